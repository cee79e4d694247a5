// C++ host-side binding of the B200 back-end: the reference's operator API for the hot path
// (x::State, x::Updater, x::VioUpdater, x::Ekf -- include/x/ekf/{state,updater,ekf}.h, include/x/vio/vio_updater.h)
// re-implemented over the C ABI of include/xb200.h.  Header-only; link with -lxb200.
//
// Scope: the filter back-end only.  The reference's VioUpdater constructor also takes Tracker / StateManager /
// TrackManager objects (vio_updater.h:45-49); those front-end components are out of scope, so the measurement
// enters at the seam their preProcess leaves behind (vio_updater.cpp:172-179): x::VioMeasurement here carries the
// five track lists + lost-feature indexes.  Matrices use Eigen when available, else the minimal column-major
// x::Matrix below (same element access syntax).
#pragma once
#include <cmath>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../xb200.h"

#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
namespace x {
using Matrix = Eigen::MatrixXd;
using Vectorx = Eigen::VectorXd;
using Vector3 = Eigen::Vector3d;
struct Quaternion : Eigen::Quaterniond { using Eigen::Quaterniond::Quaterniond; };
}  // namespace x
#else
namespace x {
// Minimal dense column-major matrix (Eigen::MatrixXd storage order) for builds without Eigen.
class Matrix {
 public:
  Matrix() = default;
  Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
  static Matrix Zero(int r, int c) { return Matrix(r, c); }
  static Matrix Identity(int r, int c) { Matrix m(r, c); for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0; return m; }
  int rows() const { return r_; }
  int cols() const { return c_; }
  size_t size() const { return d_.size(); }
  double& operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
  double operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
  double& operator()(int i) { return d_[i]; }
  double operator()(int i) const { return d_[i]; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
 private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};
using Vectorx = Matrix;
struct Vector3 {
  double v[3] = {0, 0, 0};
  Vector3() = default;
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Quaternion {  // Eigen::Quaterniond(w, x, y, z) constructor order, coeffs stored (x,y,z,w)
  double c[4] = {0, 0, 0, 1};
  Quaternion() = default;
  Quaternion(double w, double x, double y, double z) : c{x, y, z, w} {}
  double x() const { return c[0]; } double y() const { return c[1]; } double z() const { return c[2]; } double w() const { return c[3]; }
};
}  // namespace x
#endif

namespace x {

constexpr double kInvalid = -1.0;  // include/x/common/types.h:90
struct ImuNoise { double n_w = 0.0083, n_bw = 0.00083, n_a = 0.0013, n_ba = 0.00013; };  // common/types.h:65-85
struct init_bfr_mismatch {};  // include/x/ekf/ekf.h:202

using Track = std::vector<std::pair<double, double>>;  // normalised (x, y) per observation, oldest first
using TrackList = std::vector<Track>;

/** What VioUpdater::preProcess leaves behind (vio_updater.cpp:172-179). */
struct VioMeasurement {
  double timestamp = kInvalid;
  TrackList slam_trks, msckf_trks, msckf_short_trks, new_slam_std_trks, new_msckf_slam_trks;
  std::vector<unsigned int> lost_slam_trk_idxs;
};

inline void xb_throw(int rc) {
  if (rc >= 0) return;
  const std::string msg = xb_last_error();
  if (rc == XB_E_MISMATCH) throw init_bfr_mismatch{};
  if (rc == XB_E_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

/** x::SimpleState (include/x/ekf/simple_state.h:30-75, src/x/ekf/simple_state.cpp:23-65): another agent's snapshot as
 *  it arrives from the network -- dynamic state, pose window, features, anchors and the full covariance. */
class SimpleState {
 public:
  SimpleState() = delete;
  SimpleState(std::vector<double> dynamic_state, std::vector<double> positions_state, std::vector<double> orientations_state,
              std::vector<double> features_state, Matrix cov, std::vector<int> anchor_idxs)
      : dynamic_state_(std::move(dynamic_state)), positions_state_(std::move(positions_state)),
        orientations_state_(std::move(orientations_state)), features_state_(std::move(features_state)),
        anchor_idxs_(std::move(anchor_idxs)), cov_(std::move(cov)), n_poses_((int)positions_state_.size() / 3) {}
  int nPosesMax() const { return n_poses_; }
  int nFeaturesMax() const { return (int)features_state_.size() / 3; }
  const std::vector<double>& getDynamicState() const { return dynamic_state_; }
  const std::vector<double>& getPositionState() const { return positions_state_; }
  const std::vector<double>& getOrientationState() const { return orientations_state_; }
  const std::vector<double>& getFeatureState() const { return features_state_; }
  const Matrix& getCovariance() const { return cov_; }
  const std::vector<int>& getAnchorIdxs() const { return anchor_idxs_; }
  int getErrorStateSize() const { return (int)cov_.cols(); }
  Vector3 getTranslation() const { return Vector3(0.0, 0.0, 0.0); }  // simple_state.h:71 (const zero in the reference)
  /** View for the C ABI (pointers stay valid while *this lives). */
  xb_peer_state view() const {
    xb_peer_state ps{};
    ps.n_poses_max = nPosesMax(); ps.n_features_max = nFeaturesMax();
    ps.positions = positions_state_.data(); ps.orientations = orientations_state_.data();
    ps.features = features_state_.data(); ps.anchor_idxs = anchor_idxs_.data();
    ps.cov = cov_.data(); ps.cov_layout = XB_COL_MAJOR;
    return ps;
  }
 private:
  std::vector<double> dynamic_state_, positions_state_, orientations_state_, features_state_;
  std::vector<int> anchor_idxs_;
  Matrix cov_;
  int n_poses_ = -1;
};

/** include/x/vision/types.h:83-100.  The reference identifies the own track by Track::getId(); at the preProcess seam
 *  a track is identified by the list it sits in (0 = msckf_trks, 1 = msckf_short_trks) and its index there. */
struct MsckfMatch {
  std::shared_ptr<SimpleState> state;
  int uav_id = -1;
  Track received_track;
  int current_track_list = 0;
  int id_current_track = -1;
};
using MsckfMatches = std::vector<MsckfMatch>;
/** include/x/vision/types.h:102-116 */
struct SlamMatch {
  std::shared_ptr<SimpleState> state;
  int uav_id = -1;
  int current_feature_id = -1;
  int received_feature_id = -1;
};
using SlamMatches = std::vector<SlamMatch>;

/** x::State (include/x/ekf/state.h:36-337): estimates are mirrored on the host; the N x N covariance stays on the
 *  device and is fetched lazily by getCovariance(). */
class State {
 public:
  State() = default;
  State(int n_poses, int n_features) : M_(n_poses), F_(n_features), x_(XB_XVEC_LEN(n_poses, n_features), 0.0) {
    x_[9] = 1.0; x_[19] = 1.0; x_[29] = kInvalid;
  }
  double getTime() const { return x_[29]; }
  Vector3 getPosition() const { return Vector3(x_[0], x_[1], x_[2]); }
  Vector3 getVelocity() const { return Vector3(x_[3], x_[4], x_[5]); }
  Quaternion getOrientation() const { return Quaternion(x_[9], x_[6], x_[7], x_[8]); }
  Vector3 getGyroscopeBias() const { return Vector3(x_[10], x_[11], x_[12]); }
  Vector3 getAccelerometerBias() const { return Vector3(x_[13], x_[14], x_[15]); }
  Quaternion getOrientationExtrinsics() const { return Quaternion(x_[19], x_[16], x_[17], x_[18]); }
  Vector3 getPositionExtrinsics() const { return Vector3(x_[20], x_[21], x_[22]); }
  std::vector<double> getPositionArray() const { return {x_.begin() + 32, x_.begin() + 32 + 3 * M_}; }
  std::vector<double> getOrientationArray() const { return {x_.begin() + 32 + 3 * M_, x_.begin() + 32 + 7 * M_}; }
  std::vector<double> getFeatureArray() const { return {x_.begin() + 32 + 7 * M_, x_.begin() + 32 + 7 * M_ + 3 * F_}; }
  int nPosesMax() const { return M_; }
  int nFeaturesMax() const { return F_; }
  int nErrorStates() const { return XB_NERR(M_, F_); }  // state.cpp:171-175
  void setTime(double t) { x_[29] = t; }
  void setPosition(const Vector3& p) { for (int i = 0; i < 3; ++i) x_[i] = p(i); }
  void setVelocity(const Vector3& v) { for (int i = 0; i < 3; ++i) x_[3 + i] = v(i); }
  void setOrientation(const Quaternion& q) { x_[6] = q.x(); x_[7] = q.y(); x_[8] = q.z(); x_[9] = q.w(); }
  void setGyroscopeBias(const Vector3& b) { for (int i = 0; i < 3; ++i) x_[10 + i] = b(i); }
  void setAccelerometerBias(const Vector3& b) { for (int i = 0; i < 3; ++i) x_[13 + i] = b(i); }
  void setOrientationExtrinsics(const Quaternion& q) { x_[16] = q.x(); x_[17] = q.y(); x_[18] = q.z(); x_[19] = q.w(); }
  void setPositionExtrinsics(const Vector3& p) { for (int i = 0; i < 3; ++i) x_[20 + i] = p(i); }
  void setImu(double t, unsigned seq, const Vector3& w, const Vector3& a) {  // state.cpp:145-151
    x_[29] = t; x_[30] = seq;
    for (int i = 0; i < 3; ++i) { x_[23 + i] = w(i); x_[26 + i] = a(i); }
  }
  /** Host copy of the covariance (column-major like Eigen).  For states returned by Ekf it is downloaded on demand. */
  const Matrix& getCovariance() const {
    if (cov_.size() == 0 && flt_) {
      cov_ = Matrix(nErrorStates(), nErrorStates());
      xb_throw(xb_ekf_get_covariance(flt_, slot_, cov_.data(), XB_COL_MAJOR));
    }
    return cov_;
  }
  void setCovariance(const Matrix& c) { cov_ = c; flt_ = nullptr; }
  std::vector<double>& xvec() { return x_; }
  const std::vector<double>& xvec() const { return x_; }

 private:
  friend class Ekf;
  int M_ = 0, F_ = 0;
  std::vector<double> x_;
  mutable Matrix cov_;
  xb_filter* flt_ = nullptr;  // device-resident covariance: (filter, ring slot)
  int slot_ = -1;
};

/** x::Updater (include/x/ekf/updater.h:37-232): the abstract operator API.  The template method `update` and the
 *  Kalman arithmetic run on the device; a subclass supplies its measurement either natively (VioUpdater) or as dense
 *  host matrices through applyUpdate / applyCI, exactly as in the reference. */
class Updater {
 public:
  virtual ~Updater() = default;
  virtual double getTime() const = 0;
  /** updater.cpp:39-115 on the filter's work state. */
  void update(xb_filter* f) {
    if (deviceNative()) { xb_throw(xb_updater_update(f)); return; }
    throw std::logic_error("generic host-matrix updaters call applyUpdate() on the work state themselves");
  }
 protected:
  int iekf_iter_{1};
  virtual bool deviceNative() const { return false; }
  /** updater.cpp:117-141; H is m x N (column-major x::Matrix), res m x 1, R diagonal m x m. */
  void applyUpdate(xb_filter* f, const Matrix& H, const Matrix& res, const Matrix& R, Matrix& correction_total,
                   bool cov_update = true) {
    const int m = H.rows(), n = H.cols();
    std::vector<double> h((size_t)m * n), r(m), rd(m);
    for (int i = 0; i < m; ++i) { r[i] = res(i, 0); rd[i] = R(i, i); for (int j = 0; j < n; ++j) h[(size_t)i * n + j] = H(i, j); }
    std::vector<double> ct(n);
    for (int j = 0; j < n; ++j) ct[j] = correction_total(j, 0);
    xb_throw(xb_updater_apply_update(f, h.data(), r.data(), rd.data(), m, ct.data(), cov_update ? 1 : 0));
    for (int j = 0; j < n; ++j) correction_total(j, 0) = ct[j];
  }
  friend class Ekf;
};

/** x::VioUpdater (include/x/vio/vio_updater.h:35-335), back-end part. */
class VioUpdater : public Updater {
 public:
  VioUpdater(double sigma_img, double sigma_range, double rho_0, double sigma_rho_0, int min_track_length,
             double sigma_landmark = 0.0, double ci_msckf_w = -1.0, double ci_slam_w = -1.0, int iekf_iter = 1)
      : sigma_img_(sigma_img), sigma_range_(sigma_range), rho_0_(rho_0), sigma_rho_0_(sigma_rho_0),
        min_track_length_(min_track_length), sigma_landmark_(sigma_landmark), ci_msckf_w_(ci_msckf_w), ci_slam_w_(ci_slam_w) {
    iekf_iter_ = iekf_iter;
  }
  void setMeasurement(const VioMeasurement& m) { measurement_ = m; }  // vio_updater.cpp:122-124
  /** MULTI_UAV build: what Tracker::getMsckfMatches / getSlamMatches hand over (vio_updater.cpp:185,212). */
  void setMsckfMatches(const MsckfMatches& m) { msckf_matches_ = m; }
  void setSlamMatches(const SlamMatches& m) { slam_matches_ = m; }
  /** Updater::collaborativeUpdate (updater.cpp:22-36) through Ekf::processOthersMeasurement. */
  int collaborate(xb_filter* f, double timestamp, double* xvec_out) {
    std::vector<const SimpleState*> uniq;
    std::vector<xb_slam_match> cm;
    for (const auto& m : slam_matches_) {
      size_t k = 0;
      while (k < uniq.size() && uniq[k] != m.state.get()) ++k;
      if (k == uniq.size()) uniq.push_back(m.state.get());
      cm.push_back({(int)k, m.current_feature_id, m.received_feature_id});
    }
    std::vector<xb_peer_state> ps;
    for (const auto* u : uniq) ps.push_back(u->view());
    const int rc = xb_ekf_process_others(f, timestamp, ps.data(), (int)ps.size(), cm.data(), (int)cm.size(), xvec_out);
    slam_matches_.clear();
    return rc;
  }
  double getTime() const override { return measurement_.timestamp; }  // vio_updater.h:60
  void fillConfig(xb_config& c) const {
    c.sigma_img = sigma_img_; c.sigma_range = sigma_range_; c.rho_0 = rho_0_; c.sigma_rho_0 = sigma_rho_0_;
    c.min_track_length = min_track_length_; c.sigma_landmark = sigma_landmark_; c.ci_msckf_w = ci_msckf_w_;
    c.ci_slam_w = ci_slam_w_; c.iekf_iter = iekf_iter_;
  }
  /** Marshal the five track lists to the C ABI (the only host->device traffic of an update). */
  void upload(xb_filter* f) const {
    struct Csr { std::vector<int> off; std::vector<double> obs; };
    auto pack = [](const TrackList& tl, Csr& c, xb_track_list& out) {
      c.off.assign(1, 0);
      for (const auto& t : tl) { for (const auto& o : t) { c.obs.push_back(o.first); c.obs.push_back(o.second); } c.off.push_back((int)c.obs.size() / 2); }
      out.n_tracks = (int)tl.size(); out.off = c.off.data(); out.obs = c.obs.data();
    };
    Csr a, b, c, d, e;
    xb_measurement m{};
    m.timestamp = measurement_.timestamp;
    pack(measurement_.slam_trks, a, m.slam);
    pack(measurement_.msckf_trks, b, m.msckf);
    pack(measurement_.msckf_short_trks, c, m.msckf_short);
    pack(measurement_.new_slam_std_trks, d, m.new_slam_std);
    pack(measurement_.new_msckf_slam_trks, e, m.new_msckf_slam);
    std::vector<int> lost(measurement_.lost_slam_trk_idxs.begin(), measurement_.lost_slam_trk_idxs.end());
    m.n_lost = (int)lost.size(); m.lost_slam_idxs = lost.data();
    xb_throw(xb_vio_set_measurement(f, &m));
    if (!msckf_matches_.empty()) {  // consumed by the next Updater::update (msckf_update.cpp:96-139)
      std::vector<const SimpleState*> uniq;
      std::vector<xb_msckf_match> cm;
      std::vector<std::vector<double>> obs;
      for (const auto& mm : msckf_matches_) {
        size_t k = 0;
        while (k < uniq.size() && uniq[k] != mm.state.get()) ++k;
        if (k == uniq.size()) uniq.push_back(mm.state.get());
        obs.emplace_back();
        for (const auto& o : mm.received_track) { obs.back().push_back(o.first); obs.back().push_back(o.second); }
        cm.push_back({(int)k, mm.current_track_list, mm.id_current_track, (int)mm.received_track.size(), nullptr});
      }
      for (size_t j = 0; j < cm.size(); ++j) cm[j].obs = obs[j].data();
      std::vector<xb_peer_state> ps;
      for (const auto* u : uniq) ps.push_back(u->view());
      xb_throw(xb_vio_set_msckf_matches(f, ps.data(), (int)ps.size(), cm.data(), (int)cm.size()));
      msckf_matches_.clear();  // preProcess replaces the list on every update (vio_updater.cpp:185)
    }
  }
 protected:
  bool deviceNative() const override { return true; }
 private:
  VioMeasurement measurement_;
  mutable MsckfMatches msckf_matches_;
  SlamMatches slam_matches_;
  double sigma_img_, sigma_range_, rho_0_, sigma_rho_0_;
  int min_track_length_;
  double sigma_landmark_, ci_msckf_w_, ci_slam_w_;
};

/** x::Ekf (include/x/ekf/ekf.h:53-195). */
class Ekf {
 public:
  explicit Ekf(Updater& updater) : updater_(updater) {}
  ~Ekf() { if (f_) xb_destroy(f_); }
  Ekf(const Ekf&) = delete;
  /** ekf.cpp:32-41 */
  void set(const Updater&, const Vector3& g, const ImuNoise& noise, int state_buffer_sz, const State& default_state,
           double a_m_max, unsigned delta_seq_imu, const double& time_margin_bfr, int max_tracks = 1024, int device = 0) {
    xb_config c;
    xb_default_config(&c);
    c.n_poses_max = default_state.nPosesMax(); c.n_features_max = default_state.nFeaturesMax();
    c.n_slots = state_buffer_sz; c.device = device; c.max_tracks = max_tracks;
    for (int i = 0; i < 3; ++i) c.g[i] = g(i);
    c.n_w = noise.n_w; c.n_bw = noise.n_bw; c.n_a = noise.n_a; c.n_ba = noise.n_ba;
    c.a_m_max = a_m_max; c.delta_seq_imu = delta_seq_imu; c.time_margin = time_margin_bfr;
    if (auto* v = dynamic_cast<VioUpdater*>(&updater_)) v->fillConfig(c);
#ifdef MULTI_UAV
    c.multi_uav = 1;  // the reference selects this flow at compile time (CMakeLists.txt: -DMULTI_UAV)
#endif
    if (f_) { xb_destroy(f_); f_ = nullptr; }
    xb_throw(xb_create(&c, &f_));
    M_ = c.n_poses_max; F_ = c.n_features_max;
  }
  /** ekf.cpp:43-64 */
  void initializeFromState(const State& s) {
    if (!f_) throw std::runtime_error("The EKF state buffer must have non-zero size.");
    if (s.nPosesMax() != M_ || s.nFeaturesMax() != F_ || s.getCovariance().rows() != XB_NERR(M_, F_)) throw init_bfr_mismatch{};
    xb_throw(xb_ekf_initialize_from_state(f_, s.xvec().data(), s.getCovariance().data(), XB_COL_MAJOR));
  }
  /** ekf.cpp:66-140 */
  std::optional<State> processImu(double timestamp, unsigned seq, const Vector3& w_m, const Vector3& a_m) {
    std::lock_guard<std::mutex> lk(mutex_);
    State out(M_, F_);
    const double w[3] = {w_m(0), w_m(1), w_m(2)}, a[3] = {a_m(0), a_m(1), a_m(2)};
    const int rc = xb_ekf_process_imu(f_, timestamp, seq, w, a, out.xvec().data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    out.flt_ = f_; out.slot_ = xb_ekf_newest_slot(f_);
    return out;
  }
  /** ekf.cpp:179-213 */
  std::optional<State> processUpdateMeasurement() {
    std::lock_guard<std::mutex> lk(mutex_);
    auto* v = dynamic_cast<VioUpdater*>(&updater_);
    if (!v) throw std::logic_error("Ekf::processUpdateMeasurement needs a device-native updater");
    v->upload(f_);
    State out(M_, F_);
    const int rc = xb_ekf_process_update(f_, out.xvec().data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    return out;
  }
  /** ekf.cpp:143-176 (MULTI_UAV): SLAM-SLAM covariance-intersection update against the peers set on the updater. */
  std::optional<State> processOthersMeasurement(double timestamp) {
    std::lock_guard<std::mutex> lk(mutex_);
    auto* v = dynamic_cast<VioUpdater*>(&updater_);
    if (!v) throw std::logic_error("Ekf::processOthersMeasurement needs a device-native updater");
    State out(M_, F_);
    const int rc = v->collaborate(f_, timestamp, out.xvec().data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    return out;
  }
  void lock() { mutex_.lock(); }      // ekf.h:128
  void unlock() { mutex_.unlock(); }  // ekf.h:133
  xb_filter* handle() { return f_; }

 private:
  Updater& updater_;
  xb_filter* f_ = nullptr;
  int M_ = 0, F_ = 0;
  std::mutex mutex_;
};

}  // namespace x
