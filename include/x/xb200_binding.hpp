// C++ host side of the B200 back end: the reference's operator API for the hot path -- x::State, x::Updater,
// x::VioUpdater, x::Ekf, x::StateManager, x::SimpleState and the vision PODs they exchange (include/x/ekf/{state,
// updater,ekf,simple_state}.h, include/x/vio/{vio_updater,state_manager,types}.h, include/x/vision/{types,track,
// feature}.h of jpl-x/x_multi_agent) -- with the same class names, method signatures, ownership rules and error
// behaviour, implemented over the C ABI of include/xb200.h.  Header-only; link with -lxb200.  The per-file headers
// next to this one (x/ekf/ekf.h, x/vio/vio_updater.h, ...) forward here, so the reference's own
// `#include "x/ekf/ekf.h"` lines keep working.
//
// What is different underneath:
//  * the N x N covariance of a State that came out of x::Ekf stays on the GPU; getCovariance()/getCovarianceRef()
//    fetch it on demand (and getCovarianceRef() makes the host copy authoritative, as a mutable reference must);
//  * Updater::update(State&) is the reference's template method, driving the device stage by stage through the
//    same virtuals; x::VioUpdater's constructUpdate hands back a 1 x 1 token instead of the stacked Jacobian (the
//    compressed measurement never leaves the device) and Updater::applyUpdate recognises it.  A user-defined
//    Updater that returns real host matrices goes through the dense entry point (xb_updater_apply_update);
//  * Tracker / TrackManager (image front end) are opaque here: TrackManager only carries the
//    five normalised track lists VioUpdater::preProcess reads (vio_updater.cpp:172-179).
// Matrix types are Eigen's (the reference API is Eigen-typed).
#pragma once
#include <Eigen/Dense>

#include <cmath>
#include <exception>
#include <iomanip>
#include <memory>
#include <mutex>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../xb200.h"

namespace x {

// ---- include/x/common/types.h ----------------------------------------------------------------------------------
enum { kIdxP = 0, kIdxV = 3, kIdxQ = 6, kIdxBw = 9, kIdxBa = 12, kSizeCoreErr = 15, kSizeClone = 15 };
using Vector3 = Eigen::Vector3d;
using Quaternion = Eigen::Quaterniond;
using Matrix = Eigen::MatrixXd;
using Vectorx = Eigen::VectorXd;
using Matrix3 = Eigen::Matrix3d;
using Matrix4 = Eigen::Matrix4d;
constexpr double kInvalid = -1.0;
/** Ekf::mutex_ (ekf.h:185).  Recursive: the reference's VIO::initAtTime calls Ekf::initializeFromState between Ekf::lock()
 *  and Ekf::unlock() (vio.cpp:55-109), and the entry points of this binding take the lock themselves. */
using EkfMutex = std::recursive_mutex;
struct ImuNoise {
  double n_w = 0.0083, n_bw = 0.00083, n_a = 0.0013;
  double n_ba = 0.00013;  // the reference's literal `00013` is octal (= 11); VIO::setUp overrides it (vio.cpp:181-185)
};
class init_bfr_mismatch : std::exception {};  // include/x/ekf/ekf.h:202

inline void xb_throw(int rc) {
  if (rc >= 0) return;
  const std::string msg = xb_last_error();
  if (rc == XB_E_MISMATCH) throw init_bfr_mismatch{};
  if (rc == XB_E_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

// ---- include/x/vision/{feature,track,types}.h (the part the back end reads) ---------------------------------------
class Feature {
 public:
  Feature() = default;
  Feature(const double& timestamp, double x, double y, double intensity = -1.0)
      : timestamp_(timestamp), x_(x), y_(y), intensity_(intensity) {}
  Feature(const double& timestamp, unsigned int frame_number, double x, double y, double x_dist, double y_dist,
          double intensity = -1.0)
      : timestamp_(timestamp), x_(x), y_(y), intensity_(intensity), x_dist_(x_dist), y_dist_(y_dist), frame_number_(frame_number) {}
  void setX(const double x) { x_ = x; }
  void setY(const double y) { y_ = y; }
  void setXDist(const double x_dist) { x_dist_ = x_dist; }
  void setYDist(const double y_dist) { y_dist_ = y_dist; }
  [[nodiscard]] double getXDist() const { return x_dist_; }   // distorted (measured) pixel coordinates
  [[nodiscard]] double getYDist() const { return y_dist_; }
  [[nodiscard]] double getTimestamp() const { return timestamp_; }
  [[nodiscard]] double getX() const { return x_; }   // normalised image coordinates
  [[nodiscard]] double getY() const { return y_; }
  [[nodiscard]] double getIntensity() const { return intensity_; }

 private:
  double timestamp_{0}, x_{0}, y_{0}, intensity_{-1}, x_dist_{0}, y_dist_{0};
  unsigned int frame_number_{0};
};
/** include/x/vision/types.h:39-42 */
struct Match {
  Feature previous;
  Feature current;
};
using MatchList = std::vector<Match>;
class Track : public std::vector<Feature> {
 public:
  Track() : std::vector<Feature>() { id_ = ++counter(); }
  Track(const size_type count, const Feature& feature, unsigned long long id) : std::vector<Feature>(count, feature), id_(id) {}
  explicit Track(unsigned long long id) : id_(id) {}
  [[nodiscard]] unsigned long long getId() const { return id_; }

 private:
  unsigned long long id_{0};
  static unsigned long long& counter() { static unsigned long long c = 0; return c; }
};
using TrackList = std::vector<Track>;
using TrackPtr = std::shared_ptr<Track>;
using TrackListPtr = std::vector<std::shared_ptr<Track>>;   // include/x/vision/types.h:55
using uniqueId = unsigned long long;
struct Attitude {
  double ax = 0, ay = 0, az = 0, aw = 0;
  Attitude(double ax, double ay, double az, double aw) : ax(ax), ay(ay), az(az), aw(aw) {}
  Attitude() = default;
};
struct Translation {
  double tx = 0, ty = 0, tz = 0;
  Translation(double tx, double ty, double tz) : tx(tx), ty(ty), tz(tz) {}
  Translation() = default;
};
using AttitudeList = std::vector<Attitude>;
using TranslationList = std::vector<Translation>;

/** include/x/vio/types.h:223-243: laser range finder. */
struct RangeMeasurement {
  double timestamp{-1.0};   // kInvalid
  double range{0.0};
  Feature img_pt{Feature()};
  Feature img_pt_n{Feature()};
};
/** include/x/vio/types.h:250-254: sun sensor. */
struct SunAngleMeasurement {
  double timestamp{-1.0};
  double x_angle{0.0};
  double y_angle{0.0};
};
class TiledImage;
/** include/x/vio/types.h:264-324.  `matches` / `image` are the front end's members: a measurement built with them (the
 *  reference's constructor) makes VioUpdater::preProcess sort the matches into tracks itself (vio_updater.cpp:142-179);
 *  without them the track lists are the ones the caller left in the TrackManager (manageTracks / setTracks). */
struct VioMeasurement {
  double timestamp{0};
  unsigned int seq{0};
  MatchList matches;
  unsigned int n_tiles_h{1}, n_tiles_w{1};   // tile grid of `image` (TiledImage without the pixels)
  bool from_front_end{false};
  RangeMeasurement range;
  SunAngleMeasurement sun_angle;
  VioMeasurement() = default;
  VioMeasurement(const double& timestamp, const unsigned int seq) : timestamp{timestamp}, seq{seq} {}
  VioMeasurement(const double& timestamp, const unsigned int seq, RangeMeasurement range, const SunAngleMeasurement& sun_angle)
      : timestamp{timestamp}, seq{seq}, range{std::move(range)}, sun_angle{sun_angle} {}
  /** types.h:300-305 (defined below TiledImage) */
  VioMeasurement(const double& timestamp, const unsigned int seq, MatchList matches, const TiledImage& image,
                 RangeMeasurement range, const SunAngleMeasurement& sun_angle);
};

/** x::Params (include/x/vio/types.h:33-160).  The tracker / place-recognition members are carried so that parameter files
 *  and callers of the reference keep working; the back end reads the filter, camera and track-management members. */
struct Params {
  Vector3 p{0, 0, 0}, v{0, 0, 0};
  Quaternion q{1, 0, 0, 0};
  Vector3 b_w{0, 0, 0}, b_a{0, 0, 0};
  Vector3 sigma_dp{0, 0, 0}, sigma_dv{0, 0, 0}, sigma_dtheta{0, 0, 0}, sigma_dbw{0, 0, 0}, sigma_dba{0, 0, 0};
  double cam_fx{0}, cam_fy{0}, cam_cx{0}, cam_cy{0}, cam_s{0};
  int img_height{0}, img_width{0};
  Vector3 p_ic{0, 0, 0};
  Quaternion q_ic{1, 0, 0, 0};
  double sigma_img{0};
  double sigma_range{0};
  Quaternion q_sc{1, 0, 0, 0};
  Vector3 w_s{0, 0, 1};
  double n_a{0}, n_ba{0}, n_w{0}, n_bw{0};
  int fast_detection_delta{0};
  bool non_max_supp{false};
  int block_half_length{0}, margin{0}, n_feat_min{0}, outlier_method{0};
  double outlier_param1{0}, outlier_param2{0};
  int n_tiles_h{1}, n_tiles_w{1}, max_feat_per_tile{0};
  double time_offset{0};
  std::string vocabulary_path;
  double sigma_landmark{0};
  float descriptor_scale_factor{0};
  int descriptor_pyramid{0}, descriptor_patch_size{0};
  double ci_msckf_w{-1.0}, ci_slam_w{-1.0};
  int desc_type{0};
  double pr_score_thr{0}, pr_desc_ratio_thr{0}, pr_desc_min_distance{0};
  int max_level{0};
  double min_eig_thr{0};
  int win_size_w{0}, win_size_h{0};
  int n_poses_max = 15;
  int n_slam_features_max = 15;
  double rho_0 = 0.5;
  double sigma_rho_0 = 0.25;
  int iekf_iter = 1;
  double msckf_baseline = 10;
  int min_track_length = 15;
  Vector3 g{0, 0, -9.81};
  bool self_init_start_ = false;
  int state_buffer_size = 250;
};

/** x::SimpleState (include/x/ekf/simple_state.h:30-75): another agent's snapshot as it arrives from the network. */
class SimpleState {
 public:
  SimpleState() = delete;
  SimpleState(Vectorx dynamic_state, const Vectorx& positions_state, Vectorx orientations_state, Vectorx features_state,
              Matrix cov, std::vector<int> anchor_idxs)
      : dynamic_state_(std::move(dynamic_state)), positions_state_(positions_state),
        orientations_state_(std::move(orientations_state)), features_state_(std::move(features_state)),
        anchor_idxs_(std::move(anchor_idxs)), cov_(std::move(cov)), n_poses_(static_cast<int>(positions_state_.rows()) / 3) {}
  [[nodiscard]] int nPosesMax() const { return n_poses_; }
  [[nodiscard]] int nFeaturesMax() const { return static_cast<int>(features_state_.rows()) / 3; }
  [[nodiscard]] Vectorx getDynamicState() const { return dynamic_state_; }
  [[nodiscard]] Vectorx getPositionState() const { return positions_state_; }
  [[nodiscard]] Vectorx getOrientationState() const { return orientations_state_; }
  [[nodiscard]] Vectorx getFeatureState() const { return features_state_; }
  [[nodiscard]] Matrix getCovariance() const { return cov_; }
  [[nodiscard]] std::vector<int> getAnchorIdxs() const { return anchor_idxs_; }
  [[nodiscard]] int getErrorStateSize() const { return static_cast<int>(cov_.cols()); }
  [[nodiscard]] Vector3 getTranslation() const { return Vector3(0.0, 0.0, 0.0); }
  [[nodiscard]] int getAnchorIdat(int id) const { return anchor_idxs_[id]; }
  /** View for the C ABI (pointers stay valid while *this lives). */
  xb_peer_state view() const {
    xb_peer_state ps{};
    ps.n_poses_max = nPosesMax();
    ps.n_features_max = nFeaturesMax();
    ps.positions = positions_state_.data();
    ps.orientations = orientations_state_.data();
    ps.features = features_state_.data();
    ps.anchor_idxs = anchor_idxs_.data();
    ps.cov = cov_.data();
    ps.cov_layout = XB_COL_MAJOR;
    return ps;
  }

 private:
  const Vectorx dynamic_state_, positions_state_, orientations_state_, features_state_;
  const std::vector<int> anchor_idxs_;
  const Matrix cov_;
  const int n_poses_ = -1;
};
/** include/x/vision/types.h:83-100 */
struct MsckfMatch {
  std::shared_ptr<SimpleState> state;
  int uav_id = -1;
  TrackPtr received_track_ptr;
  uniqueId id_current_track = static_cast<uniqueId>(-1);
  uniqueId id_received_track = static_cast<uniqueId>(-1);
  MsckfMatch(int uav_id, uniqueId id_current_track, uniqueId id_received_track, TrackPtr received_track,
             std::shared_ptr<SimpleState> state)
      : state(std::move(state)), uav_id(uav_id), received_track_ptr(std::move(received_track)),
        id_current_track(id_current_track), id_received_track(id_received_track) {}
};
using MsckfMatches = std::vector<MsckfMatch>;
/** include/x/vision/types.h:102-116 */
struct SlamMatch {
  std::shared_ptr<SimpleState> state;
  int uav_id = -1;
  int current_feature_id = -1;
  int received_feature_id = -1;
  SlamMatch(int uav_id, int current_feature_id, int received_feature_id, std::shared_ptr<SimpleState> state)
      : state(std::move(state)), uav_id(uav_id), current_feature_id(current_feature_id),
        received_feature_id(received_feature_id) {}
  SlamMatch() = delete;
};
using SlamMatches = std::vector<SlamMatch>;

// ---- x::State (include/x/ekf/state.h:36-337) ---------------------------------------------------------------------
/** Host object with the reference's members and accessors.  A State handed out by x::Ekf additionally refers to the
 *  device-resident copy it was read from (ring slot or work state): estimates are mirrored on the host, the covariance
 *  is fetched on first use.  A State built by the caller is a plain host object. */
class State {
 public:
  State() = default;
  State(int n_poses, int n_features) {   // state.cpp:23-37
    p_array_ = Matrix::Zero(n_poses * 3, 1);
    q_array_ = Matrix::Zero(n_poses * 4, 1);
    f_array_ = Matrix::Zero(n_features * 3, 1);
    const int n = kSizeCoreErr + n_poses * 6 + n_features * 3;
    cov_ = Matrix::Identity(n, n);
  }
  State(const double time, const unsigned int seq, const Vector3& p, const Vector3& v, const Quaternion& q, const Vector3& b_w,
        const Vector3& b_a, const Matrix& p_array, const Matrix& q_array, const Matrix& f_array, const Matrix& cov,
        const Quaternion& q_ic, const Vector3& p_ic, const Vector3& w_m, const Vector3& a_m)
      : time_{time}, seq_{seq}, p_{p}, v_{v}, q_{q}, b_w_{b_w}, b_a_{b_a}, p_array_{p_array}, q_array_{q_array},
        f_array_{f_array}, cov_{cov}, q_ic_{q_ic}, p_ic_{p_ic}, w_m_{w_m}, a_m_{a_m} {}

  [[nodiscard]] double getTime() const { return time_; }
  [[nodiscard]] unsigned int getSeq() const { return seq_; }
  [[nodiscard]] Vector3 getPosition() const { return p_; }
  [[nodiscard]] Vector3 getVelocity() const { return v_; }
  [[nodiscard]] Quaternion getOrientation() const { return q_; }
  [[nodiscard]] Matrix getPositionArray() const { return p_array_; }
  [[nodiscard]] Matrix getOrientationArray() const { return q_array_; }
  [[nodiscard]] Matrix getFeatureArray() const { return f_array_; }
  [[nodiscard]] Quaternion getOrientationExtrinsics() const { return q_ic_; }
  [[nodiscard]] Vector3 getPositionExtrinsics() const { return p_ic_; }
  [[nodiscard]] Eigen::VectorXd getDynamicStates() const {   // state.cpp:87-99
    Eigen::VectorXd d(kSizeCoreErr + 1);
    d.segment(0, 3) = p_;
    d.segment(3, 3) = v_;
    d.segment(6, 4) = q_.coeffs();
    d.segment(10, 3) = b_w_;
    d.segment(13, 3) = b_a_;
    return d;
  }
  [[nodiscard]] Matrix getCovariance() const { fetchCov(); return cov_; }
  [[nodiscard]] Matrix getPoseCovariance() const {   // state.cpp:103-120
    fetchCov();
    Matrix pc(6, 6);
    pc.topLeftCorner(3, 3) = cov_.topLeftCorner(3, 3);
    pc.bottomRightCorner(3, 3) = cov_.block(kIdxQ, kIdxQ, 3, 3);
    pc.topRightCorner(3, 3) = cov_.block(kIdxP, kIdxQ, 3, 3);
    pc.bottomLeftCorner(3, 3) = cov_.block(kIdxQ, kIdxP, 3, 3);
    return pc;
  }
  [[nodiscard]] Matrix getDynamicCovariance() const { fetchCov(); return cov_.topLeftCorner(kSizeCoreErr, kSizeCoreErr); }
  /** Mutable host reference (state.h:115): the host copy becomes the authoritative one; the next device operation on
   *  this State uploads it. */
  Matrix& getCovarianceRef() { fetchCov(); dev_.kind = kHost; return cov_; }

  void setTime(const double time) { time_ = time; est_dirty_ = true; }
  void setPositionArray(const Matrix& p_array) { p_array_ = p_array; est_dirty_ = true; }
  void setOrientationArray(const Matrix& q_array) { q_array_ = q_array; est_dirty_ = true; }
  void setFeatureArray(const Matrix& f_array) { f_array_ = f_array; est_dirty_ = true; }
  void setCovariance(const Matrix& cov) { cov_ = cov; cov_valid_ = true; dev_.kind = kHost; }
  void setImu(double time, unsigned int seq, const Vector3& w_m, const Vector3& a_m) {   // state.cpp:145-151
    time_ = time; seq_ = seq; w_m_ = w_m; a_m_ = a_m; est_dirty_ = true;
  }
  void setStaticStatesFrom(const State& s) {   // state.cpp:153-161
    b_w_ = s.b_w_; b_a_ = s.b_a_; q_ic_ = s.q_ic_; p_ic_ = s.p_ic_;
    p_array_ = s.p_array_; q_array_ = s.q_array_; f_array_ = s.f_array_; est_dirty_ = true;
  }
  void reset() { time_ = kInvalid; }
  [[nodiscard]] int nPosesMax() const { return static_cast<int>(p_array_.rows() / 3); }
  [[nodiscard]] int nFeaturesMax() const { return static_cast<int>(f_array_.rows() / 3); }
  [[nodiscard]] int nErrorStates() const {
    return kSizeCoreErr + static_cast<int>(p_array_.rows()) + static_cast<int>(q_array_.rows() / 4) * 3 +
           static_cast<int>(f_array_.rows());
  }
  void computeUnbiasedImuMeasurements(Vector3& e_w, Vector3& e_a) const { e_w = w_m_ - b_w_; e_a = a_m_ - b_a_; }
  [[nodiscard]] Attitude computeCameraAttitude() const {
    const Quaternion quat = q_.normalized() * q_ic_.normalized();
    return {quat.x(), quat.y(), quat.z(), quat.w()};
  }
  [[nodiscard]] Vector3 computeCameraPosition() const { return p_ + q_.normalized().toRotationMatrix() * p_ic_; }
  [[nodiscard]] Quaternion computeCameraOrientation() const { return q_.normalized() * q_ic_.normalized(); }
  /** state.cpp:197-249: additive correction + quaternion product with the exact angle-axis error quaternion. */
  void correct(const Eigen::VectorXd& correction) {
    const int n_p = static_cast<int>(p_array_.rows()), n_f = static_cast<int>(f_array_.rows());
    p_ += correction.segment(kIdxP, 3);
    v_ += correction.segment(kIdxV, 3);
    b_w_ += correction.segment(kIdxBw, 3);
    b_a_ += correction.segment(kIdxBa, 3);
    for (int i = 0; i < n_p; ++i) p_array_(i, 0) += correction(kSizeCoreErr + i);
    for (int i = 0; i < n_f; ++i) f_array_(i, 0) += correction(kSizeCoreErr + 2 * n_p + i);
    const Vector3 dth = correction.segment(kIdxQ, 3);
    q_ = (q_ * errorQuat(dth)).normalized();
    for (int i = 0; i < n_p / 3; ++i) {
      const Vector3 d = correction.segment(kSizeCoreErr + n_p + 3 * i, 3);
      Quaternion q_i(q_array_(4 * i + 3, 0), q_array_(4 * i, 0), q_array_(4 * i + 1, 0), q_array_(4 * i + 2, 0));
      q_i = q_i * errorQuat(d);
      q_i.normalize();
      q_array_(4 * i, 0) = q_i.x(); q_array_(4 * i + 1, 0) = q_i.y(); q_array_(4 * i + 2, 0) = q_i.z(); q_array_(4 * i + 3, 0) = q_i.w();
    }
    est_dirty_ = true;
  }
  [[nodiscard]] std::string toString() const {
    std::stringstream s;
    s << "Timestamp: " << time_ << "\n" << std::setprecision(6) << "p [x,y,z]: " << p_.transpose() << "\nv [x,y,z]: "
      << v_.transpose() << "\nq [w,x,y,z]: " << q_.w() << " " << q_.x() << " " << q_.y() << " " << q_.z() << "\nb_w [x,y,z]: "
      << b_w_.transpose() << "\nb_a [x,y,z]: " << b_a_.transpose() << "\n";
    return s.str();
  }

  // ---- device side (not part of the reference API) ----
  /** Flat estimates in the layout of include/xb200.h (XV_*). */
  std::vector<double> xvec() const {
    const int M = nPosesMax(), F = nFeaturesMax();
    std::vector<double> x(XB_XVEC_LEN(M, F), 0.0);
    for (int i = 0; i < 3; ++i) { x[i] = p_(i); x[3 + i] = v_(i); x[10 + i] = b_w_(i); x[13 + i] = b_a_(i); x[20 + i] = p_ic_(i); x[23 + i] = w_m_(i); x[26 + i] = a_m_(i); }
    x[6] = q_.x(); x[7] = q_.y(); x[8] = q_.z(); x[9] = q_.w();
    x[16] = q_ic_.x(); x[17] = q_ic_.y(); x[18] = q_ic_.z(); x[19] = q_ic_.w();
    x[29] = time_; x[30] = seq_;
    for (int i = 0; i < 3 * M; ++i) x[32 + i] = p_array_(i, 0);
    for (int i = 0; i < 4 * M; ++i) x[32 + 3 * M + i] = q_array_(i, 0);
    for (int i = 0; i < 3 * F; ++i) x[32 + 7 * M + i] = f_array_(i, 0);
    return x;
  }
  void setFromXvec(const double* x, int M, int F) {
    p_ = Vector3(x[0], x[1], x[2]); v_ = Vector3(x[3], x[4], x[5]); q_ = Quaternion(x[9], x[6], x[7], x[8]);
    b_w_ = Vector3(x[10], x[11], x[12]); b_a_ = Vector3(x[13], x[14], x[15]);
    q_ic_ = Quaternion(x[19], x[16], x[17], x[18]); p_ic_ = Vector3(x[20], x[21], x[22]);
    w_m_ = Vector3(x[23], x[24], x[25]); a_m_ = Vector3(x[26], x[27], x[28]);
    time_ = x[29]; seq_ = static_cast<unsigned int>(x[30]);
    p_array_.resize(3 * M, 1); q_array_.resize(4 * M, 1); f_array_.resize(3 * F, 1);
    for (int i = 0; i < 3 * M; ++i) p_array_(i, 0) = x[32 + i];
    for (int i = 0; i < 4 * M; ++i) q_array_(i, 0) = x[32 + 3 * M + i];
    for (int i = 0; i < 3 * F; ++i) f_array_(i, 0) = x[32 + 7 * M + i];
    est_dirty_ = false;
  }
  /** Is this State's covariance the device's work covariance of filter f (as left by Ekf / a previous stage)? */
  bool boundToWork(const xb_filter* f) const { return dev_.kind == kWork && dev_.f == f; }

 private:
  friend class Ekf;
  friend class Updater;
  friend class VioUpdater;
  friend class StateManager;
  enum { kHost = 0, kSlot = 1, kWork = 2 };
  struct DevRef {
    int kind = kHost;
    xb_filter* f = nullptr;
    EkfMutex* mtx = nullptr;  // Ekf::mutex_: downloads from a ring slot take it
    int slot = -1, serial = -1;
    double time = kInvalid;
  };
  static Quaternion errorQuat(const Vector3& d) {   // state.cpp:273-283
    const double n = d.norm();
    if (n == 0.0) return Quaternion::Identity();
    return Quaternion(Eigen::AngleAxisd(n, d / n));
  }
  /** Lazy download of a device-resident covariance; fails loudly when the ring slot has been rewritten meanwhile. */
  void fetchCov() const {
    if (cov_valid_ || dev_.kind == kHost) return;
    const int n = nErrorStates();
    cov_.resize(n, n);
    if (dev_.kind == kWork) {
      xb_throw(xb_work_get(dev_.f, nullptr, cov_.data(), XB_COL_MAJOR));
    } else {
      std::unique_lock<EkfMutex> lk;
      if (dev_.mtx) lk = std::unique_lock<EkfMutex>(*dev_.mtx);
      double t = kInvalid;
      int serial = -1;
      xb_throw(xb_ekf_slot_info(dev_.f, dev_.slot, &t, &serial));
      if (serial != dev_.serial || t != dev_.time)
        throw std::runtime_error("x::State: the ring-buffer slot this state was read from has been rewritten; "
                                 "its covariance is no longer on the device (copy it with getCovariance() earlier)");
      xb_throw(xb_ekf_get_covariance(dev_.f, dev_.slot, cov_.data(), XB_COL_MAJOR));
    }
    cov_valid_ = true;
  }
  void bindSlot(xb_filter* f, EkfMutex* m, int slot) {
    dev_ = DevRef{kSlot, f, m, slot, -1, kInvalid};
    xb_ekf_slot_info(f, slot, &dev_.time, &dev_.serial);
    cov_valid_ = false;
  }
  void bindWork(xb_filter* f) { dev_ = DevRef{kWork, f, nullptr, -1, -1, time_}; cov_valid_ = false; }
  /** A State of the given dimensions whose estimates and covariance are still on the device (no N x N host matrix). */
  static State shell(int n_poses, int n_features) {
    State s;
    s.p_array_ = Matrix::Zero(n_poses * 3, 1);
    s.q_array_ = Matrix::Zero(n_poses * 4, 1);
    s.f_array_ = Matrix::Zero(n_features * 3, 1);
    s.cov_valid_ = false;
    return s;
  }

  double time_{kInvalid};
  unsigned int seq_{0};
  Vector3 p_{Vector3::Zero()}, v_{Vector3::Zero()};
  Quaternion q_{Quaternion(1.0, 0.0, 0.0, 0.0)};
  Vector3 b_w_{Vector3::Zero()}, b_a_{Vector3::Zero()};
  Matrix p_array_, q_array_, f_array_;
  mutable Matrix cov_;
  Quaternion q_ic_{Quaternion(1.0, 0.0, 0.0, 0.0)};
  Vector3 p_ic_{Vector3::Zero()}, w_m_{Vector3::Zero()}, a_m_{Vector3::Zero()};
  mutable bool cov_valid_ = true;   // cov_ holds this State's covariance
  bool est_dirty_ = false;          // host estimates changed since they were last in sync with the device
  DevRef dev_;
};

// ---- front end, opaque (include/x/vision/tracker.h, include/x/vio/track_manager.h) --------------------------------
class Camera;
class Tracker {
 public:
  Tracker() = default;
  /** tracker.h:39-62: camera + FAST / KLT / RANSAC parameters of the image front end; accepted and ignored (matches are
   *  delivered by the caller). */
  template <typename... FrontEndParams>
  explicit Tracker(const Camera&, FrontEndParams&&...) {}
  /** tracker.cpp: draws the matches into the GUI image; there are no pixels here. */
  static void plotMatches(MatchList&, TiledImage&) {}
};
/** Carries what VioUpdater::preProcess reads from the reference's TrackManager (vio_updater.cpp:172-179): the five
 *  normalised track lists and the lost SLAM feature indexes.  Sorting matches into these lists
 *  (TrackManager::manageTracks) is the step before the hot path. */
/** x::Camera (include/x/vision/camera.h): the intrinsics the track manager needs (FOV distortion model). */
class Camera {
 public:
  Camera() = default;
  Camera(double fx, double fy, double cx, double cy, double s, unsigned int img_width, unsigned int img_height)
      : fx_(fx), fy_(fy), cx_(cx), cy_(cy), s_(s), img_width_(img_width), img_height_(img_height) {}
  [[nodiscard]] unsigned int getWidth() const { return img_width_; }
  [[nodiscard]] unsigned int getHeight() const { return img_height_; }
  /** camera.cpp:69-87: FOV model; the undistorted pixel coordinates go into the feature's x / y. */
  void undistort(Feature& feature) const {
    const double fx = img_width_ * fx_, fy = img_height_ * fy_, cx = img_width_ * cx_, cy = img_height_ * cy_;
    const double inv_fx = 1.0 / fx, inv_fy = 1.0 / fy, cx_n = cx * inv_fx, cy_n = cy * inv_fy;
    const double cam_dist_x = feature.getXDist() * inv_fx - cx_n, cam_dist_y = feature.getYDist() * inv_fy - cy_n;
    const double dist_r = std::sqrt(cam_dist_x * cam_dist_x + cam_dist_y * cam_dist_y);
    double distortion_factor = 1.0;
    if (dist_r > 0.01) distortion_factor = inverseTf(dist_r) / dist_r;
    feature.setX(distortion_factor * cam_dist_x * fx + cx);
    feature.setY(distortion_factor * cam_dist_y * fy + cy);
  }
  /** camera.cpp:122-135: pixel coordinates (x, y) -> normalised image coordinates. */
  [[nodiscard]] Feature normalize(const Feature& feature) const {
    const double fx = img_width_ * fx_, fy = img_height_ * fy_, cx = img_width_ * cx_, cy = img_height_ * cy_;
    const double inv_fx = 1.0 / fx, inv_fy = 1.0 / fy;
    return Feature(feature.getTimestamp(), feature.getX() * inv_fx - cx * inv_fx, feature.getY() * inv_fy - cy * inv_fy,
                   feature.getIntensity());
  }
  /** camera.cpp:163-169 */
  [[nodiscard]] double inverseTf(const double& dist) const {
    return s_ == 0.0 ? dist : std::tan(dist * s_) * (1.0 / (2.0 * std::tan(s_ / 2.0)));
  }
  /** Camera::undistort followed by Camera::normalize (camera.cpp:69-87, 122-135) of a measured image point, through the
   *  camera model of the library (xb_tm_normalize_point).  Returns a Feature with the normalised coordinates. */
  [[nodiscard]] Feature undistortAndNormalize(const Feature& feature) const {
    xb_tm_config c{};
    c.fx = fx_; c.fy = fy_; c.cx = cx_; c.cy = cy_; c.s = s_;
    c.img_width = img_width_; c.img_height = img_height_;
    c.n_tiles_h = c.n_tiles_w = 1;
    xb_track_manager* tm = xb_tm_create(&c);
    if (!tm) throw std::invalid_argument("Camera: invalid intrinsics");
    double out[2] = {0.0, 0.0};
    xb_tm_normalize_point(tm, feature.getXDist(), feature.getYDist(), out);
    xb_tm_destroy(tm);
    return Feature(feature.getTimestamp(), out[0], out[1], feature.getIntensity());
  }
  double fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0, s_ = 0;  // as given: fractions of the image size (camera.cpp:27-35)
  unsigned int img_width_ = 0, img_height_ = 0;
};
/** x::TiledImage (include/x/vision/tiled_image.h) without the pixels: the tile grid manageTracks balances over. */
class TiledImage {
 public:
  TiledImage() = default;
  TiledImage(unsigned int n_tiles_h, unsigned int n_tiles_w) : n_tiles_h_(n_tiles_h), n_tiles_w_(n_tiles_w) {}
  [[nodiscard]] unsigned int getNTilesH() const { return n_tiles_h_; }
  [[nodiscard]] unsigned int getNTilesW() const { return n_tiles_w_; }
 private:
  unsigned int n_tiles_h_ = 1, n_tiles_w_ = 1;
};

inline VioMeasurement::VioMeasurement(const double& timestamp, const unsigned int seq, MatchList matches, const TiledImage& image,
                                      RangeMeasurement range, const SunAngleMeasurement& sun_angle)
    : timestamp{timestamp}, seq{seq}, matches{std::move(matches)}, n_tiles_h{image.getNTilesH()}, n_tiles_w{image.getNTilesW()},
      from_front_end{true}, range{std::move(range)}, sun_angle{sun_angle} {}

/** x::TrackManager (include/x/vio/track_manager.h).  manageTracks (track_manager.cpp:115-436) runs in libxb200.so
 *  (xb_tm_*, host code); the lists can also be injected directly (setTracks) when the caller has its own front end. */
class TrackManager {
 public:
  TrackManager() = default;
  TrackManager(const Camera& camera, const double min_baseline_x_n, const double min_baseline_y_n)
      : camera_(camera), min_baseline_x_n_(min_baseline_x_n), min_baseline_y_n_(min_baseline_y_n) {}
  TrackManager(const TrackManager& o) { *this = o; }
  TrackManager& operator=(const TrackManager& o) {
    if (this == &o) return *this;
    release();
    camera_ = o.camera_; min_baseline_x_n_ = o.min_baseline_x_n_; min_baseline_y_n_ = o.min_baseline_y_n_;
    slam_ = o.slam_; msckf_ = o.msckf_; short_ = o.short_; new_std_ = o.new_std_; new_msckf_ = o.new_msckf_; lost_ = o.lost_;
    injected_ = o.injected_; tiles_h_ = o.tiles_h_; tiles_w_ = o.tiles_w_; facet_ = o.facet_;
    tm_ = o.tm_; refs_ = o.refs_;
    if (refs_) ++*refs_;
    return *this;
  }
  ~TrackManager() { release(); }
  void setCamera(Camera camera) { camera_ = camera; }
  void setTracks(TrackList slam, TrackList msckf, TrackList msckf_short, TrackList new_slam_std, TrackList new_slam_msckf,
                 std::vector<unsigned int> lost_slam_idxs) {
    slam_ = std::move(slam); msckf_ = std::move(msckf); short_ = std::move(msckf_short);
    new_std_ = std::move(new_slam_std); new_msckf_ = std::move(new_slam_msckf); lost_ = std::move(lost_slam_idxs);
    injected_ = true;
  }
  /** track_manager.cpp:115-436; matches carry distorted pixel coordinates (getXDist / getYDist), as after
   *  VIO::importMatches before undistortion -- the undistortion of vio.cpp:399-405 is part of the call. */
  void manageTracks(MatchList& matches, const AttitudeList cam_rots, const size_t n_poses_max, const size_t n_slam_features_max,
                    const size_t min_track_length, TiledImage& img) {
    if (!tm_ || tiles_h_ != img.getNTilesH() || tiles_w_ != img.getNTilesW()) {
      release();
      xb_tm_config c{};
      c.fx = camera_.fx_; c.fy = camera_.fy_; c.cx = camera_.cx_; c.cy = camera_.cy_; c.s = camera_.s_;
      c.img_width = camera_.img_width_; c.img_height = camera_.img_height_;
      c.min_baseline_x_n = min_baseline_x_n_; c.min_baseline_y_n = min_baseline_y_n_;
      c.n_tiles_h = tiles_h_ = img.getNTilesH(); c.n_tiles_w = tiles_w_ = img.getNTilesW();
      tm_ = xb_tm_create(&c);
      if (!tm_) throw std::invalid_argument("TrackManager: invalid camera / tile configuration");
      refs_ = new int(1);
    }
    std::vector<double> mv(10 * matches.size(), 0.0), rots(4 * cam_rots.size());
    for (size_t i = 0; i < matches.size(); ++i) {
      mv[10 * i + 1] = matches[i].previous.getTimestamp(); mv[10 * i + 2] = matches[i].previous.getXDist();
      mv[10 * i + 3] = matches[i].previous.getYDist(); mv[10 * i + 4] = matches[i].current.getTimestamp();
      mv[10 * i + 5] = matches[i].current.getXDist(); mv[10 * i + 6] = matches[i].current.getYDist();
    }
    for (size_t i = 0; i < cam_rots.size(); ++i) {
      rots[4 * i] = cam_rots[i].ax; rots[4 * i + 1] = cam_rots[i].ay; rots[4 * i + 2] = cam_rots[i].az; rots[4 * i + 3] = cam_rots[i].aw;
    }
    detail_check(xb_tm_manage_tracks(tm_, mv.data(), static_cast<int>(matches.size()), rots.data(), static_cast<int>(cam_rots.size()),
                                     static_cast<int>(n_poses_max), static_cast<int>(n_slam_features_max),
                                     static_cast<int>(min_track_length)));
    matches.clear();  // the reference consumes the matched entries (track_manager.cpp:167)
    injected_ = false;
  }
  TrackList normalizeSlamTracks(const int size_out) const { return injected_ || !tm_ ? slam_ : fetch(XB_TM_SLAM, size_out); }
  TrackList getMsckfTracks() const { return injected_ || !tm_ ? msckf_ : fetch(XB_TM_MSCKF, 0); }
  TrackList getShortMsckfTracks() const { return injected_ || !tm_ ? short_ : fetch(XB_TM_MSCKF_SHORT, 0); }
  TrackList getNewSlamStdTracks() const { return injected_ || !tm_ ? new_std_ : fetch(XB_TM_NEW_SLAM_STD, 0); }
  TrackList getNewSlamMsckfTracks() const { return injected_ || !tm_ ? new_msckf_ : fetch(XB_TM_NEW_SLAM_MSCKF, 0); }
  TrackList getOppTracks() const { return tm_ ? fetch(XB_TM_OPP, 0) : TrackList(); }
  /** track_manager.cpp:443-560: the Delaunay facet of SLAM features around the LRF image point (distorted pixels).  With
   *  injected track lists (setTracks) the facet is the one given to setFacet(). */
  std::vector<int> featureTriangleAtPoint(const Feature& lrf_img_pt, TiledImage&) const {
    if (injected_ || !tm_) return facet_;
    int ids[3] = {0, 0, 0};
    const int n = xb_tm_feature_triangle_at_point(tm_, lrf_img_pt.getXDist(), lrf_img_pt.getYDist(), ids);
    detail_check(n);
    return n == 3 ? std::vector<int>(ids, ids + 3) : std::vector<int>();
  }
  void setFacet(std::vector<int> ids) { facet_ = std::move(ids); }
  std::vector<unsigned int> getLostSlamTrackIndexes() const {
    if (injected_ || !tm_) return lost_;
    std::vector<int> tmp(static_cast<size_t>(std::max(1, xb_tm_lost_slam_idxs(tm_, nullptr, 0))));
    const int n = xb_tm_lost_slam_idxs(tm_, tmp.data(), static_cast<int>(tmp.size()));
    return std::vector<unsigned int>(tmp.begin(), tmp.begin() + n);
  }
  void removePersistentTracksAtIndex(const unsigned int idx) { if (tm_) detail_check(xb_tm_remove_persistent_track(tm_, idx)); }
  void removeNewPersistentTracksAtIndexes(const std::vector<unsigned int> invalid_tracks_idx) {
    if (tm_) detail_check(xb_tm_remove_new_persistent_tracks(tm_, invalid_tracks_idx.data(), static_cast<int>(invalid_tracks_idx.size())));
  }
  void clear() {
    slam_.clear(); msckf_.clear(); short_.clear(); new_std_.clear(); new_msckf_.clear(); lost_.clear();
    if (tm_) xb_tm_clear(tm_);
  }

 private:
  static void detail_check(int rc) { if (rc < 0) throw std::invalid_argument("TrackManager: xb_tm call failed"); }
  TrackList fetch(int which, int size_out) const {
    int nt = 0, no = 0;
    detail_check(xb_tm_list_size(tm_, which, size_out, &nt, &no));
    std::vector<int> off(static_cast<size_t>(nt) + 1);
    std::vector<double> xy(2 * static_cast<size_t>(std::max(1, no)));
    std::vector<unsigned long long> ids(static_cast<size_t>(std::max(1, nt)));
    detail_check(xb_tm_get_list(tm_, which, size_out, off.data(), xy.data(), ids.data()));
    TrackList out;
    for (int i = 0; i < nt; ++i) {
      Track t(static_cast<size_t>(off[i + 1] - off[i]), Feature(), ids[static_cast<size_t>(i)]);
      for (int j = off[i]; j < off[i + 1]; ++j) { t[static_cast<size_t>(j - off[i])].setX(xy[2 * j]); t[static_cast<size_t>(j - off[i])].setY(xy[2 * j + 1]); }
      out.push_back(t);
    }
    return out;
  }
  void release() {
    if (refs_ && --*refs_ == 0) { xb_tm_destroy(tm_); delete refs_; }
    tm_ = nullptr; refs_ = nullptr;
  }
  Camera camera_;
  double min_baseline_x_n_ = 0.0, min_baseline_y_n_ = 0.0;
  TrackList slam_, msckf_, short_, new_std_, new_msckf_;
  std::vector<unsigned int> lost_;
  bool injected_ = false;
  std::vector<int> facet_;
  xb_track_manager* tm_ = nullptr;   // shared between copies (VioUpdater keeps a copy of the manager it is given)
  int* refs_ = nullptr;
  unsigned int tiles_h_ = 0, tiles_w_ = 0;
};

/** x::StateManager (include/x/vio/state_manager.h): the window / feature bookkeeping lives in the device filter; this
 *  object is the handle the reference API passes around and reads it back. */
class StateManager {
 public:
  StateManager(int n_poses_max = 0, int n_features_max = 0) : n_poses_max_(n_poses_max), n_features_max_(n_features_max) {}
  void attach(xb_filter* f) { f_ = f; }
  size_t getNFeatures() const { return f_ ? static_cast<size_t>(xb_sm_n_features(f_)) : 0; }
  size_t poseSize() const { return f_ ? static_cast<size_t>(xb_sm_n_poses(f_)) : 0; }
  std::vector<int> getAnchorIdxs() const {
    std::vector<int> a(static_cast<size_t>(std::max(1, n_features_max_)), -1);
    if (f_) xb_sm_anchor_idxs(f_, a.data());
    a.resize(static_cast<size_t>(n_features_max_));
    return a;
  }
  void clear() {
    std::vector<int> a(static_cast<size_t>(std::max(1, n_features_max_)), -1);
    if (f_) xb_sm_set(f_, 0, 0, a.data(), 0);
  }
  /** state_manager.cpp:538-565: the last min(max_size, n_poses) camera attitudes of the window (all of them for
   *  max_size <= 0); `state` must carry its estimates on the host. */
  [[nodiscard]] AttitudeList convertCameraAttitudesToList(const State& state, const int max_size = 0) const {
    const int n_poses = static_cast<int>(poseSize());
    const int size_out = max_size > 0 ? std::min(max_size, n_poses) : n_poses;
    const Matrix orientation_array = state.getOrientationArray();
    AttitudeList attitude_list(static_cast<size_t>(size_out), Attitude());
    const int start_idx = n_poses - size_out;
    for (int i = start_idx; i < n_poses; i++)
      attitude_list[static_cast<size_t>(i - start_idx)] = Attitude(orientation_array(4 * i, 0), orientation_array(4 * i + 1, 0),
                                                                   orientation_array(4 * i + 2, 0), orientation_array(4 * i + 3, 0));
    return attitude_list;
  }
  /** state_manager.cpp:232-271: inverse-depth SLAM features of `state` in world coordinates. */
  [[nodiscard]] std::vector<Eigen::Vector3d> computeSLAMCartesianFeaturesForState(const State& state) const {
    const std::vector<int> anchor_idxs = getAnchorIdxs();
    const size_t n_features = getNFeatures();
    const Matrix feats = state.getFeatureArray(), poss = state.getPositionArray(), atts = state.getOrientationArray();
    std::vector<Eigen::Vector3d> features_xyz(n_features);
    for (size_t i = 0; i < n_features; ++i) {
      const double alpha = feats(3 * i, 0), beta = feats(3 * i + 1, 0), rho = feats(3 * i + 2, 0);
      const int a = anchor_idxs[i];
      const Quaternion q_a(atts(4 * a + 3, 0), atts(4 * a, 0), atts(4 * a + 1, 0), atts(4 * a + 2, 0));
      const Vector3 p_a(poss(3 * a, 0), poss(3 * a + 1, 0), poss(3 * a + 2, 0));
      features_xyz[i] = p_a + 1.0 / rho * (q_a.normalized().toRotationMatrix() * Vector3(alpha, beta, 1.0));
    }
    return features_xyz;
  }
  /** state_manager.cpp:31-149 on a device-bound state. */
  void manage(State& state, std::vector<unsigned int> del_feat_idx) {
    if (!f_ || !state.boundToWork(f_)) throw std::logic_error("StateManager::manage needs a State bound to the device work state");
    std::vector<int> lost(del_feat_idx.begin(), del_feat_idx.end());
    xb_throw(xb_sm_manage(f_, lost.data(), static_cast<int>(lost.size())));
  }

 private:
  int n_poses_max_, n_features_max_;
  xb_filter* f_ = nullptr;
};

// ---- x::Updater (include/x/ekf/updater.h:37-232) ---------------------------------------------------------------------
class Updater {
 public:
  virtual ~Updater() = default;
  virtual double getTime() const = 0;

  /** The reference's template method (updater.cpp:39-115), single-UAV and -DMULTI_UAV flavour. */
  void update(State& state) {
    Matrix h, res, r;
    Matrix correction = Matrix::Zero(state.nErrorStates(), 1);
    bindForUpdate(state);
    preProcess(state);
    const bool short_update_requested = preUpdateShortMsckf();
    if (short_update_requested) {
#ifdef MULTI_UAV
      std::vector<std::shared_ptr<Matrix>> S_list, P_list, H_list, res_list;
      constructShortMsckfUpdate(state, h, res, r, S_list, P_list, H_list, res_list);
      for (size_t j = 0; j < P_list.size(); j++) applyCI(state, *P_list[j], *H_list[j], *res_list[j], *S_list[j]);
#else
      constructShortMsckfUpdate(state, h, res, r);
      if (h.size() > 0) applyUpdate(state, h, res, r, correction, true);
#endif
    }
    const bool update_requested = preUpdate(state);
    if (update_requested) {
      correction = Matrix::Zero(state.nErrorStates(), 1);
#ifdef MULTI_UAV
      std::vector<std::shared_ptr<Matrix>> S_list, P_list, H_list, res_list;
      constructUpdate(state, h, res, r, S_list, P_list, H_list, res_list);
      for (size_t j = 0; j < P_list.size(); j++) applyCI(state, *P_list[j], *H_list[j], *res_list[j], *S_list[j]);
      if (h.size() > 0) applyUpdate(state, h, res, r, correction, true);
#else
      for (int i = 0; i < iekf_iter_; i++) {
        const bool is_last_iter = i == iekf_iter_ - 1;
        constructUpdate(state, h, res, r);
        if (h.size() > 0) applyUpdate(state, h, res, r, correction, is_last_iter);
      }
#endif
      postUpdate(state, correction);
    }
    finishUpdate(state);
  }
#ifdef MULTI_UAV
  /** updater.cpp:22-36 */
  void collaborativeUpdate(State& state) {
    bindForUpdate(state);
    if (preUpdateCI()) {
      std::vector<std::shared_ptr<Matrix>> S_list, P_list, H_list, res_list;
      constructSlamCIUpdate(state, S_list, P_list, H_list, res_list);
      for (size_t i = 0; i < P_list.size(); i++) applyCI(state, *P_list[i], *H_list[i], *res_list[i], *S_list[i]);
    }
    finishUpdate(state);
  }
#endif

 protected:
  int iekf_iter_{1};

  /** A 1 x 1 matrix holding this value stands for "the measurement constructed on the device" (see the header note). */
  static Matrix deviceToken() { return Matrix::Constant(1, 1, kDeviceToken()); }
  static bool isDeviceToken(const Matrix& h) { return h.rows() == 1 && h.cols() == 1 && h(0, 0) == kDeviceToken(); }

  /** updater.cpp:117-141.  H m x N, res m x 1, R m x m (its diagonal is used: every R the reference builds is diagonal). */
  void applyUpdate(State& state, const Eigen::MatrixXd& H, const Eigen::MatrixXd& res, const Eigen::MatrixXd& R,
                   Matrix& correction_total, bool cov_update = true) {
    xb_filter* f = deviceOf(state);
    if (isDeviceToken(H)) {
      xb_throw(xb_updater_apply_constructed(f, cov_update ? 1 : 0));
      return;
    }
    const int m = static_cast<int>(H.rows()), n = static_cast<int>(H.cols());
    std::vector<double> h(static_cast<size_t>(m) * n), rr(m), rd(m), ct(n);
    for (int i = 0; i < m; ++i) {
      rr[i] = res(i, 0);
      rd[i] = R(i, i);
      for (int j = 0; j < n; ++j) h[static_cast<size_t>(i) * n + j] = H(i, j);
    }
    for (int j = 0; j < n; ++j) ct[j] = correction_total(j, 0);
    xb_throw(xb_updater_apply_update(f, h.data(), rr.data(), rd.data(), m, ct.data(), cov_update ? 1 : 0));
    for (int j = 0; j < n; ++j) correction_total(j, 0) = ct[j];
  }
#ifdef MULTI_UAV
  /** updater.cpp:144-161 with caller-supplied host matrices. */
  void applyCI(State& state, Matrix& ci_P, const Matrix& H, const Matrix& res, Matrix& S) {
    xb_filter* f = deviceOf(state);
    xb_throw(xb_work_set(f, nullptr, ci_P.data(), XB_COL_MAJOR));
    const int m = static_cast<int>(H.rows()), n = static_cast<int>(H.cols());
    std::vector<double> h(static_cast<size_t>(m) * n), rr(m), ss(static_cast<size_t>(m) * m);
    for (int i = 0; i < m; ++i) {
      rr[i] = res(i, 0);
      for (int j = 0; j < n; ++j) h[static_cast<size_t>(i) * n + j] = H(i, j);
      for (int j = 0; j < m; ++j) ss[static_cast<size_t>(i) * m + j] = S(i, j);
    }
    xb_throw(xb_updater_apply_ci(f, h.data(), rr.data(), ss.data(), m, nullptr, 0, 1.0));
  }
#endif

  virtual void preProcess(const State& state) = 0;
  virtual bool preUpdate(State& state) = 0;
  virtual bool preUpdateShortMsckf() = 0;
#ifdef MULTI_UAV
  virtual bool preUpdateCI() = 0;
  virtual void constructSlamCIUpdate(const State& state, std::vector<std::shared_ptr<Matrix>>& S_list,
                                     std::vector<std::shared_ptr<Matrix>>& P_list, std::vector<std::shared_ptr<Matrix>>& H_list,
                                     std::vector<std::shared_ptr<Matrix>>& res_list) = 0;
  virtual void constructUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r,
                               std::vector<std::shared_ptr<Matrix>>& S_list, std::vector<std::shared_ptr<Matrix>>& P_list,
                               std::vector<std::shared_ptr<Matrix>>& H_list, std::vector<std::shared_ptr<Matrix>>& res_list) = 0;
  virtual void constructShortMsckfUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r,
                                         std::vector<std::shared_ptr<Matrix>>& S_list,
                                         std::vector<std::shared_ptr<Matrix>>& P_list,
                                         std::vector<std::shared_ptr<Matrix>>& H_list,
                                         std::vector<std::shared_ptr<Matrix>>& res_list) = 0;
#else
  virtual void constructUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r) = 0;
  virtual void constructShortMsckfUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r) = 0;
#endif
  virtual void postUpdate(State& state, const Matrix& correction) = 0;

  /** The device filter this updater works on (set by Ekf::set; the Ekf owns it). */
  xb_filter* device_ = nullptr;
  xb_filter* deviceOf(const State& s) const {
    if (s.dev_.kind == State::kWork && s.dev_.f) return s.dev_.f;
    if (device_) return device_;
    throw std::logic_error("x::Updater: no device filter (construct an x::Ekf with this updater and call Ekf::set first)");
  }

 private:
  friend class Ekf;
  static double kDeviceToken() { return -7.2057594037927936e16; }
  /** Make `state` the device's work state: a State that is not already bound to it is uploaded. */
  void bindForUpdate(State& state) {
    xb_filter* f = deviceOf(state);
    if (state.boundToWork(f) && !state.est_dirty_) return;
    const std::vector<double> x = state.xvec();
    if (state.boundToWork(f)) {
      xb_throw(xb_work_set(f, x.data(), nullptr, XB_COL_MAJOR));
    } else {
      state.fetchCov();
      xb_throw(xb_work_set(f, x.data(), state.cov_.data(), XB_COL_MAJOR));
      state.bindWork(f);
    }
    state.est_dirty_ = false;
  }
  /** Mirror the device's work estimates back into the host object at the end of an update. */
  void finishUpdate(State& state) {
    xb_filter* f = deviceOf(state);
    std::vector<double> x(XB_XVEC_LEN(state.nPosesMax(), state.nFeaturesMax()));
    xb_throw(xb_work_get(f, x.data(), nullptr, XB_COL_MAJOR));
    state.setFromXvec(x.data(), state.nPosesMax(), state.nFeaturesMax());
    state.cov_valid_ = false;
  }
};

// ---- x::VioUpdater (include/x/vio/vio_updater.h:35-335), back-end part ------------------------------------------------
class VioUpdater : public Updater {
 public:
  VioUpdater() = default;
  VioUpdater(Tracker& tracker, StateManager& state_manager, TrackManager& track_manager, double sigma_img, double sigma_range,
             double rho_0, double sigma_rho_0, int min_track_length, double sigma_landmark = 0, double ci_msckf_w = -1.0,
             double ci_slam_w = -1.0, int iekf_iter = 1)
      : tracker_(tracker), state_manager_(state_manager), track_manager_(track_manager), sigma_img_{sigma_img},
        sigma_landmark_{sigma_landmark}, sigma_range_{sigma_range}, rho_0_{rho_0}, sigma_rho_0_{sigma_rho_0},
        min_track_length_{min_track_length}, ci_msckf_w_{ci_msckf_w}, ci_slam_w_(ci_slam_w) {
    iekf_iter_ = iekf_iter;
  }
  void setMeasurement(const VioMeasurement& measurement) { measurement_ = measurement; }   // vio_updater.cpp:122-124
  [[nodiscard]] double getTime() const override { return measurement_.timestamp; }
  /** The front-end seam (what the reference's preProcess pulls out of its TrackManager / Tracker copies). */
  TrackManager& trackManager() { return track_manager_; }
  StateManager& stateManager() { return state_manager_; }
  [[nodiscard]] TiledImage& getFeatureImage() { return feature_img_; }   // vio_updater.h:62
  /** True when preProcess needs the estimates of the update state on the host (a measurement that carries matches). */
  [[nodiscard]] bool needsHostState() const { return measurement_.from_front_end; }
#ifdef MULTI_UAV
  void getMsckfTracks(TrackList& tracks) { tracks = track_manager_.getMsckfTracks(); }
  void getOppTracks(TrackList& tracks) { tracks = track_manager_.getOppTracks(); }   // vio_updater.h:75
  void getSlamTracks(TrackList& tracks, std::vector<int>& anchor_idxs, const int n_poses_max) {
    tracks = track_manager_.normalizeSlamTracks(n_poses_max);
    anchor_idxs = state_manager_.getAnchorIdxs();
  }
  /** Tracker::getMsckfMatches / getSlamMatches of the reference (vio_updater.cpp:185,212). */
  void setMsckfMatches(const MsckfMatches& m) { msckf_matches_ = m; }
  void setSlamMatches(const SlamMatches& m) { slam_matches_ = m; }
#endif
  void fillConfig(xb_config& c) const {
    c.sigma_img = sigma_img_; c.sigma_range = sigma_range_; c.rho_0 = rho_0_; c.sigma_rho_0 = sigma_rho_0_;
    c.min_track_length = min_track_length_; c.sigma_landmark = sigma_landmark_; c.ci_msckf_w = ci_msckf_w_;
    c.ci_slam_w = ci_slam_w_; c.iekf_iter = iekf_iter_;
  }

 private:
  VioMeasurement measurement_;
  Tracker tracker_;
  StateManager state_manager_;
  TrackManager track_manager_;
  double sigma_img_{}, sigma_landmark_{}, sigma_range_{}, rho_0_{}, sigma_rho_0_{};
  int min_track_length_{};
  double ci_msckf_w_{}, ci_slam_w_{};
  TrackList msckf_trks_, msckf_short_trks_, new_slam_std_trks_, new_msckf_slam_trks_, slam_trks_;
  std::vector<unsigned int> lost_slam_trk_idxs_;
  MsckfMatches msckf_matches_;
  SlamMatches slam_matches_;
  TiledImage feature_img_;

  struct Csr { std::vector<int> off; std::vector<double> obs; };
  static void pack(const TrackList& tl, Csr& c, xb_track_list& out) {
    c.off.assign(1, 0);
    for (const auto& t : tl) {
      for (const auto& o : t) { c.obs.push_back(o.getX()); c.obs.push_back(o.getY()); }
      c.off.push_back(static_cast<int>(c.obs.size() / 2));
    }
    out.n_tracks = static_cast<int>(tl.size());
    out.off = c.off.data();
    out.obs = c.obs.data();
  }

  /** vio_updater.cpp:126-198 from the track-list seam on: the lists go to the device (the only host-to-device traffic
   *  of an update), MSCKF-MSCKF matches are resolved to (list, index) by Track id as msckf_update.cpp:96-98 does. */
  void preProcess(const State& state) override {
    xb_filter* f = deviceOf(state);
    const int n_poses_max = state.nPosesMax();
    if (measurement_.from_front_end) {
      // vio_updater.cpp:142-170: camera attitudes of the window without its oldest pose, plus the current one (the pose
      // window has not been slid yet), then the matches are sorted into tracks.  Ekf::processUpdateMeasurement hands a
      // state with its estimates on the host for such a measurement.
      AttitudeList cam_rots = state_manager_.convertCameraAttitudesToList(state, n_poses_max - 1);
      cam_rots.push_back(state.computeCameraAttitude());
      feature_img_ = TiledImage(measurement_.n_tiles_h, measurement_.n_tiles_w);
      track_manager_.manageTracks(measurement_.matches, cam_rots, static_cast<size_t>(n_poses_max),
                                  static_cast<size_t>(state.nFeaturesMax()), static_cast<size_t>(min_track_length_), feature_img_);
    }
    slam_trks_ = track_manager_.normalizeSlamTracks(n_poses_max);
    msckf_trks_ = track_manager_.getMsckfTracks();
    msckf_short_trks_ = track_manager_.getShortMsckfTracks();
    new_slam_std_trks_ = track_manager_.getNewSlamStdTracks();
    new_msckf_slam_trks_ = track_manager_.getNewSlamMsckfTracks();
    lost_slam_trk_idxs_ = track_manager_.getLostSlamTrackIndexes();
    Csr a, b, c, d, e;
    xb_measurement m{};
    m.timestamp = measurement_.timestamp;
    pack(slam_trks_, a, m.slam);
    pack(msckf_trks_, b, m.msckf);
    pack(msckf_short_trks_, c, m.msckf_short);
    pack(new_slam_std_trks_, d, m.new_slam_std);
    pack(new_msckf_slam_trks_, e, m.new_msckf_slam);
    std::vector<int> lost(lost_slam_trk_idxs_.begin(), lost_slam_trk_idxs_.end());
    m.n_lost = static_cast<int>(lost.size());
    m.lost_slam_idxs = lost.data();
    xb_throw(xb_vio_set_measurement(f, &m));
    {  // range / sun-sensor measurements (vio_updater.cpp:352-403; the facet lookup uses the hard-coded image point of :360-363)
      xb_range_measurement rm{};
      xb_sun_angle_measurement sm{};
      const bool with_range = measurement_.range.timestamp > 0.1 && !slam_trks_.empty();
      if (with_range) {
        Feature lrf_img_pt;
        lrf_img_pt.setXDist(320.5);
        lrf_img_pt.setYDist(240.5);
        TiledImage img;
        const std::vector<int> ids = track_manager_.featureTriangleAtPoint(lrf_img_pt, img);
        rm.timestamp = measurement_.range.timestamp;
        rm.range = measurement_.range.range;
        rm.img_pt_n[0] = measurement_.range.img_pt_n.getX();
        rm.img_pt_n[1] = measurement_.range.img_pt_n.getY();
        rm.n_tr_feat_ids = static_cast<int>(ids.size());
        for (size_t j = 0; j < ids.size() && j < 3; ++j) rm.tr_feat_ids[j] = ids[j];
      }
      sm.timestamp = measurement_.sun_angle.timestamp;
      sm.x_angle = measurement_.sun_angle.x_angle;
      sm.y_angle = measurement_.sun_angle.y_angle;
      xb_throw(xb_vio_set_sensors(f, with_range ? &rm : nullptr, &sm));
    }
    xb_throw(xb_updater_reset_correction(f));
#ifdef MULTI_UAV
    std::vector<const SimpleState*> uniq;
    std::vector<xb_msckf_match> cm;
    std::vector<std::vector<double>> obs;
    for (const auto& mm : msckf_matches_) {
      int which = -1, idx = -1;
      for (size_t j = 0; j < msckf_trks_.size() && idx < 0; ++j)
        if (msckf_trks_[j].getId() == mm.id_current_track) { which = 0; idx = static_cast<int>(j); }
      for (size_t j = 0; j < msckf_short_trks_.size() && idx < 0; ++j)
        if (msckf_short_trks_[j].getId() == mm.id_current_track) { which = 1; idx = static_cast<int>(j); }
      if (idx < 0 || !mm.received_track_ptr) continue;   // no own track with that id: the reference's loop never consumes it
      size_t k = 0;
      while (k < uniq.size() && uniq[k] != mm.state.get()) ++k;
      if (k == uniq.size()) uniq.push_back(mm.state.get());
      obs.emplace_back();
      for (const auto& o : *mm.received_track_ptr) { obs.back().push_back(o.getX()); obs.back().push_back(o.getY()); }
      cm.push_back({static_cast<int>(k), which, idx, static_cast<int>(mm.received_track_ptr->size()), nullptr});
    }
    for (size_t j = 0; j < cm.size(); ++j) cm[j].obs = obs[j].data();
    std::vector<xb_peer_state> ps;
    for (const auto* u : uniq) ps.push_back(u->view());
    xb_throw(xb_vio_set_msckf_matches(f, ps.data(), static_cast<int>(ps.size()), cm.data(), static_cast<int>(cm.size())));
    msckf_matches_.clear();   // preProcess replaces the list on every update (vio_updater.cpp:185)
#endif
  }
  /** vio_updater.cpp:200-207 */
  [[nodiscard]] bool preUpdate(State& state) override {
    xb_filter* f = deviceOf(state);
    std::vector<int> lost(lost_slam_trk_idxs_.begin(), lost_slam_trk_idxs_.end());
    xb_throw(xb_sm_manage(f, lost.data(), static_cast<int>(lost.size())));
    xb_throw(xb_updater_reset_correction(f));
    first_iter_ = true;
    return !(msckf_trks_.empty() && slam_trks_.empty() && new_slam_std_trks_.empty() && new_msckf_slam_trks_.empty());
  }
  [[nodiscard]] bool preUpdateShortMsckf() override { return !msckf_short_trks_.empty(); }   // vio_updater.cpp:209-215
#ifdef MULTI_UAV
  [[nodiscard]] bool preUpdateCI() override { return !slam_matches_.empty(); }   // vio_updater.cpp:76-79
  /** vio_updater.cpp:81-115: MultiSlamUpdate + pair fuseCI + applyCI over the list, on the device (the lists stay empty). */
  void constructSlamCIUpdate(const State& state, std::vector<std::shared_ptr<Matrix>>&, std::vector<std::shared_ptr<Matrix>>&,
                             std::vector<std::shared_ptr<Matrix>>&, std::vector<std::shared_ptr<Matrix>>&) override {
    xb_filter* f = deviceOf(state);
    std::vector<const SimpleState*> uniq;
    std::vector<xb_slam_match> cm;
    for (const auto& m : slam_matches_) {
      size_t k = 0;
      while (k < uniq.size() && uniq[k] != m.state.get()) ++k;
      if (k == uniq.size()) uniq.push_back(m.state.get());
      cm.push_back({static_cast<int>(k), m.current_feature_id, m.received_feature_id});
    }
    std::vector<xb_peer_state> ps;
    for (const auto* u : uniq) ps.push_back(u->view());
    xb_throw(xb_updater_collaborative_update(f, ps.data(), static_cast<int>(ps.size()), cm.data(), static_cast<int>(cm.size())));
    slam_matches_.clear();   // tracker_.cleanSlamMatches() (vio_updater.cpp:107)
  }
  /** vio_updater.cpp:266-423: the stacked, compressed measurement is built on the device; its CI lists are applied there
   *  too (the reference's loop over P_list finds the host lists empty), then the token makes applyUpdate finish the job. */
  void constructUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r, std::vector<std::shared_ptr<Matrix>>&,
                       std::vector<std::shared_ptr<Matrix>>&, std::vector<std::shared_ptr<Matrix>>&,
                       std::vector<std::shared_ptr<Matrix>>&) override {
    construct(state, 0, h, res, r);
    xb_throw(xb_updater_apply_ci_lists(deviceOf(state)));
  }
  void constructShortMsckfUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r,
                                 std::vector<std::shared_ptr<Matrix>>&, std::vector<std::shared_ptr<Matrix>>&,
                                 std::vector<std::shared_ptr<Matrix>>&, std::vector<std::shared_ptr<Matrix>>&) override {
    construct(state, 1, h, res, r);
    xb_throw(xb_updater_apply_ci_lists(deviceOf(state)));
    h.resize(0, 0);   // this build applies only the CI lists of the short tracks (updater.cpp:58-70)
  }
#else
  void constructUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r) override { construct(state, 0, h, res, r); }
  void constructShortMsckfUpdate(const State& state, Matrix& h, Matrix& res, Matrix& r) override { construct(state, 1, h, res, r); }
#endif
  /** vio_updater.cpp:425-449 */
  void postUpdate(State& state, const Matrix&) override { xb_throw(xb_vio_post_update(deviceOf(state))); }

  void construct(const State& state, int which, Matrix& h, Matrix& res, Matrix& r) {
    xb_filter* f = deviceOf(state);
    xb_throw(xb_vio_construct_update(f, which));
    const bool any = which == 1 ? !msckf_short_trks_.empty()
                                : !(msckf_trks_.empty() && slam_trks_.empty() && new_msckf_slam_trks_.empty());
    if (any) { h = deviceToken(); res = Matrix::Zero(1, 1); r = Matrix::Identity(1, 1); }
    else { h.resize(0, 0); res.resize(0, 0); r.resize(0, 0); }
    first_iter_ = false;
  }
  bool first_iter_ = true;
  friend class Ekf;
  friend class VIO;
};

// ---- x::Ekf (include/x/ekf/ekf.h:53-195) -----------------------------------------------------------------------------
class Ekf {
 public:
  explicit Ekf(Updater& updater) : updater_{updater} {}
  /** ekf.h:63: the copy refers to the same updater and shares the device filter (which holds propagator + ring buffer). */
  Ekf(const Ekf& ekf) : updater_{ekf.updater_}, dev_{ekf.dev_}, M_{ekf.M_}, F_{ekf.F_} {}

  /** ekf.cpp:32-41.  max_tracks / device are this back end's additions (capacity of one update, CUDA ordinal). */
  void set(const Updater& updater, const Vector3& g, const ImuNoise& imu_noise, const int state_buffer_sz,
           const State& default_state, double a_m_max, unsigned int delta_seq_imu, const double& time_margin_bfr,
           int max_tracks = 1024, int device = 0) {
    updater_.iekf_iter_ = updater.iekf_iter_;   // `updater_ = updater` of the reference: a base-class slice (ekf.cpp:36)
    xb_config c;
    xb_default_config(&c);
    c.n_poses_max = default_state.nPosesMax();
    c.n_features_max = default_state.nFeaturesMax();
    c.n_slots = state_buffer_sz;
    c.device = device;
    c.max_tracks = max_tracks;
    for (int i = 0; i < 3; ++i) c.g[i] = g(i);
    c.n_w = imu_noise.n_w; c.n_bw = imu_noise.n_bw; c.n_a = imu_noise.n_a; c.n_ba = imu_noise.n_ba;
    c.a_m_max = a_m_max; c.delta_seq_imu = delta_seq_imu; c.time_margin = time_margin_bfr;
    if (auto* v = dynamic_cast<const VioUpdater*>(&updater)) v->fillConfig(c);
    c.iekf_iter = updater.iekf_iter_;
#ifdef MULTI_UAV
    c.multi_uav = 1;   // the reference selects this flow at compile time (CMakeLists.txt:36,67-71)
#endif
    xb_filter* f = nullptr;
    xb_throw(xb_create(&c, &f));
    dev_ = std::shared_ptr<Dev>(new Dev{f, {}}, [](Dev* d) { xb_destroy(d->f); delete d; });
    M_ = c.n_poses_max;
    F_ = c.n_features_max;
    updater_.device_ = f;
    if (auto* v = dynamic_cast<VioUpdater*>(&updater_)) v->state_manager_.attach(f);
  }
  /** ekf.cpp:43-64 */
  void initializeFromState(const State& init_state) {
    if (!dev_) throw std::runtime_error("The EKF state buffer must have non-zero size.");
    if (init_state.nPosesMax() != M_ || init_state.nFeaturesMax() != F_ || init_state.q_array_.rows() != 4 * M_ ||
        init_state.getCovariance().rows() != XB_NERR(M_, F_) || init_state.getCovariance().cols() != XB_NERR(M_, F_))
      throw init_bfr_mismatch{};
    const std::vector<double> x = init_state.xvec();
    const Matrix cov = init_state.getCovariance();
    std::lock_guard<EkfMutex> lk(dev_->mutex);
    xb_throw(xb_ekf_initialize_from_state(dev_->f, x.data(), cov.data(), XB_COL_MAJOR));
  }
  /** ekf.cpp:66-140 */
  std::optional<State> processImu(const double timestamp, unsigned int seq, const Vector3& w_m, const Vector3& a_m) {
    if (!dev_) return std::nullopt;
    std::lock_guard<EkfMutex> lk(dev_->mutex);
    std::vector<double> x(XB_XVEC_LEN(M_, F_));
    const double w[3] = {w_m(0), w_m(1), w_m(2)}, a[3] = {a_m(0), a_m(1), a_m(2)};
    const int rc = xb_ekf_process_imu(dev_->f, timestamp, seq, w, a, x.data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    State out;
    out.setFromXvec(x.data(), M_, F_);
    out.bindSlot(dev_->f, &dev_->mutex, xb_ekf_newest_slot(dev_->f));
    return out;
  }
  /** Addition to the reference API: a run of processImu calls in one (xb_ekf_process_imu_batch: three launches for up to
   *  32 samples instead of one per sample).  Returns the newest propagated state, nullopt if no sample produced one. */
  std::optional<State> processImuBatch(const std::vector<double>& timestamps, const std::vector<unsigned int>& seqs,
                                       const std::vector<Vector3>& w_m, const std::vector<Vector3>& a_m) {
    if (!dev_ || timestamps.empty()) return std::nullopt;
    if (seqs.size() != timestamps.size() || w_m.size() != timestamps.size() || a_m.size() != timestamps.size())
      throw std::invalid_argument("Ekf::processImuBatch: argument lengths differ");
    std::lock_guard<EkfMutex> lk(dev_->mutex);
    std::vector<double> x(XB_XVEC_LEN(M_, F_)), w(3 * timestamps.size()), a(3 * timestamps.size());
    for (size_t i = 0; i < timestamps.size(); ++i)
      for (int e = 0; e < 3; ++e) { w[3 * i + e] = w_m[i](e); a[3 * i + e] = a_m[i](e); }
    const int rc = xb_ekf_process_imu_batch(dev_->f, static_cast<int>(timestamps.size()), timestamps.data(), seqs.data(),
                                            w.data(), a.data(), x.data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    State out;
    out.setFromXvec(x.data(), M_, F_);
    out.bindSlot(dev_->f, &dev_->mutex, xb_ekf_newest_slot(dev_->f));
    return out;
  }
  /** ekf.cpp:179-213: closestIdx + copy of the buffered state, Updater::update on it, write-back + re-propagation. */
  std::optional<State> processUpdateMeasurement() {
    if (!dev_) return std::nullopt;
    std::lock_guard<EkfMutex> lk(dev_->mutex);
    // the host mirror of the estimates is only filled in for updaters that read it (a device-native VioUpdater does not)
    const auto* vio_updater = dynamic_cast<VioUpdater*>(&updater_);
    const bool native = vio_updater != nullptr && !vio_updater->needsHostState();
    std::vector<double> x0(native ? 0 : XB_XVEC_LEN(M_, F_));
    int rc = xb_ekf_update_begin(dev_->f, updater_.getTime(), native ? nullptr : x0.data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    State update_state = State::shell(M_, F_);
    if (!native) update_state.setFromXvec(x0.data(), M_, F_);
    update_state.bindWork(dev_->f);
    updater_.update(update_state);
    std::vector<double> x(XB_XVEC_LEN(M_, F_));
    rc = xb_ekf_update_end(dev_->f, x.data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    update_state.setFromXvec(x.data(), M_, F_);
    update_state.bindSlot(dev_->f, &dev_->mutex, xb_ekf_last_update_slot(dev_->f));
    return update_state;
  }
#ifdef MULTI_UAV
  /** ekf.cpp:143-176 */
  std::optional<State> processOthersMeasurement(double timestamp) {
    if (!dev_) return std::nullopt;
    std::lock_guard<EkfMutex> lk(dev_->mutex);
    int rc = xb_ekf_update_begin(dev_->f, timestamp, nullptr);
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    State update_state = State::shell(M_, F_);
    update_state.bindWork(dev_->f);
    updater_.collaborativeUpdate(update_state);
    std::vector<double> x(XB_XVEC_LEN(M_, F_));
    rc = xb_ekf_update_end(dev_->f, x.data());
    xb_throw(rc);
    if (rc == 0) return std::nullopt;
    update_state.setFromXvec(x.data(), M_, F_);
    update_state.bindSlot(dev_->f, &dev_->mutex, xb_ekf_last_update_slot(dev_->f));
    return update_state;
  }
#endif
  void lock() { if (dev_) dev_->mutex.lock(); }       // ekf.h:128
  void unlock() { if (dev_) dev_->mutex.unlock(); }   // ekf.h:133
  xb_filter* handle() { return dev_ ? dev_->f : nullptr; }

 private:
  struct Dev { xb_filter* f; EkfMutex mutex; };
  Updater& updater_;
  std::shared_ptr<Dev> dev_;
  int M_ = 0, F_ = 0;
};

}  // namespace x
