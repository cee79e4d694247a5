// Test infrastructure (oracle/): C entry points around the reference's OWN filter back end, compiled where it lies
// (/root/reference/src/x/{ekf,vio,vision}/*.cpp, unmodified) against the stand-in headers of shim/.  This file plays
// the role of x::VIO (src/x/vio/vio.cpp:40,54-111,176-214): it owns Tracker / StateManager / TrackManager /
// VioUpdater / Ekf, builds them exactly as VIO::setUp does, and feeds them through the reference's public API
// (Ekf::initializeFromState / processImu / processUpdateMeasurement / processOthersMeasurement).  Synthetic track
// lists enter at the VioUpdater::preProcess seam through the TrackManager stand-in (shim/x/vio/track_manager.h).
// Used by tests/ (to pin the numpy oracle and to check the CUDA path) and by bench.py's reference arm.  Never part
// of the product path.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <utility>

#include <Eigen/Dense>
#include <Eigen/QR>
#include <boost/log/trivial.hpp>
#include <boost/math/distributions.hpp>
#include <opencv2/core/core.hpp>
#ifdef MULTI_UAV
#include <nlopt.hpp>
#endif

// The harness needs the ring buffer of x::Ekf (newest state / covariance), the IMU sample of x::State and the
// protected stage methods of x::Updater / x::VioUpdater.  Access specifiers do not change the object layout under the
// Itanium ABI, so this translation unit alone sees them as public.
#define private public
#define protected public
#include "x/ekf/ekf.h"
#include "x/vio/msckf_update.h"
#include "x/vio/vio_updater.h"
#include "x/vio/range_update.h"
#include "x/vio/solar_update.h"
#undef private
#undef protected

namespace x {
// VioUpdater declares `friend class VIO` (vio_updater.h:334); the harness takes VIO's place.
class VIO {
 public:
  int M, F;
  Tracker tracker;
  StateManager state_manager;
  TrackManager track_manager;
  VioUpdater updater;
  Ekf ekf;
  std::optional<State> last;

  VIO(const double* c)
      : M(int(c[0])),
        F(int(c[1])),
        state_manager(int(c[0]), int(c[1])),
        updater(tracker, state_manager, track_manager, c[3], c[4], c[5], c[6], int(c[7]), c[8], c[9], c[10],
                int(c[11])),
        ekf(updater) {
    ImuNoise noise;
    noise.n_w = c[15];
    noise.n_bw = c[16];
    noise.n_a = c[17];
    noise.n_ba = c[18];
    // vio.cpp:206-214
    ekf.set(updater, Vector3(c[12], c[13], c[14]), noise, int(c[2]), State(M, F), c[19], 1, c[20]);
  }

  // VIO::initAtTime (vio.cpp:54-111) with the initial state given by the caller
  int init(const State& s) {
    ekf.lock();
    updater.track_manager_.clear();
    updater.state_manager_.clear();
    int rc = 0;
    try {
      ekf.initializeFromState(s);
    } catch (std::runtime_error&) {
      rc = -1;
    } catch (init_bfr_mismatch&) {
      rc = -2;
    }
    ekf.unlock();
    return rc;
  }

  TrackManager::Lists& lists() { return *updater.track_manager_.lists; }
  StateManager& sm() { return updater.state_manager_; }
#ifdef MULTI_UAV
  Tracker::Shared& matches() { return *updater.tracker_.sh; }
#endif
};
}  // namespace x

using namespace x;

namespace {
int xlen(int M, int F) { return 32 + 7 * M + 3 * F; }

State state_from(const double* xv, const double* cov_rm, int M, int F) {
  const int N = 15 + 6 * M + 3 * F;
  Matrix pa(3 * M, 1), qa(4 * M, 1), fa(3 * F, 1), cov(N, N);
  for (int i = 0; i < 3 * M; ++i) pa(i, 0) = xv[32 + i];
  for (int i = 0; i < 4 * M; ++i) qa(i, 0) = xv[32 + 3 * M + i];
  for (int i = 0; i < 3 * F; ++i) fa(i, 0) = xv[32 + 7 * M + i];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) cov(i, j) = cov_rm[size_t(i) * N + j];
  return State(xv[29], static_cast<unsigned>(xv[30]), Vector3(xv[0], xv[1], xv[2]), Vector3(xv[3], xv[4], xv[5]),
               Quaternion(xv[9], xv[6], xv[7], xv[8]), Vector3(xv[10], xv[11], xv[12]),
               Vector3(xv[13], xv[14], xv[15]), pa, qa, fa, cov, Quaternion(xv[19], xv[16], xv[17], xv[18]),
               Vector3(xv[20], xv[21], xv[22]), Vector3(xv[23], xv[24], xv[25]), Vector3(xv[26], xv[27], xv[28]));
}

void state_to(const State& s, double* xv, int M, int F) {
  if (!xv) return;
  std::memset(xv, 0, sizeof(double) * size_t(xlen(M, F)));
  const Vector3 p = s.getPosition();
  const Quaternion q = s.getOrientation(), qic = s.getOrientationExtrinsics();
  const Vector3 pic = s.getPositionExtrinsics();
  const Eigen::VectorXd dyn = s.getDynamicStates();  // p v q(x,y,z,w) b_w b_a
  for (int i = 0; i < 16; ++i) xv[i] = dyn(i);
  (void)p;
  (void)q;
  xv[16] = qic.x(), xv[17] = qic.y(), xv[18] = qic.z(), xv[19] = qic.w();
  for (int i = 0; i < 3; ++i) xv[20 + i] = pic(i);
  for (int i = 0; i < 3; ++i) xv[23 + i] = s.w_m_(i), xv[26 + i] = s.a_m_(i);
  xv[29] = s.getTime();
  xv[30] = double(s.getSeq());
  const Matrix pa = s.getPositionArray(), qa = s.getOrientationArray(), fa = s.getFeatureArray();
  for (int i = 0; i < 3 * M; ++i) xv[32 + i] = pa(i, 0);
  for (int i = 0; i < 4 * M; ++i) xv[32 + 3 * M + i] = qa(i, 0);
  for (int i = 0; i < 3 * F; ++i) xv[32 + 7 * M + i] = fa(i, 0);
}

void cov_to(const Matrix& c, double* out_rm) {
  const Eigen::Index n = c.rows();
  for (Eigen::Index i = 0; i < n; ++i)
    for (Eigen::Index j = 0; j < n; ++j) out_rm[size_t(i) * n + j] = c(i, j);
}

TrackList make_tracks(double ts, int n, const int* off, const double* obs, const unsigned long long* ids) {
  TrackList tl;
  tl.reserve(size_t(n));
  for (int j = 0; j < n; ++j) {
    Track t = ids ? Track(0, Feature(), ids[j]) : Track();
    for (int i = off[j]; i < off[j + 1]; ++i) t.push_back(Feature(ts, obs[2 * i], obs[2 * i + 1], 0.0));
    tl.push_back(t);
  }
  return tl;
}
}  // namespace

extern "C" {

// Bind the BLAS/LAPACK the Eigen stand-in routes its large products / QR / LU to (numpy's ILP64 OpenBLAS).
// threads <= 0 keeps the library default.  Returns 0 on success.
int xref_set_blas(const char* path, int threads) {
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return -1;
  xref::Blas& b = xref::blas();
  b.dgemm = reinterpret_cast<decltype(b.dgemm)>(dlsym(h, "scipy_dgemm_64_"));
  b.dgeqrf = reinterpret_cast<decltype(b.dgeqrf)>(dlsym(h, "scipy_dgeqrf_64_"));
  b.dgetrf = reinterpret_cast<decltype(b.dgetrf)>(dlsym(h, "scipy_dgetrf_64_"));
  b.dgetri = reinterpret_cast<decltype(b.dgetri)>(dlsym(h, "scipy_dgetri_64_"));
  b.set_threads = reinterpret_cast<decltype(b.set_threads)>(dlsym(h, "scipy_openblas_set_num_threads64_"));
  if (!b.dgemm || !b.dgeqrf || !b.dgetrf || !b.dgetri) {
    b = xref::Blas();
    return -2;
  }
  if (threads > 0 && b.set_threads) b.set_threads(threads);
  return 0;
}
void xref_unset_blas() { xref::blas() = xref::Blas(); }

// cfg: [M, F, n_slots, sigma_img, sigma_range, rho_0, sigma_rho_0, min_track_length, sigma_landmark, ci_msckf_w,
//       ci_slam_w, iekf_iter, g[3], n_w, n_bw, n_a, n_ba, a_m_max, time_margin]   (21 doubles)
void* xref_create(const double* cfg) {
  try {
    return new VIO(cfg);
  } catch (...) {
    return nullptr;
  }
}
void xref_destroy(void* h) { delete static_cast<VIO*>(h); }

int xref_flavour() {
#ifdef MULTI_UAV
  return 1;
#else
  return 0;
#endif
}

int xref_init(void* h, const double* xvec, const double* cov_rm) {
  VIO& v = *static_cast<VIO*>(h);
  return v.init(state_from(xvec, cov_rm, v.M, v.F));
}

// Ekf::processImu; returns 1 and the propagated state, 0 for std::nullopt
int xref_process_imu(void* h, double t, unsigned seq, const double* w, const double* a, double* xvec_out) {
  VIO& v = *static_cast<VIO*>(h);
  std::optional<State> s = v.ekf.processImu(t, seq, Vector3(w[0], w[1], w[2]), Vector3(a[0], a[1], a[2]));
  if (!s) return 0;
  state_to(*s, xvec_out, v.M, v.F);
  return 1;
}

// The five track lists of vio_updater.cpp:172-179 in CSR form (n[k] tracks, off_k offsets into obs_k (x,y) pairs)
// + lost SLAM feature indexes; ids (optional, per list) become Track ids.
int xref_set_measurement(void* h, double ts, const int* n, const int* const* off, const double* const* obs,
                         const unsigned long long* const* ids, const int* lost, int n_lost) {
  VIO& v = *static_cast<VIO*>(h);
  TrackManager::Lists& L = v.lists();
  L.slam = make_tracks(ts, n[0], off[0], obs[0], ids ? ids[0] : nullptr);
  L.msckf = make_tracks(ts, n[1], off[1], obs[1], ids ? ids[1] : nullptr);
  L.msckf_short = make_tracks(ts, n[2], off[2], obs[2], ids ? ids[2] : nullptr);
  L.new_slam_std = make_tracks(ts, n[3], off[3], obs[3], ids ? ids[3] : nullptr);
  L.new_slam_msckf = make_tracks(ts, n[4], off[4], obs[4], ids ? ids[4] : nullptr);
  L.lost.assign(lost, lost + n_lost);
  VioMeasurement m;
  m.timestamp = ts;
  v.updater.setMeasurement(m);
  return 0;
}

// VioMeasurement::range / sun_angle (include/x/vio/types.h:223-254, 300-305) of the measurement set by the last
// xref_set_measurement, and the facet TrackManager::featureTriangleAtPoint reports for the LRF beam (vio_updater.cpp:
// 365-366; the Delaunay lookup itself is front-end code).  range_ts <= 0.1 / sun_ts <= -1 leave the sensor unused.
int xref_set_sensors(void* h, double range_ts, double range, double img_x_n, double img_y_n, const int* tri, int n_tri,
                     double sun_ts, double sun_x, double sun_y) {
  VIO& v = *static_cast<VIO*>(h);
  VioMeasurement& m = v.updater.measurement_;
  m.range.timestamp = range_ts;
  m.range.range = range;
  m.range.img_pt_n.setX(img_x_n);
  m.range.img_pt_n.setY(img_y_n);
  v.lists().tri.assign(tri, tri + n_tri);
  m.sun_angle.timestamp = sun_ts;
  m.sun_angle.x_angle = sun_x;
  m.sun_angle.y_angle = sun_y;
  return 0;
}

// Ekf::processUpdateMeasurement; returns 1 and the updated state, 0 for std::nullopt.  seconds (optional) receives
// the wall time of the call.
int xref_process_update(void* h, double* xvec_out, double* seconds) {
  VIO& v = *static_cast<VIO*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  std::optional<State> s = v.ekf.processUpdateMeasurement();
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (!s) return 0;
  v.last = s;
  state_to(*s, xvec_out, v.M, v.F);
  return 1;
}

// which = -1: newest state of the ring buffer; -2: the state returned by the last update; >= 0: ring slot
int xref_get_state(void* h, int which, double* xvec_out, double* cov_rm) {
  VIO& v = *static_cast<VIO*>(h);
  const State* s = nullptr;
  if (which == -2) {
    if (!v.last) return -1;
    s = &*v.last;
  } else if (which == -1) {
    s = &v.ekf.state_buffer_.getTailStateRef();
  } else {
    if (which >= int(v.ekf.state_buffer_.size())) return -1;
    s = &v.ekf.state_buffer_[size_t(which)];
  }
  state_to(*s, xvec_out, v.M, v.F);
  if (cov_rm) cov_to(s->getCovariance(), cov_rm);
  return 0;
}

int xref_sm_info(void* h, int* n_poses, int* n_features, int* anchors) {
  VIO& v = *static_cast<VIO*>(h);
  *n_poses = int(v.sm().poseSize());
  *n_features = int(v.sm().getNFeatures());
  const std::vector<int> a = v.sm().getAnchorIdxs();
  for (size_t i = 0; i < a.size(); ++i) anchors[i] = a[i];
  return 0;
}

// StateManager bookkeeping (n_poses_, n_features_, anchor_idxs_, stateHasBeenFilledBefore_) set directly, so that an
// update can start from a prior produced elsewhere (e.g. downloaded from the device).
int xref_sm_set(void* h, int n_poses, int n_features, const int* anchors, int filled_before) {
  VIO& v = *static_cast<VIO*>(h);
  StateManager& sm = v.sm();
  sm.n_poses_ = n_poses;
  sm.n_features_ = size_t(n_features);
  for (size_t i = 0; i < sm.anchor_idxs_.size(); ++i) sm.anchor_idxs_[i] = anchors[i];
  sm.stateHasBeenFilledBefore_ = filled_before != 0;
  return 0;
}

// Updater::update (updater.cpp:39-115) on a caller-provided state with the measurement set by xref_set_measurement.
int xref_updater_update(void* h, double* xvec, double* cov_rm, double* seconds) {
  VIO& v = *static_cast<VIO*>(h);
  State s = state_from(xvec, cov_rm, v.M, v.F);
  const auto t0 = std::chrono::steady_clock::now();
  v.updater.update(s);
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  state_to(s, xvec, v.M, v.F);
  cov_to(s.getCovariance(), cov_rm);
  return 0;
}

// ---- stage-level entry points (methods VioUpdater / Updater keep protected; VIO is a friend) ----------------

// Updater::applyUpdate (updater.cpp:117-141) on a caller-provided state: H (m x N), res (m), R diagonal (m),
// correction_total (N, in/out).  State and covariance are overwritten with the result.
int xref_apply_update(void* h, double* xvec, double* cov_rm, const double* H_rm, const double* res,
                      const double* r_diag, int m, double* correction_total, int cov_update) {
  VIO& v = *static_cast<VIO*>(h);
  const int N = 15 + 6 * v.M + 3 * v.F;
  State s = state_from(xvec, cov_rm, v.M, v.F);
  Matrix H(m, N), r(m, 1), R = Matrix::Zero(m, m), ct(N, 1);
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < N; ++j) H(i, j) = H_rm[size_t(i) * N + j];
    r(i, 0) = res[i];
    R(i, i) = r_diag[i];
  }
  for (int j = 0; j < N; ++j) ct(j, 0) = correction_total[j];
  v.updater.applyUpdate(s, H, r, R, ct, cov_update != 0);
  for (int j = 0; j < N; ++j) correction_total[j] = ct(j, 0);
  state_to(s, xvec, v.M, v.F);
  cov_to(s.getCovariance(), cov_rm);
  return 0;
}

// VioUpdater::applyQRDecomposition (vio_updater.cpp:487-512): H (m x n) and res (m) in, compressed H (n x n) and
// res (n) out when m > n + 1 (returns the output row count).
int xref_qr_compress(void* h, const double* H_rm, const double* res, int m, int n, double* H_out_rm,
                     double* res_out) {
  VIO& v = *static_cast<VIO*>(h);
  Matrix H(m, n), r(m, 1), R = Matrix::Identity(m, m);
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) H(i, j) = H_rm[size_t(i) * n + j];
    r(i, 0) = res[i];
  }
  v.updater.applyQRDecomposition(H, r, R);
  const int mo = int(H.rows());
  for (int i = 0; i < mo; ++i) {
    for (int j = 0; j < n; ++j) H_out_rm[size_t(i) * n + j] = H(i, j);
    res_out[i] = r(i, 0);
  }
  return mo;
}

// StateManager::manage (state_manager.cpp:31-149) on a caller-provided state, with the harness's bookkeeping
// (n_poses, n_features, anchors) as it stands.
int xref_manage(void* h, double* xvec, double* cov_rm, const int* lost, int n_lost) {
  VIO& v = *static_cast<VIO*>(h);
  State s = state_from(xvec, cov_rm, v.M, v.F);
  std::vector<unsigned int> del(lost, lost + n_lost);
  v.sm().manage(s, del);
  state_to(s, xvec, v.M, v.F);
  cov_to(s.getCovariance(), cov_rm);
  return 0;
}

// Propagator::propagateState + propagateCovariance (propagator.cpp:30-72) from state 0 to the IMU sample of state 1
int xref_propagate(void* h, const double* xvec0, const double* cov0_rm, double t1, const double* w1, const double* a1,
                   double* xvec1, double* cov1_rm) {
  VIO& v = *static_cast<VIO*>(h);
  State s0 = state_from(xvec0, cov0_rm, v.M, v.F);
  State s1(v.M, v.F);
  s1.setImu(t1, s0.getSeq() + 1, Vector3(w1[0], w1[1], w1[2]), Vector3(a1[0], a1[1], a1[2]));
  v.ekf.propagator_.propagateState(s0, s1);
  v.ekf.propagator_.propagateCovariance(s0, s1);
  state_to(s1, xvec1, v.M, v.F);
  cov_to(s1.getCovariance(), cov1_rm);
  return 0;
}

// MsckfUpdate on a caller-provided state (msckf_update.cpp:27-63): returns the number of inlier rows; gamma is not
// exposed by the reference, the stacked Jacobian / residual are (rows x N row-major, rows).
int xref_msckf_rows(void* h, const double* xvec, const double* cov_rm, double ts, int n, const int* off,
                    const double* obs, double* jac_rm, double* res, int max_rows) {
#ifdef MULTI_UAV
  (void)h, (void)xvec, (void)cov_rm, (void)ts, (void)n, (void)off, (void)obs, (void)jac_rm, (void)res, (void)max_rows;
  return -1;
#else
  VIO& v = *static_cast<VIO*>(h);
  const int N = 15 + 6 * v.M + 3 * v.F;
  State s = state_from(xvec, cov_rm, v.M, v.F);
  const TranslationList G_p_C = v.sm().convertCameraPositionsToList(s);
  const AttitudeList C_q_G = v.sm().convertCameraAttitudesToList(s);
  const TrackList trks = make_tracks(ts, n, off, obs, nullptr);
  MsckfUpdate msckf(trks, C_q_G, G_p_C, s.getCovariance(), s.nPosesMax(), v.updater.sigma_img_);
  const Matrix& J = msckf.getJacobian();
  const Matrix& r = msckf.getResidual();
  const Eigen::VectorXd& d = msckf.getCovDiag();
  // rows beyond the inliers keep zero Jacobian and unit noise (msckf_update.cpp:50-52)
  int rows = 0;
  const double var = v.updater.sigma_img_ * v.updater.sigma_img_;
  for (Eigen::Index i = 0; i < J.rows(); ++i)
    if (d(i) == var) rows = int(i) + 1;
  if (rows > max_rows) return -2;
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < N; ++j) jac_rm[size_t(i) * N + j] = J(i, j);
    res[i] = r(i, 0);
  }
  return rows;
#endif
}

// RangeUpdate (range_update.cpp:24-265) and SolarUpdate (solar_update.cpp:25-94) on a caller-provided state: the
// Jacobian rows (1 x N, 2 x N), residuals and noise variances as the classes leave them.  sm bookkeeping (n_poses,
// anchors) is the harness filter's own (xref_sm_set).
int xref_sensor_rows(void* h, const double* xvec, const double* cov_rm, double range, double img_x_n, double img_y_n,
                     const int* tri, double sun_x, double sun_y, double* jac_rm, double* res, double* r_diag) {
  VIO& v = *static_cast<VIO*>(h);
  const int N = 15 + 6 * v.M + 3 * v.F;
  State s = state_from(xvec, cov_rm, v.M, v.F);
  const TranslationList G_p_C = v.sm().convertCameraPositionsToList(s);
  const AttitudeList C_q_G = v.sm().convertCameraAttitudesToList(s);
  RangeMeasurement rm;
  rm.timestamp = 1.0;
  rm.range = range;
  rm.img_pt_n.setX(img_x_n);
  rm.img_pt_n.setY(img_y_n);
  const std::vector<int> ids(tri, tri + 3);
  const RangeUpdate ru(rm, ids, C_q_G, G_p_C, s.getFeatureArray(), v.sm().getAnchorIdxs(), s.getCovariance(),
                       s.nPosesMax(), v.updater.sigma_range_);
  SunAngleMeasurement sa;
  sa.timestamp = 1.0;
  sa.x_angle = sun_x;
  sa.y_angle = sun_y;
  const SolarUpdate su(sa, s.getOrientation(), s.getCovariance());
  for (int j = 0; j < N; ++j) {
    jac_rm[j] = ru.getJacobian()(0, j);
    jac_rm[N + j] = su.getJacobian()(0, j);
    jac_rm[2 * N + j] = su.getJacobian()(1, j);
  }
  res[0] = ru.getResidual()(0, 0), res[1] = su.getResidual()(0, 0), res[2] = su.getResidual()(1, 0);
  r_diag[0] = ru.getCovDiag()(0), r_diag[1] = su.getCovDiag()(0), r_diag[2] = su.getCovDiag()(1);
  return 0;
}

// VioUpdater::preProcess + constructUpdate (vio_updater.cpp:126-198, 266-423; single-UAV build) on a caller-provided
// state AFTER StateManager::manage: the stacked (and, when rows > N + 1, QR-compressed) h, res and diag(R).
int xref_construct_update(void* h, const double* xvec, const double* cov_rm, double* H_rm, double* res, double* r_diag,
                          int max_rows) {
#ifdef MULTI_UAV
  (void)h, (void)xvec, (void)cov_rm, (void)H_rm, (void)res, (void)r_diag, (void)max_rows;
  return -1;
#else
  VIO& v = *static_cast<VIO*>(h);
  const int N = 15 + 6 * v.M + 3 * v.F;
  State s = state_from(xvec, cov_rm, v.M, v.F);
  const VioMeasurement keep = v.updater.measurement_;
  v.updater.preProcess(s);
  v.updater.measurement_ = keep;
  Matrix H, r, R;
  v.updater.constructUpdate(s, H, r, R);
  const int rows = int(H.rows());
  if (rows > max_rows) return -2;
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < N; ++j) H_rm[size_t(i) * N + j] = H(i, j);
    res[i] = r(i, 0);
    r_diag[i] = R(i, i);
  }
  return rows;
#endif
}

#ifdef MULTI_UAV
// A peer snapshot (x::SimpleState, simple_state.h:30-75)
struct XrefPeer {
  int M, F;
  const double* dynamic;       // 16
  const double* positions;     // 3M
  const double* orientations;  // 4M
  const double* features;      // 3F
  const int* anchors;          // F
  const double* cov_rm;        // N x N
};
static std::shared_ptr<SimpleState> make_peer(const XrefPeer& p) {
  const int N = 15 + 6 * p.M + 3 * p.F;
  Vectorx dyn(16), pos(3 * p.M), ori(4 * p.M), fe(3 * p.F);
  for (int i = 0; i < 16; ++i) dyn(i) = p.dynamic ? p.dynamic[i] : 0.0;
  for (int i = 0; i < 3 * p.M; ++i) pos(i) = p.positions[i];
  for (int i = 0; i < 4 * p.M; ++i) ori(i) = p.orientations[i];
  for (int i = 0; i < 3 * p.F; ++i) fe(i) = p.features[i];
  Matrix cov(N, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) cov(i, j) = p.cov_rm[size_t(i) * N + j];
  std::vector<int> an(p.anchors, p.anchors + p.F);
  return std::make_shared<SimpleState>(dyn, pos, ori, fe, cov, an);
}

// VioUpdater::msckf_matches_ for the next update (tracker_.getMsckfMatches(), vio_updater.cpp:185):
// match j = (peer index, own track id, received track obs).
int xref_set_msckf_matches(void* h, const XrefPeer* peers, int n_peers, int n_matches, const int* peer_of,
                           const unsigned long long* own_track_id, const int* off, const double* obs, double ts) {
  VIO& v = *static_cast<VIO*>(h);
  std::vector<std::shared_ptr<SimpleState>> ps;
  for (int i = 0; i < n_peers; ++i) ps.push_back(make_peer(peers[i]));
  MsckfMatches& mm = v.matches().msckf_matches;
  mm.clear();
  for (int j = 0; j < n_matches; ++j) {
    TrackPtr t = std::make_shared<Track>();
    for (int i = off[j]; i < off[j + 1]; ++i) t->push_back(Feature(ts, obs[2 * i], obs[2 * i + 1], 0.0));
    mm.emplace_back(peer_of[j], own_track_id[j], t->getId(), t, ps[size_t(peer_of[j])]);
  }
  return 0;
}

// Ekf::processOthersMeasurement (ekf.cpp:143-176) with SLAM-SLAM matches (peer, current feature id, received id)
int xref_process_others(void* h, double t, const XrefPeer* peers, int n_peers, int n_matches, const int* peer_of,
                        const int* cur_id, const int* rcv_id, double* xvec_out) {
  VIO& v = *static_cast<VIO*>(h);
  std::vector<std::shared_ptr<SimpleState>> ps;
  for (int i = 0; i < n_peers; ++i) ps.push_back(make_peer(peers[i]));
  SlamMatches& sm = v.matches().slam_matches;
  sm.clear();
  for (int j = 0; j < n_matches; ++j) sm.emplace_back(peer_of[j], cur_id[j], rcv_id[j], ps[size_t(peer_of[j])]);
  // VioUpdater reads the list in preUpdateShortMsckf (vio_updater.cpp:209-215), i.e. during a visual update; the
  // collaborative entry point uses the member directly (vio_updater.cpp:76-79)
  v.updater.slam_matches_ = sm;
  std::optional<State> s;
  try {
    s = v.ekf.processOthersMeasurement(t);
  } catch (std::exception& e) {
    std::fprintf(stderr, "xref_process_others: %s\n", e.what());
    return -1;
  }
  if (!s) return 0;
  v.last = s;
  state_to(*s, xvec_out, v.M, v.F);
  return 1;
}
#endif

}  // extern "C"
