/* xb200 -- C ABI of the B200-native xVIO EKF/MSCKF hot path (libxb200.so).
 *
 * This is the drop-in boundary for jpl-x/x_multi_agent's filter back-end.  The reference has no
 * FFI of its own; its operator API is the C++ classes x::Ekf / x::Updater / x::VioUpdater /
 * x::State / x::StateManager.  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  The C++ shim in include/x/ re-implements those
 * classes' method bodies on top of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; all floating-point data is fp64.
 *   - every function returns an int status: >=0 success (1/0 = "state produced"/"std::nullopt"
 *     where the reference returns std::optional<State>), <0 an XB_E_* error;
 *     xb_last_error() returns the message of the last failure on the calling thread.
 *   - one CUDA stream per filter; calls on one filter must be serialised by the caller exactly
 *     as Ekf::lock()/unlock() does in the reference (include/x/ekf/ekf.h:128-133).
 *   - the library has NO CPU fallback: if no sm_100 device is usable xb_create fails.
 *
 * State vector ("xvec", XB_XVEC_LEN(M,F) doubles), reference members include/x/ekf/state.h:240-337:
 *   [0:3) p   [3:6) v   [6:10) q (x,y,z,w)   [10:13) b_w   [13:16) b_a   [16:20) q_ic (x,y,z,w)
 *   [20:23) p_ic   [23:26) w_m   [26:29) a_m   [29] time   [30] seq   [31] pad
 *   [32:32+3M) p_array   [..+4M) q_array (x,y,z,w per pose)   [..+3F) f_array (alpha,beta,rho)
 * Error-state / covariance order (src/x/ekf/state.cpp:201-214):
 *   [p v theta b_w b_a | p_array 3M | theta_array 3M | f_array 3F],  N = 15 + 6M + 3F.
 * Covariances cross the ABI as dense N x N fp64, ROW-major P(i,j) at [i*N+j]
 * (an Eigen column-major caller passes layout = XB_COL_MAJOR).
 */
#ifndef XB200_H_
#define XB200_H_

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define XB_API __attribute__((visibility("default")))
#else
#define XB_API
#endif

#define XB_CORE 15
#define XB_XVEC_LEN(M, F) (32 + 7 * (M) + 3 * (F))
#define XB_NERR(M, F) (15 + 6 * (M) + 3 * (F))

enum { XB_ROW_MAJOR = 0, XB_COL_MAJOR = 1 };

enum {
  XB_OK = 0,
  XB_E_INVALID = -1,      /* bad argument (std::invalid_argument in the reference) */
  XB_E_CUDA = -2,         /* CUDA runtime failure / no usable sm_100 device */
  XB_E_RUNTIME = -3,      /* std::runtime_error sites (ekf.cpp:46, ci.cpp:59-62, multi_slam_update.cpp:83-88) */
  XB_E_MISMATCH = -4,     /* init_bfr_mismatch (ekf.h:202, ekf.cpp:50-59) */
  XB_E_CAPACITY = -5,     /* more tracks/observations/features than the filter was created for */
  XB_E_UNSUPPORTED = -6,  /* reserved (no entry point returns it any more) */
  XB_E_STALE = -7         /* the addressed ring-buffer state exists but its covariance generation was recycled
                           * (measurement older than xb_config.n_generations - 1 covariance updates) */
};

typedef struct xb_filter xb_filter; /* one agent's sliding-window filter, resident on one GPU */

/* Ekf::set (src/x/ekf/ekf.cpp:32-41) + VioUpdater ctor (include/x/vio/vio_updater.h:45-49)
 * + State(n_poses, n_features) (src/x/ekf/state.cpp:23-37). */
typedef struct xb_config {
  int n_poses_max;        /* M */
  int n_features_max;     /* F */
  int n_slots;            /* state_buffer_sz (include/x/vio/types.h:188), default 250 */
  int n_generations;      /* covariance generations kept on the device; 0 = automatic (32); see DESIGN.md */
  int device;             /* CUDA device ordinal */
  int max_tracks;         /* capacity: MSCKF + short + MSCKF-SLAM tracks per update */
  int max_obs;            /* capacity: total observations over those tracks */
  int iekf_iter;          /* Updater::iekf_iter_ (updater.h:65) */
  int min_track_length;   /* kept for API parity; used by the (out-of-scope) TrackManager */
  unsigned delta_seq_imu; /* Ekf::set */
  double g[3];            /* gravity, Ekf::set */
  double n_w, n_bw, n_a, n_ba; /* ImuNoise (include/x/common/types.h:65-85) */
  double a_m_max;         /* accel-spike threshold (ekf.cpp:84) */
  double time_margin;     /* StateBuffer time margin (state_buffer.cpp:26-46) */
  double sigma_img, sigma_range, rho_0, sigma_rho_0;
  double sigma_landmark, ci_msckf_w, ci_slam_w;
  int downdate_precision; /* 0 = fp64 CUDA cores, 1 = 3xTF32 tcgen05 tensor-core downdate */
  int multi_uav;          /* 1 = Updater::update as compiled with -DMULTI_UAV (updater.cpp:58-97): CI lists of the
                           * MSCKF-MSCKF matches first, one applyUpdate, no IEKF loop; 0 = single-UAV build */
  int oc_projection;      /* 1 (default) = the reference's observability-constrained projection of the MSCKF pose
                           * Jacobians exactly as written (msckf_update.cpp:393-406); 0 = plain Jacobians.  Not a
                           * reference option: as written the projection removes real information (u_pos = C(q) g
                           * nulls the vertical position column of every observation) and the filter loses
                           * consistency within seconds on synthetic data (tests/test_cpu.py,
                           * test_oc_projection_as_written_breaks_consistency); 0 exists so that a benchmark can run
                           * the path on a consistent filter.  The oracle has the same switch. */
} xb_config;

/* One track list in CSR form: track t owns observations [off[t], off[t+1]) of `obs`, each
 * observation 2 doubles (normalised image x,y), oldest first (include/x/vision/track.h:32-85). */
typedef struct xb_track_list {
  int n_tracks;
  const int* off;      /* n_tracks+1 */
  const double* obs;   /* 2*off[n_tracks] */
} xb_track_list;

/* What VioUpdater::preProcess leaves behind (src/x/vio/vio_updater.cpp:172-179): the seam at which
 * the hot path starts. */
typedef struct xb_measurement {
  double timestamp;
  xb_track_list slam;            /* slam_trks_: entry j belongs to SLAM feature j */
  xb_track_list msckf;           /* msckf_trks_ */
  xb_track_list msckf_short;     /* msckf_short_trks_ */
  xb_track_list new_slam_std;    /* new_slam_std_trks_ */
  xb_track_list new_msckf_slam;  /* new_msckf_slam_trks_ */
  int n_lost;
  const int* lost_slam_idxs;     /* lost_slam_trk_idxs_ */
} xb_measurement;

/* RangeMeasurement (include/x/vio/types.h:223-243) + the facet of SLAM features its beam hits, as
 * TrackManager::featureTriangleAtPoint reports it to VioUpdater::constructUpdate (src/x/vio/vio_updater.cpp:358-369).
 * Used when timestamp > 0.1 and n_tr_feat_ids == 3 (0: no facet found -> no range row). */
typedef struct xb_range_measurement {
  double timestamp;
  double range;          /* [m] */
  double img_pt_n[2];    /* normalised image coordinates of the LRF beam */
  int n_tr_feat_ids;
  int tr_feat_ids[3];    /* indexes of the facet's SLAM features (slam_trks_ order) */
} xb_range_measurement;

/* SunAngleMeasurement (include/x/vio/types.h:250-254); used when timestamp > -1. */
typedef struct xb_sun_angle_measurement {
  double timestamp;
  double x_angle, y_angle;  /* [deg] */
} xb_sun_angle_measurement;

/* Another agent's snapshot: SimpleState (include/x/ekf/simple_state.h:30-75) + the match lists of
 * include/x/vision/types.h:83-116.  cov is N_peer x N_peer (layout as given). */
typedef struct xb_peer_state {
  int n_poses_max, n_features_max;
  const double* positions;     /* 3*M */
  const double* orientations;  /* 4*M (x,y,z,w) */
  const double* features;      /* 3*F */
  const int* anchor_idxs;      /* F */
  const double* cov;           /* N x N, may be NULL when `cov_blocks` is given */
  int cov_layout;
  double translation[3];
} xb_peer_state;

typedef struct xb_slam_match {   /* SlamMatch, include/x/vision/types.h:102-116 */
  int peer;                      /* index into the peers array */
  int current_feature_id;
  int received_feature_id;
} xb_slam_match;

typedef struct xb_msckf_match {  /* MsckfMatch, include/x/vision/types.h:83-100 */
  int peer;                      /* index into the peers array / slot of the gathered pose payloads */
  int which;                     /* list the own track is in: 0 = msckf_trks_, 1 = msckf_short_trks_ */
  int id_current_track;          /* index of the own track in that list (stands for Track::getId) */
  int n_obs;
  const double* obs;             /* 2*n_obs, the peer's track (received_track_ptr), oldest first */
} xb_msckf_match;

/* ---- lifetime ------------------------------------------------------------------------------- */
XB_API void xb_default_config(xb_config* cfg);                 /* reference defaults (types.h:65-85, vio/types.h:33-189) */
XB_API int xb_create(const xb_config* cfg, xb_filter** out);   /* Ekf::Ekf + Ekf::set, ekf.cpp:25-41 */
XB_API int xb_destroy(xb_filter* f);
XB_API const char* xb_last_error(void);
XB_API const char* xb_version(void);
XB_API int xb_set_stream(xb_filter* f, void* cuda_stream);     /* adopt the caller's CUDA stream */
XB_API int xb_synchronize(xb_filter* f);
XB_API int xb_n_error_states(const xb_filter* f);              /* State::nErrorStates, state.cpp:171-175 */
XB_API int xb_xvec_len(const xb_filter* f);

/* ---- x::Ekf --------------------------------------------------------------------------------- */
/* Ekf::initializeFromState (ekf.cpp:43-64).  Also clears the StateManager (vio.cpp:54-111). */
XB_API int xb_ekf_initialize_from_state(xb_filter* f, const double* xvec, const double* cov, int cov_layout);
/* Ekf::processImu (ekf.cpp:66-140): 1 = propagated state written to xvec_out (may be NULL), 0 = nullopt. */
XB_API int xb_ekf_process_imu(xb_filter* f, double timestamp, unsigned seq, const double w_m[3],
                       const double a_m[3], double* xvec_out);
/* n consecutive Ekf::processImu calls in one (ekf.cpp:66-140 applied per sample: non-increasing timestamps are skipped,
 * accelerometer spikes repeat the previous reading); the arithmetic of up to 32 samples is three launches instead of one
 * per sample (the re-propagation kernels of ekf.cpp:227-255).  w_m / a_m: 3 n doubles.  Returns the number of samples that
 * produced a state; xvec_out (may be NULL) receives the newest state.  Addition to the reference API (a caller that needs
 * every intermediate state reads them with xb_ekf_get_state). */
XB_API int xb_ekf_process_imu_batch(xb_filter* f, int n, const double* timestamps, const unsigned* seqs, const double* w_m,
                                    const double* a_m, double* xvec_out);
/* VioUpdater::setMeasurement (vio_updater.cpp:122-124) at the preProcess seam: copies the track lists
 * to the device (the only host->device traffic of an update).  Observation arrays that live in page-locked host
 * memory (xb_host_alloc, or the caller's own cudaHostRegister) are copied asynchronously straight from the caller's
 * buffer -- they must then stay unchanged until the update that uses them has completed (xb_synchronize or a returned
 * state); pageable buffers are staged through the library's pinned ring and may be reused at once. */
XB_API int xb_vio_set_measurement(xb_filter* f, const xb_measurement* m);
/* VioMeasurement::range / sun_angle of the measurement set last (either may be NULL): the range row (RangeUpdate,
 * src/x/vio/range_update.cpp:61-265) and the two sun-sensor rows (SolarUpdate, src/x/vio/solar_update.cpp:39-94) that
 * VioUpdater::constructUpdate stacks under the visual rows (vio_updater.cpp:352-403).  Call after
 * xb_vio_set_measurement (which, like VioUpdater::setMeasurement, replaces the whole measurement); each sensor
 * measurement is used by one constructUpdate only (vio_updater.cpp:381, 402). */
XB_API int xb_vio_set_sensors(xb_filter* f, const xb_range_measurement* range, const xb_sun_angle_measurement* sun);
/* cudaMallocHost / cudaFreeHost for measurement buffers (no reference counterpart: VioMeasurement is host memory). */
XB_API void* xb_host_alloc(size_t bytes);
XB_API void xb_host_free(void* p);
/* Ekf::processUpdateMeasurement (ekf.cpp:179-213): closest state, Updater::update, re-propagation.
 * 1 = updated state written to xvec_out (may be NULL: no device->host copy), 0 = nullopt. */
XB_API int xb_ekf_process_update(xb_filter* f, double* xvec_out);
/* Ekf::processOthersMeasurement (ekf.cpp:143-176) + Updater::collaborativeUpdate (updater.cpp:22-36):
 * SLAM-SLAM covariance-intersection update against peers' snapshots. */
XB_API int xb_ekf_process_others(xb_filter* f, double timestamp, const xb_peer_state* peers, int n_peers,
                          const xb_slam_match* matches, int n_matches, double* xvec_out);
/* MULTI_UAV build of Updater::update (updater.cpp:58-97): MSCKF-MSCKF matches used by the next
 * xb_ekf_process_update (VioUpdater::msckf_matches_, vio_updater.cpp:185). */
XB_API int xb_vio_set_msckf_matches(xb_filter* f, const xb_peer_state* peers, int n_peers,
                             const xb_msckf_match* matches, int n_matches);
/* Same with the peers given as gathered pose payloads resident on the device (one slot of
 * xb_ci_pose_payload_len doubles per agent): [8 header | 3M camera positions | 4M attitudes | 6M x 6M pose block
 * of the covariance] -- all a peer contributes to msckf_update.cpp:175-279. */
XB_API int xb_vio_set_msckf_matches_packed(xb_filter* f, const double* dev_gathered, int n_agents,
                                    const xb_msckf_match* matches, int n_matches);
XB_API int xb_ci_pose_payload_len(const xb_filter* f);
XB_API int xb_ci_pack_poses(xb_filter* f, int slot, double* dev_payload);
/* Per consumed match group of the last update: out[3*j..] = (inlier, gamma, chi2); returns the group count. */
XB_API int xb_mm_last_gates(xb_filter* f, int which, double* out, int max_groups);
/* Updater::applyCI over the (S, P_j, H, res) lists of the last xb_vio_construct_update (updater.cpp:64-69,88-92). */
XB_API int xb_updater_apply_ci_lists(xb_filter* f);
/* State getters on the newest state / on ring slot `slot` (state.h:74-115). slot<0: newest. */
XB_API int xb_ekf_get_state(xb_filter* f, int slot, double* xvec_out);
XB_API int xb_ekf_get_covariance(xb_filter* f, int slot, double* cov_out, int cov_layout);
XB_API int xb_ekf_newest_slot(const xb_filter* f);

/* ---- x::StateManager bookkeeping (include/x/vio/state_manager.h) ---------------------------- */
XB_API int xb_sm_n_poses(const xb_filter* f);
XB_API int xb_sm_n_features(const xb_filter* f);
XB_API int xb_sm_anchor_idxs(const xb_filter* f, int* out /* F */);
XB_API int xb_sm_set(xb_filter* f, int n_poses, int n_features, const int* anchor_idxs, int filled_before);

/* ---- stage-level entry points (Updater / VioUpdater / StateManager / Propagator methods) ------
 * They act on the filter's *work state*: xb_work_load(slot) copies a ring slot into it (the
 * `State update_state = buffer[i]` of ekf.cpp:196), xb_work_store writes it back.            */
XB_API int xb_work_load(xb_filter* f, int slot);
XB_API int xb_work_store(xb_filter* f, int slot);
XB_API int xb_work_set(xb_filter* f, const double* xvec, const double* cov, int cov_layout);
XB_API int xb_work_get(xb_filter* f, double* xvec_out, double* cov_out, int cov_layout);
/* StateManager::manage (state_manager.cpp:31-149) */
XB_API int xb_sm_manage(xb_filter* f, const int* lost_idxs, int n_lost);
/* VioUpdater::constructUpdate (vio_updater.cpp:266-423) / constructShortMsckfUpdate (:217-264):
 * builds the compressed (H, r) on the device. which: 0 = main update, 1 = short-MSCKF. */
XB_API int xb_vio_construct_update(xb_filter* f, int which);
/* `Matrix correction = Matrix::Zero(...)` (updater.cpp:43,86): clears the device-resident correction_total. */
XB_API int xb_updater_reset_correction(xb_filter* f);
/* Updater::applyUpdate (updater.cpp:117-141) on the device-resident compressed (H, r). */
XB_API int xb_updater_apply_constructed(xb_filter* f, int cov_update);
/* Updater::applyUpdate with caller-supplied dense H (m x N row-major), res (m), R diagonal (m). */
XB_API int xb_updater_apply_update(xb_filter* f, const double* H, const double* res, const double* r_diag,
                            int m, double* correction_total /* N, in/out */, int cov_update);
/* Updater::applyCI (updater.cpp:144-161): K = P_j H^T S^-1, P = (I-KH) P_j.  P_j = work cov with the
 * listed 3x3 diagonal blocks scaled by w (msckf_update.cpp:258-267, multi_slam_update.cpp:229-239). */
XB_API int xb_updater_apply_ci(xb_filter* f, const double* H, const double* res, const double* S, int m,
                        const int* scaled_block_cols, int n_blocks, double w_result);
/* VioUpdater::postUpdate (vio_updater.cpp:425-449) */
XB_API int xb_vio_post_update(xb_filter* f);
/* Updater::update (updater.cpp:39-115) = the whole template method on the work state. */
XB_API int xb_updater_update(xb_filter* f);
/* Propagator::propagateState + propagateCovariance (propagator.cpp:30-72) slot_from -> slot_to. */
XB_API int xb_propagate(xb_filter* f, int slot_from, int slot_to);

/* ---- multi-agent compressed payload ----------------------------------------------
 * A SLAM-SLAM match uses the peer only through the matched feature's world position and the 3x3 projection
 * h P h^T of the peer covariance (multi_slam_update.cpp:116-220).  Each agent packs these per SLAM feature on
 * its own GPU ([8 header doubles | 13 doubles per feature: valid, G_p_f(3), hPh^T(9)]); the slots are exchanged
 * with one all-gather and consumed by xb_ekf_process_others_packed.  Replaces shipping SimpleState's N x N
 * covariance (simple_state.h:65-74) with identical arithmetic. */
XB_API int xb_ci_payload_len(const xb_filter* f);                       /* doubles per agent slot */
XB_API int xb_ci_pack(xb_filter* f, int slot, double* dev_payload);     /* device pointer; slot<0: newest */
/* Ekf::processOthersMeasurement on gathered payloads (device pointer, n_agents slots); match.peer = slot index. */
XB_API int xb_ekf_process_others_packed(xb_filter* f, double timestamp, const double* dev_gathered, int n_agents,
                                        const xb_slam_match* matches, int n_matches, double* xvec_out);
/* introspection: per match [inlier, gamma] of the last CI step */
XB_API int xb_ci_last_gates(xb_filter* f, double* out /* 2*n_matches */, int max_matches);
/* Updater::collaborativeUpdate (src/x/ekf/updater.cpp:22-36) on the work state: stage-level twin of
 * xb_ekf_process_others. */
XB_API int xb_updater_collaborative_update(xb_filter* f, const xb_peer_state* peers, int n_peers,
                                           const xb_slam_match* matches, int n_matches);
/* The two halves of Ekf::processUpdateMeasurement (src/x/ekf/ekf.cpp:179-213) around Updater::update, for a host-side
 * template method that drives the stages itself: begin = StateBuffer::closestIdx + copy of the buffered state into the
 * work state (returns 1, 0 for std::nullopt), end = write-back + repropagateFromStateAtIdx (ekf.cpp:227-255). */
XB_API int xb_ekf_update_begin(xb_filter* f, double timestamp, double* xvec_out /* may be NULL */);
XB_API int xb_ekf_update_end(xb_filter* f, double* xvec_out);
/* Bookkeeping a host-side x::State needs to refer to a buffered covariance: time stamp of a ring slot and the serial
 * number of the covariance it holds (-1: gone); the slot the last update was applied to. */
XB_API int xb_ekf_slot_info(const xb_filter* f, int slot, double* time_out, int* serial_out);
XB_API int xb_ekf_last_update_slot(const xb_filter* f);

/* ---- introspection for tests / profiling ------------------------------------------------------ */
XB_API int xb_debug_read(xb_filter* f, const char* name, double* out, int max_doubles); /* returns count */
/* Test / measurement hook for the dense contraction kernels of Updater::applyUpdate (updater.cpp:124-136): C = beta C +
 * alpha A B^T on host buffers (op 0: as the update dispatches it, 1: cp.async kernel, 2: symmetric downdate
 * C <- (C + C^T)/2 - A A^T).  Returns 1 if the TMA-staged kernel ran; ms_out: average device time of `reps` launches. */
XB_API int xb_debug_gemm(int op, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double alpha,
                         double beta, double* C, int ldc, int reps, double* ms_out);
XB_API int xb_debug_read_int(xb_filter* f, const char* name, int* out, int max_ints);
/* per-stage CUDA-event timers on the filter's stream (bench.py roofline line); names/ms/counts must hold XB_MAX_STAGES entries */
#define XB_MAX_STAGES 32
XB_API int xb_profile_enable(xb_filter* f, int on);
XB_API int xb_profile_read(xb_filter* f, const char** names, double* ms, long long* counts, int reset);
XB_API long long xb_kernel_launches(const xb_filter* f);                /* kernels launched so far */
XB_API double xb_chi2_quantile(double p, double dof);                   /* boost::math::quantile(chi_squared) */

/* ---- track management and match import (part of f-1): the step in front of the hot path ---------------
 * Host-side list logic (no device work): TrackManager::manageTracks + checkBaseline (src/x/vio/track_manager.cpp:
 * 115-436, 576-636) fed by the 10-double match vector of VIO::importMatches (src/x/vio/vio.cpp:372-434).  The lists it
 * returns are in the CSR form xb_vio_set_measurement takes (normalised image coordinates). */
typedef struct xb_track_manager xb_track_manager;
typedef struct xb_tm_config {
  double fx, fy, cx, cy;          /* Camera (camera.cpp:27-48): fractions of the image width / height */
  double s;                       /* FOV distortion parameter (0: none) */
  unsigned img_width, img_height;
  double min_baseline_x_n, min_baseline_y_n;  /* TrackManager ctor (track_manager.cpp:27-32) */
  unsigned n_tiles_h, n_tiles_w;  /* TiledImage (tiled_image.cpp:38-52) */
  int multi_uav;                  /* 1: the -DMULTI_UAV flavour of the short-track rule (track_manager.cpp:239-262) */
} xb_tm_config;
enum { XB_TM_MSCKF = 0, XB_TM_MSCKF_SHORT = 1, XB_TM_NEW_SLAM_STD = 2, XB_TM_NEW_SLAM_MSCKF = 3, XB_TM_SLAM = 4, XB_TM_OPP = 5 };
XB_API xb_track_manager* xb_tm_create(const xb_tm_config* cfg);
XB_API void xb_tm_destroy(xb_track_manager* tm);
XB_API void xb_tm_clear(xb_track_manager* tm);                                   /* TrackManager::clear, :74-81 */
/* VIO::importMatches + TrackManager::manageTracks.  match_vector: 10 doubles per match [cam_id, t_prev, x_prev, y_prev,
 * t_cur, x_cur, y_cur, landmark xyz] (distorted pixel coordinates); cam_rots: n_rots x 4 (Attitude ax, ay, az, aw), the
 * window's camera attitudes followed by the current one (vio_updater.cpp:150-153). */
XB_API int xb_tm_manage_tracks(xb_track_manager* tm, const double* match_vector, int n_matches, const double* cam_rots,
                               int n_rots, int n_poses_max, int n_slam_features_max, int min_track_length);
/* TrackManager::get*Tracks / normalizeSlamTracks(size_out) (:36-61): sizes, then offsets[n_tracks + 1] + xy[2 * n_obs]
 * (+ optional track ids).  xb_tm_get_list returns the number of tracks. */
XB_API int xb_tm_list_size(const xb_track_manager* tm, int which, int size_out, int* n_tracks, int* n_obs);
XB_API int xb_tm_get_list(const xb_track_manager* tm, int which, int size_out, int* offsets, double* xy,
                          unsigned long long* ids);
XB_API int xb_tm_lost_slam_idxs(const xb_track_manager* tm, int* idxs, int cap);  /* getLostSlamTrackIndexes, :99-101 */
XB_API int xb_tm_remove_persistent_track(xb_track_manager* tm, unsigned idx);     /* removePersistentTracksAtIndex, :83-85 */
XB_API int xb_tm_remove_new_persistent_tracks(xb_track_manager* tm, const unsigned* idxs, int n);  /* :87-97 */
XB_API int xb_tm_set_opp_ids(xb_track_manager* tm, const unsigned long long* ids, int n);  /* setOppUpgradesMSCKF (MULTI_UAV) */
XB_API int xb_tm_counts(const xb_track_manager* tm, int* n_slam, int* n_new_slam, int* n_opp);
/* TrackManager::featureTriangleAtPoint (:443-560): the Delaunay facet of the SLAM features' last image positions that
 * contains the LRF image point (distorted pixel coordinates).  Returns 3 and the feature indexes (slam_trks_ order) for
 * xb_range_measurement::tr_feat_ids, 0 when the point lies in no facet of features.  xb_tm_delaunay_facet is the same
 * lookup on a caller-provided point set xy (n x 2). */
XB_API int xb_tm_feature_triangle_at_point(const xb_track_manager* tm, double x_dist, double y_dist, int* ids);
/* Camera::undistort + Camera::normalize of one image point (src/x/vision/camera.cpp:69-101), e.g. the LRF beam's image
 * point -> RangeMeasurement::img_pt_n (src/x/vio/vio.cpp:288-294).  out: normalised (x, y). */
XB_API int xb_tm_normalize_point(const xb_track_manager* tm, double x_dist, double y_dist, double* out);
XB_API int xb_tm_delaunay_facet(const double* xy, int n, int img_width, int img_height, double qx, double qy, int* ids);

#ifdef __cplusplus
}
#endif
#endif /* XB200_H_ */
