// C ABI + host orchestration of the device-resident filter (see include/xb200.h).
// Host logic mirrors: src/x/ekf/ekf.cpp (driver), src/x/ekf/state_buffer.cpp (ring buffer),
// src/x/ekf/updater.cpp (template method), src/x/vio/vio_updater.cpp (construct/post update),
// src/x/vio/state_manager.cpp (integer bookkeeping only; the arithmetic runs in k_manage.cu).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/xb200.h"
#include "xb_kernels.h"

namespace xb {
static thread_local long long g_launches = 0;
void count_launch() { ++g_launches; }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("XB_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}
}  // namespace xb

using namespace xb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(XB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
  } while (0)

// ---- chi-square quantile (boost::math::quantile(chi_squared_distribution<>(dof), p)) ------------------
static double gammap(double a, double x) {  // regularised lower incomplete gamma P(a, x)
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 1000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
  for (int i = 1; i < 1000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
extern "C" double xb_chi2_quantile(double p, double dof) {
  if (!(p > 0.0 && p < 1.0) || !(dof > 0.0)) return NAN;
  const double a = 0.5 * dof;
  // Wilson-Hilferty start, then safeguarded Newton on P(a, x/2) = p
  const double t = std::sqrt(2.0) * [&] {  // inverse error function via Newton on erf
    double y = 2.0 * p - 1.0, z = 0.0;
    for (int i = 0; i < 60; ++i) z -= (std::erf(z) - y) / (2.0 / std::sqrt(M_PI) * std::exp(-z * z));
    return z;
  }();
  double x = dof * std::pow(1.0 - 2.0 / (9.0 * dof) + t * std::sqrt(2.0 / (9.0 * dof)), 3.0);
  if (!(x > 0.0)) x = 1e-3;
  double lo = 0.0, hi = INFINITY;
  const double lg = std::lgamma(a);
  for (int it = 0; it < 200; ++it) {
    const double f = gammap(a, 0.5 * x) - p;
    if (f > 0.0) hi = std::min(hi, x); else lo = std::max(lo, x);
    const double pdf = 0.5 * std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - lg);
    double xn = x - f / pdf;
    if (!(xn > lo && xn < hi)) xn = std::isinf(hi) ? 2.0 * x : 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-15 * std::fabs(x)) { x = xn; break; }
    x = xn;
  }
  return x;
}

// ---- filter object -----------------------------------------------------------------------------------
struct ListDev {
  int cap_tracks = 0, cap_obs = 0;
  int n = 0, n_obs = 0, Lmax = 0;
  int* d_off = nullptr;
  double* d_obs = nullptr;
  std::vector<int> h_off;
};

enum { ST_ASSEMBLE, ST_MANAGE, ST_TRACKS, ST_GRAM, ST_CHOLG, ST_SLAMROWS, ST_BUILD, ST_TALLCHOL, ST_CORRECT, ST_DOWNDATE,
       ST_POST, ST_STORE, ST_PROPAGATE, ST_SIDE_SLAM, ST_SIDE_CHOL, ST_SIDE_MEANS, ST_SIDE_DD, ST_MM_TRI, ST_MM_CON, ST_MM_APPLY,
       ST_COUNT };
// the "side_*" stages run on the filter's internal side streams, concurrently with the stages listed before them
static const char* kStageNames[ST_COUNT] = {"assemble", "manage", "tracks", "gram", "chol_gram", "slam_rows", "build_s_pht",
                                            "tallchol", "correct", "downdate", "post_update", "store", "propagate",
                                            "side_slam_part", "side_tallchol_slam_cols", "side_prop_means", "side_downdate_slam_cols",
                                            "mm_triangulate", "mm_construct", "mm_apply_ci"};  // the mm_* spans lie inside "tracks" / between stages
static_assert(ST_COUNT <= XB_MAX_STAGES, "xb_profile_read callers size their arrays with XB_MAX_STAGES");
struct ProfSpan { int stage; cudaEvent_t e0, e1; };

struct xb_filter {
  xb_config cfg;
  // optional per-stage CUDA-event timers (xb_profile_enable / xb_profile_read)
  bool prof = false;
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> ev_pool;
  double stage_ms[ST_COUNT] = {0};
  long long stage_n[ST_COUNT] = {0};
  int M, F, N, LX, NS, NG;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // internal side streams: (a) the SLAM-column part of the Kalman update (rows, P H^T, S block, the first tile columns of
  // the Cholesky factorisation) runs next to the MSCKF track pipeline; (b) the means of the re-propagation run next to the
  // covariance downdate.  Joined back into `stream` with events; nothing outside the library sees them.
  cudaStream_t side = nullptr, side2 = nullptr, side3 = nullptr;
  cudaEvent_t ev_b0 = nullptr, ev_b1 = nullptr, ev_b2 = nullptr, ev_g0 = nullptr, ev_g1 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_side = nullptr, ev_corr = nullptr, ev_means = nullptr;
  bool side_pending = false;   // a construct call has forked the SLAM part; apply_constructed joins it
  bool slam_part_done = false; // the tall buffer already holds the SLAM-column part for the pending update
  bool side_used_corr = false; // ... and it was built with a non-zero correction_total
  // early part of the covariance downdate: sym(P) - W1s W1s^T (the SLAM columns of W1 exist after the side factorisation) is
  // written into a freshly claimed generation on side2 while the main stream is still busy with the MSCKF pipeline; apply finishes it in place
  cudaEvent_t ev_dd = nullptr;
  bool dd_pending = false;
  int dd_cols = 0;
  bool xw_final = false;       // no estimate changed since ev_corr was recorded
  bool overlap = true;         // XB_NO_OVERLAP=1 runs everything on the one stream
  int chol_share = 4;          // concurrent dataflow launches each take 1/chol_share of the co-resident CTA slots
  double* d_Bc = nullptr;      // Wsym = (W1s + W2s)/2 on the pose rows (6M x s_pad)
  double* d_Gp = nullptr;      // gathered P[pose rows, Omega]^T (32 x 6M)
  int* d_flags_g = nullptr;    // dataflow flags of the Gram factorisation (may run while the tall buffer is being factored)
  double* d_FQ2 = nullptr;     // F_d / Q_d of the re-propagation steps
  // ring buffer (state_buffer.cpp)
  double* d_xv = nullptr;      // NS x LX
  double* d_strip = nullptr;   // NS x 15 x N: [P_ii | P_iv]
  double* d_strip2 = nullptr;  // NS x 15 x N: P_vi^T, maintained only for slots with slot_asym > 0
  // Number of newest clones whose covariance blocks are unsymmetric in a slot (0: core block only, P_vi = P_iv^T).
  // > 0 only after updates WITHOUT measurement rows: the reference then never symmetrises (updater.cpp:106) and carries
  // P_vi != P_iv^T through propagation (propagator.cpp:197-203).
  std::vector<int> slot_asym;
  int asym_clones = 0;         // the same count for the work covariance
  double* d_getcov = nullptr;  // scratch of xb_ekf_get_covariance (never a work buffer)
  std::vector<double> h_time;  // mirror of State::time_
  std::vector<double> h_am;    // mirror of a_m (accel-spike substitution, ekf.cpp:119-128)
  std::vector<int> slot_gen;
  std::vector<int> slot_serial;   // stamp of the covariance a slot refers to (changes whenever the slot is rewritten)
  int serial = 0, last_update_slot = -1;
  int tail = 0, head = 0, n_valid = 0;
  int status = 0;  // 0 not initialised, 1 stand-by, 2 initialised
  unsigned last_seq = 0;
  // covariance generations
  // NG generation slots, each naming one N x N buffer of a pool of NG + 3 (generations + the two work buffers + the
  // early-downdate destination).  Storing a work covariance SWAPS its buffer into the generation table (no copy);
  // the buffer that falls out becomes the new scratch.
  std::vector<double*> gen_buf;
  double* d_spare = nullptr; // destination of the early downdate; enters the table only when the update is stored
  int cur_gen = 0;
  // work state
  double* d_xw = nullptr;   // LX
  double* d_WA = nullptr;   // N x N scratch / assembled covariance
  double* d_WB = nullptr;   // second scratch: work_load / manage ping-pong between the two, generations are claimed only
                            // for a covariance that a ring slot is going to reference
  double* d_dd = nullptr;   // destination of the pending early downdate (a claimed generation)
  // The work covariance right after xb_work_load is "virtual": d_Pw names d_WB, but its content is still the slot's strip +
  // generation.  xb_sm_manage (the first consumer in Ekf::processUpdateMeasurement unless a short-MSCKF update precedes it)
  // reads that form directly; every other consumer materialises it first (one assemble pass).
  bool virt = false;
  const double *virt_strip = nullptr, *virt_gen = nullptr, *virt_strip2 = nullptr;
  double* d_Pw = nullptr;   // points at WA or a generation
  double* d_corr = nullptr; // N correction_total
  double* d_delta = nullptr;
  double* d_FQ = nullptr;
  // state manager bookkeeping (state_manager.h)
  int n_poses = 0, n_features = 0, filled_before = 0;
  std::vector<int> anchor;
  // measurement
  double meas_time = 0.0;
  ListDev l_slam, l_msckf, l_short, l_newstd, l_newms;
  std::vector<int> lost;
  double* d_slam_chi2 = nullptr;
  // pinned staging rings: a region is reused only after the copies issued from it have completed (event wait, no
  // stream-wide synchronisation on the hot path)
  static constexpr int kPinRing = 4;
  double* h_pin = nullptr;  // kPinRing regions of pin_bytes: measurement lists
  size_t pin_bytes = 0;
  cudaEvent_t pin_ev[kPinRing] = {nullptr, nullptr, nullptr, nullptr};
  int pin_cur = 0;
  size_t ipin_ints = 0;     // kPinRing regions of ipin_ints: manage tables
  cudaEvent_t ipin_ev[kPinRing] = {nullptr, nullptr, nullptr, nullptr};
  int ipin_cur = 0;
  std::vector<double> chi90_cache;  // chi2(0.9, dof) by integer dof (slam_update.cpp:196-197), filled lazily
  // device tables / scratch
  double* d_chi95 = nullptr;
  int chi_len = 0;
  // track outputs (mode 0 / mode 1)
  double *d_ivd0, *d_gamma0, *d_B0, *d_J0;
  int* d_inl0;
  double *d_ivd1, *d_gamma1, *d_B1, *d_J1, *d_H1, *d_H2, *d_D1;
  int* d_inl1;
  // slam rows
  int* d_scols; double *d_svals, *d_sres, *d_sgamma; int* d_sinl; int* d_anchor;
  // range / sun-sensor rows (xb_vio_set_sensors): pending measurement + device rows
  int* d_wcols = nullptr; double *d_wvals = nullptr, *d_wres = nullptr, *d_wgamma = nullptr; int* d_winl = nullptr;
  xb_range_measurement sens_range{};
  xb_sun_angle_measurement sens_sun{};
  bool sens_range_on = false, sens_sun_on = false;
  int last_nw = 0;
  // gram
  double *d_partB, *d_partD, *d_blocks, *d_Tg, *d_Rg, *d_diag0;
  int gcols_pad = 0, grows_pad = 0, nz = 1;
  int* d_flags = nullptr;
  int* d_err = nullptr;
  int* h_err = nullptr;  // pinned mirror of d_err
  long long* d_trace = nullptr;  // optional tile-Cholesky timeline (XB_CHOL_TRACE=1)
  long long* d_track_prof = nullptr;  // optional per-track phase clocks (XB_TRACK_PROF=1)
  int trace_tiles = 0;
  // kalman tall buffer
  double* d_T = nullptr;
  size_t T_doubles = 0;
  int last_m = 0, last_nslam = 0, last_which = 0;
  bool constructed_any = false;
  // manage
  int *d_rowmap, *d_ccols, *d_featsrc, *d_reanch;
  double *d_cvals, *d_mscratch, *d_Tm, *d_T2;
  int* h_ipin = nullptr;
  // feature-init scratch
  double* d_fscratch = nullptr;
  // dense-H scratch
  double* d_Hdense = nullptr; size_t Hdense_doubles = 0;
  void* d_tcws = nullptr;
  // Omega (core + newest clone) bookkeeping for the non-symmetric part of P
  int *d_omega = nullptr, *d_omega_inv = nullptr, *d_tileflag = nullptr;
  double *d_om = nullptr, *d_Zb = nullptr, *d_Yb = nullptr, *d_Qb = nullptr, *d_Cb = nullptr;
  // covariance intersection
  double *d_ci_own = nullptr, *d_ci_gather = nullptr, *d_ci_rec = nullptr, *d_ci_K = nullptr, *d_ci_delta = nullptr, *d_ci_HP = nullptr;
  int *d_ci_matches = nullptr, *d_ci_last = nullptr;
  int ci_payload_len = 0, ci_max_matches = 256, ci_max_agents = 16, ci_last_n = 0;
  // multi-agent MSCKF-MSCKF matches (VioUpdater::msckf_matches_, vio_updater.h:282)
  struct MmMatch { int peer, which, trk, n_obs; size_t obs_off; };
  std::vector<MmMatch> mm_matches;
  std::vector<double> mm_obs_h;
  const double* mm_gather = nullptr;  // device: pose payload slots (own buffer or the caller's)
  double *d_mm_gather = nullptr, *d_mm_pobs = nullptr, *d_mm_chi2 = nullptr, *d_mm_ivd = nullptr, *d_mm_F0 = nullptr,
         *d_mm_rec = nullptr, *d_mm_V = nullptr, *d_mm_D = nullptr, *d_mm_K3 = nullptr, *d_mm_HP3 = nullptr;
  int *d_mm_grp = nullptr, *d_mm_ent = nullptr, *d_mm_trkgrp = nullptr, *d_mm_last = nullptr;
  int mm_pp_len = 0, mm_max_groups = 512, mm_max_entries = 1024, mm_G = 0, mm_max_tracks = 0;
  int mm_last_G[2] = {0, 0};
  // MULTI_UAV with MSCKF-MSCKF matches: the Gram stage of constructUpdate is deferred to applyUpdate so that it runs next
  // to the SLAM-column part, which can only start once the CI corrections have changed P (updater.cpp:84-97)
  bool gram_deferred = false;
  GramParams gram_gp{};
  // page-locked staging of the match-group tables (two regions, each guarded by an event: no stream synchronisation)
  char* h_mm = nullptr;
  size_t mm_pin_bytes = 0;
  int mm_pin_cur = 0;
  cudaEvent_t mm_pin_ev[2] = {nullptr, nullptr};
  std::vector<double> chi95_cache;  // chi2(0.95, 2 n_obs - 3) by n_obs (msckf_update.cpp:243-247), filled lazily
  int omega_slot = -2;
  bool corr_zero = false;  // correction_total is known to be all-zero (first IEKF iteration)
  std::vector<void*> allocs;
};

static int invalidate_early(xb_filter* f);
static cudaEvent_t prof_event(xb_filter* f);
struct StageTimer;
static void materialize(xb_filter* f);
static cudaEvent_t prof_event(xb_filter* f) {
  cudaEvent_t e;
  if (!f->ev_pool.empty()) { e = f->ev_pool.back(); f->ev_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
struct StageTimer {  // RAII: records a CUDA event pair around a stage on the given stream when profiling is on
  xb_filter* f; int idx = -1; cudaStream_t st;
  StageTimer(xb_filter* f_, int stage, cudaStream_t on = nullptr) : f(f_), st(on ? on : f_->stream) {
    if (!f->prof) return;
    ProfSpan sp{stage, prof_event(f), prof_event(f)};
    cudaEventRecord(sp.e0, st);
    f->spans.push_back(sp);
    idx = (int)f->spans.size() - 1;
  }
  ~StageTimer() { if (idx >= 0) cudaEventRecord(f->spans[idx].e1, st); }
};
extern "C" int xb_profile_enable(xb_filter* f, int on) {
  f->prof = on != 0;
  return XB_OK;
}
// Accumulated device time per stage since the last reset. names: ST_COUNT pointers (may be NULL). Returns stage count.
extern "C" int xb_profile_read(xb_filter* f, const char** names, double* ms, long long* counts, int reset) {
  cudaStreamSynchronize(f->stream);
  cudaStreamSynchronize(f->side);
  cudaStreamSynchronize(f->side2);
  cudaStreamSynchronize(f->side3);
  for (auto& sp : f->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess) { f->stage_ms[sp.stage] += t; f->stage_n[sp.stage] += 1; }
    f->ev_pool.push_back(sp.e0);
    f->ev_pool.push_back(sp.e1);
  }
  f->spans.clear();
  for (int i = 0; i < ST_COUNT; ++i) {
    if (names) names[i] = kStageNames[i];
    if (ms) ms[i] = f->stage_ms[i];
    if (counts) counts[i] = f->stage_n[i];
    if (reset) { f->stage_ms[i] = 0; f->stage_n[i] = 0; }
  }
  return ST_COUNT;
}

static int dalloc(xb_filter* f, void** p, size_t bytes) {
  if (bytes == 0) bytes = 8;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return fail(XB_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  cudaMemset(*p, 0, bytes);
  f->allocs.push_back(*p);
  return 0;
}
#define DA(ptr, count, type)                                                        \
  do {                                                                              \
    int rc_ = dalloc(f, (void**)&(ptr), sizeof(type) * (size_t)(count));            \
    if (rc_) return rc_;                                                            \
  } while (0)

static int pad32(int x) { return (x + 31) / 32 * 32; }

extern "C" void xb_default_config(xb_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->n_poses_max = 15;      // vio/types.h:141
  c->n_features_max = 15;   // vio/types.h:146
  c->n_slots = 250;         // vio/types.h:188
  c->n_generations = 0;     // 0 = automatic (32): a buffered state stays usable for 31 later covariance updates
  c->device = 0;
  c->max_tracks = 1024;
  c->max_obs = 0;
  c->iekf_iter = 1;
  c->min_track_length = 10;
  c->delta_seq_imu = 1;
  c->g[0] = 0.0; c->g[1] = 0.0; c->g[2] = -9.81;
  c->n_w = 0.0083; c->n_bw = 0.00083; c->n_a = 0.0013; c->n_ba = 0.00013;  // common/types.h:65-85 (n_ba: intended value)
  c->a_m_max = 50.0;
  c->time_margin = 0.005;
  c->sigma_img = 1.0 / 320.0;
  c->sigma_range = 0.05;
  c->rho_0 = 0.5;
  c->sigma_rho_0 = 0.25;
  c->sigma_landmark = 0.0;
  c->ci_msckf_w = -1.0;
  c->ci_slam_w = -1.0;
  c->downdate_precision = 0;
  c->oc_projection = 1;     // msckf_update.cpp:393-406 as written
}

extern "C" const char* xb_last_error(void) { return g_err.c_str(); }
extern "C" const char* xb_version(void) { return "xb200 0.1 (sm_100a)"; }
extern "C" long long xb_kernel_launches(const xb_filter*) { return g_launches; }

static int list_alloc(xb_filter* f, ListDev& l, int cap_tracks, int cap_obs) {
  l.cap_tracks = cap_tracks;
  l.cap_obs = cap_obs;
  DA(l.d_off, cap_tracks + 1, int);
  DA(l.d_obs, 2 * (size_t)cap_obs, double);
  return 0;
}

extern "C" int xb_create(const xb_config* cfg, xb_filter** out) {
  if (!cfg || !out) return fail(XB_E_INVALID, "null argument");
  if (cfg->n_poses_max < 1 || cfg->n_poses_max > 64 || cfg->n_features_max < 0 || cfg->n_slots < 1)
    return fail(XB_E_INVALID, "n_poses_max must be in [1,64], n_features_max >= 0, n_slots >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(XB_E_CUDA, "no CUDA device: libxb200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(XB_E_INVALID, "bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return fail(XB_E_CUDA, "libxb200 is built for sm_100a only");

  xb_filter* f = new xb_filter();
  f->cfg = *cfg;
  f->M = cfg->n_poses_max;
  f->F = cfg->n_features_max;
  f->N = XB_NERR(f->M, f->F);
  f->LX = XB_XVEC_LEN(f->M, f->F);
  f->NS = cfg->n_slots;
  f->NG = cfg->n_generations > 0 ? std::max(3, cfg->n_generations) : 32;
  const int M = f->M, F = f->F, N = f->N, LX = f->LX, NS = f->NS;
  if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete f;
    return fail(XB_E_CUDA, "cudaStreamCreate failed");
  }
  f->own_stream = true;
  if (cudaStreamCreateWithFlags(&f->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&f->side2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&f->side3, cudaStreamNonBlocking) != cudaSuccess)
    return fail(XB_E_CUDA, "cudaStreamCreate failed");
  for (cudaEvent_t* e : {&f->ev_fork, &f->ev_side, &f->ev_corr, &f->ev_means, &f->ev_dd, &f->ev_b0, &f->ev_b1, &f->ev_b2,
                         &f->ev_g0, &f->ev_g1})
    CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  if (const char* e = getenv("XB_NO_OVERLAP")) f->overlap = atoi(e) == 0;
  // cfg-2 sized covariances (N <= 1024): the dataflow launches run out of tiles before they run out of CTAs, a sixth of
  // the slots each leaves more of every SM to the per-track kernel next to them (measured: tools/gpu/knobs.sh)
  f->chol_share = f->N <= 1024 ? 6 : 4;
  if (const char* e = getenv("XB_CHOL_SHARE")) f->chol_share = std::max(1, atoi(e));
  const int W = 6 * M + 1;
  const int maxT = std::max(1, cfg->max_tracks);
  const int maxO = cfg->max_obs > 0 ? cfg->max_obs : maxT * M;
  const int maxT1 = std::max(1, F), maxO1 = std::max(1, F) * M;

  DA(f->d_xv, (size_t)NS * LX, double);
  DA(f->d_strip, (size_t)NS * 15 * N, double);
  DA(f->d_strip2, (size_t)NS * 15 * N, double);
  DA(f->d_getcov, (size_t)N * N, double);
  f->slot_asym.assign(NS, 0);
  f->gen_buf.assign(f->NG, nullptr);
  for (int g = 0; g < f->NG; ++g) DA(f->gen_buf[g], (size_t)N * N, double);
  DA(f->d_spare, (size_t)N * N, double);
  DA(f->d_xw, LX, double);
  DA(f->d_WA, (size_t)N * N, double);
  DA(f->d_WB, (size_t)N * N, double);
  DA(f->d_corr, N, double);
  DA(f->d_delta, N, double);
  DA(f->d_FQ, (size_t)128 * 450, double);
  f->h_time.assign(NS, -1.0);
  f->h_am.assign(3 * (size_t)NS, 0.0);
  f->slot_gen.assign(NS, -1);
  f->slot_serial.assign(NS, 0);
  f->anchor.assign(std::max(F, 1), -1);

  int rc;
  if ((rc = list_alloc(f, f->l_slam, std::max(1, F), std::max(1, F) * 4 * M))) return rc;
  if ((rc = list_alloc(f, f->l_msckf, maxT, maxO))) return rc;
  if ((rc = list_alloc(f, f->l_short, maxT, maxO))) return rc;
  if ((rc = list_alloc(f, f->l_newstd, maxT1, maxO1))) return rc;
  if ((rc = list_alloc(f, f->l_newms, maxT1, maxO1))) return rc;
  DA(f->d_slam_chi2, std::max(1, F), double);

  f->chi_len = 2 * 64 + 2;
  {
    std::vector<double> tab(f->chi_len, 0.0);
    for (int d = 1; d < f->chi_len; ++d) tab[d] = xb_chi2_quantile(0.95, (double)d);
    DA(f->d_chi95, f->chi_len, double);
    CK(cudaMemcpy(f->d_chi95, tab.data(), sizeof(double) * f->chi_len, cudaMemcpyHostToDevice));
  }
  DA(f->d_ivd0, 3 * (size_t)maxT, double);
  DA(f->d_gamma0, maxT, double);
  DA(f->d_inl0, maxT, int);
  DA(f->d_B0, (size_t)maxT * 3 * W, double);
  DA(f->d_J0, 14 * (size_t)maxO, double);
  DA(f->d_ivd1, 3 * (size_t)maxT1, double);
  DA(f->d_gamma1, maxT1, double);
  DA(f->d_inl1, maxT1, int);
  DA(f->d_B1, (size_t)maxT1 * 3 * W, double);
  DA(f->d_J1, 14 * (size_t)maxO1, double);
  DA(f->d_H1, (size_t)maxT1 * 3 * W, double);
  DA(f->d_H2, 9 * (size_t)maxT1, double);
  DA(f->d_D1, (size_t)2 * maxO1 * W, double);
  DA(f->d_scols, 15 * (size_t)std::max(1, F), int);
  DA(f->d_svals, 30 * (size_t)std::max(1, F), double);
  DA(f->d_sres, 2 * (size_t)std::max(1, F), double);
  DA(f->d_sgamma, std::max(1, F), double);
  DA(f->d_sinl, std::max(1, F), int);
  DA(f->d_anchor, std::max(1, F), int);
  DA(f->d_wcols, XB_WMAX * XB_WNZ, int);
  DA(f->d_wvals, XB_WMAX * XB_WNZ, double);
  DA(f->d_wres, XB_WMAX, double);
  DA(f->d_wgamma, 1, double);
  DA(f->d_winl, 1, int);

  const int tiles = (W + 63) / 64;
  f->nz = std::max(1, std::min(32, (148 + tiles * tiles - 1) / (tiles * tiles)));
  DA(f->d_partB, (size_t)f->nz * W * W, double);
  DA(f->d_partD, (size_t)f->nz * W * W, double);
  DA(f->d_blocks, 28 * (size_t)M, double);
  f->gcols_pad = pad32(6 * M);
  f->grows_pad = f->gcols_pad + 32;
  DA(f->d_Tg, (size_t)f->grows_pad * f->gcols_pad, double);
  DA(f->d_Rg, (size_t)f->gcols_pad * f->gcols_pad, double);
  DA(f->d_diag0, f->gcols_pad, double);

  // structured path: SLAM columns and slab columns are padded separately; dense-H path allows up to N rows
  // (+ XB_WMAX: the range / sun-sensor rows share the column group of the SLAM rows)
  const int m_pad = std::max(pad32(2 * F + XB_WMAX) + pad32(6 * M), pad32(N)), n_pad = pad32(N);
  f->T_doubles = (size_t)(m_pad + n_pad + 96) * m_pad;
  DA(f->d_T, f->T_doubles, double);
  DA(f->d_flags, (size_t)((m_pad + n_pad + 96) / 32 + 2) * (m_pad / 32) + 128, int);
  DA(f->d_flags_g, (size_t)(f->grows_pad / 32 + 2) * (f->gcols_pad / 32) + 128, int);
  DA(f->d_Bc, (size_t)6 * M * std::max(32, pad32(2 * F + XB_WMAX)), double);
  DA(f->d_FQ2, (size_t)128 * 450, double);
  DA(f->d_Gp, (size_t)32 * 6 * M, double);
  DA(f->d_omega, 32, int);
  DA(f->d_omega_inv, n_pad, int);
  DA(f->d_tileflag, n_pad / 32 + 4, int);
  DA(f->d_om, 21 * 21 + 32, double);
  f->ci_payload_len = 8 + 13 * std::max(1, F);
  DA(f->d_ci_own, f->ci_payload_len, double);
  DA(f->d_ci_gather, (size_t)f->ci_max_agents * f->ci_payload_len, double);
  DA(f->d_ci_rec, 64 * (size_t)f->ci_max_matches, double);
  DA(f->d_ci_K, 3 * (size_t)f->ci_max_matches * N, double);
  DA(f->d_ci_delta, (size_t)f->ci_max_matches * N, double);
  DA(f->d_ci_HP, 3 * (size_t)N, double);
  DA(f->d_ci_matches, 3 * (size_t)f->ci_max_matches, int);
  DA(f->d_ci_last, 4, int);
  f->mm_pp_len = 8 + 7 * M + 36 * M * M;
  f->mm_max_tracks = maxT;
  DA(f->d_mm_gather, (size_t)f->ci_max_agents * f->mm_pp_len, double);
  DA(f->d_mm_pobs, 2 * (size_t)f->mm_max_entries * M, double);
  DA(f->d_mm_chi2, f->mm_max_groups, double);
  DA(f->d_mm_ivd, 3 * (size_t)f->mm_max_groups, double);
  DA(f->d_mm_F0, 9 * (size_t)f->mm_max_groups, double);
  DA(f->d_mm_rec, (size_t)2 * XB_MM_REC * f->mm_max_groups, double);
  DA(f->d_mm_V, (size_t)6 * M * f->mm_max_groups, double);
  DA(f->d_mm_D, (size_t)N * f->mm_max_groups, double);
  DA(f->d_mm_K3, 3 * (size_t)N, double);
  DA(f->d_mm_HP3, 3 * (size_t)N, double);
  DA(f->d_mm_grp, 4 * (size_t)f->mm_max_groups, int);
  DA(f->d_mm_ent, 3 * (size_t)f->mm_max_entries, int);
  DA(f->d_mm_trkgrp, maxT, int);
  DA(f->d_mm_last, 4, int);
  DA(f->d_Zb, (size_t)n_pad * 32, double);
  DA(f->d_Yb, (size_t)n_pad * 32, double);
  DA(f->d_Qb, (size_t)n_pad * 32, double);
  DA(f->d_Cb, (size_t)4 * (n_pad + 96) * 96, double);
  DA(f->d_err, 4, int);
  if (getenv("XB_TRACK_PROF")) DA(f->d_track_prof, 12 * (size_t)maxT, long long);
  if (getenv("XB_CHOL_TRACE")) DA(f->d_trace, 10 * ((size_t)((m_pad + n_pad + 96) / 32) * (m_pad / 32) + 64), long long);

  {  // the manage tables travel as ONE host-to-device copy: [rowmap N | ccols 15 (6 + 3F) | featsrc F | reanch F]
    const size_t nc = 15 * (size_t)(6 + 3 * std::max(1, F));
    int* tab;
    DA(tab, (size_t)N + nc + 2 * (size_t)std::max(1, F), int);
    f->d_rowmap = tab;
    f->d_ccols = tab + N;
    f->d_featsrc = f->d_ccols + nc;
    f->d_reanch = f->d_featsrc + std::max(1, F);
  }
  DA(f->d_cvals, 15 * (size_t)(6 + 3 * std::max(1, F)), double);
  DA(f->d_mscratch, 7 * (size_t)M + 3 * (size_t)F + 8, double);
  DA(f->d_Tm, (size_t)(6 + 3 * std::max(1, F)) * N, double);
  DA(f->d_T2, (size_t)(6 + 3 * std::max(1, F)) * N, double);
  {
    const size_t n3 = 3 * (size_t)maxT1;
    DA(f->d_fscratch, n3 * 6 * M + 9 * (size_t)maxT1 + n3 * N + n3 * n3, double);
  }
  f->Hdense_doubles = (size_t)m_pad * N + 2 * (size_t)m_pad + 64;
  DA(f->d_Hdense, f->Hdense_doubles, double);
  DA(f->d_tcws, downdate_tc_workspace_bytes(N, m_pad), char);

  // pinned staging: measurement lists + manage tables
  f->pin_bytes = sizeof(double) * (2 * (size_t)(maxO * 2 + maxO1 * 2 + std::max(1, F) * 4 * M) + (size_t)LX + 4096) +
                 sizeof(int) * (size_t)(2 * maxT + 2 * maxT1 + 4 * std::max(1, F) + 64);
  f->pin_bytes = (f->pin_bytes + 255) / 256 * 256;
  if (cudaMallocHost((void**)&f->h_pin, f->pin_bytes * xb_filter::kPinRing) != cudaSuccess) return fail(XB_E_CUDA, "cudaMallocHost failed");
  f->ipin_ints = ((size_t)(N + 15 * (6 + 3 * std::max(1, F)) + 2 * std::max(1, F) + 64) + 63) / 64 * 64;
  if (cudaMallocHost((void**)&f->h_ipin, sizeof(int) * f->ipin_ints * xb_filter::kPinRing) != cudaSuccess)
    return fail(XB_E_CUDA, "cudaMallocHost failed");
  if (cudaMallocHost((void**)&f->h_err, sizeof(int)) != cudaSuccess) return fail(XB_E_CUDA, "cudaMallocHost failed");
  f->mm_pin_bytes = sizeof(int) * (4 * (size_t)f->mm_max_groups + 3 * (size_t)f->mm_max_entries + (size_t)f->mm_max_tracks + 8) +
                    sizeof(double) * (2 * (size_t)f->mm_max_entries * f->M + (size_t)f->mm_max_groups + 8);
  f->mm_pin_bytes = (f->mm_pin_bytes + 255) / 256 * 256;
  if (cudaMallocHost((void**)&f->h_mm, 2 * f->mm_pin_bytes) != cudaSuccess) return fail(XB_E_CUDA, "cudaMallocHost failed");
  for (auto& e : f->mm_pin_ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(XB_E_CUDA, "cudaEventCreate failed");
  *f->h_err = 0;
  for (int i = 0; i < xb_filter::kPinRing; ++i) {
    CK(cudaEventCreateWithFlags(&f->pin_ev[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&f->ipin_ev[i], cudaEventDisableTiming));
  }
  CK(cudaDeviceSynchronize());
  *out = f;
  return XB_OK;
}

extern "C" int xb_destroy(xb_filter* f) {
  if (!f) return XB_OK;
  cudaSetDevice(f->cfg.device);
  cudaStreamSynchronize(f->stream);
  if (f->side) { cudaStreamSynchronize(f->side); cudaStreamDestroy(f->side); }
  if (f->side2) { cudaStreamSynchronize(f->side2); cudaStreamDestroy(f->side2); }
  if (f->side3) { cudaStreamSynchronize(f->side3); cudaStreamDestroy(f->side3); }
  for (cudaEvent_t e : {f->ev_fork, f->ev_side, f->ev_corr, f->ev_means, f->ev_dd, f->ev_b0, f->ev_b1, f->ev_b2, f->ev_g0, f->ev_g1})
    if (e) cudaEventDestroy(e);
  for (void* p : f->allocs) cudaFree(p);
  if (f->h_pin) cudaFreeHost(f->h_pin);
  if (f->h_ipin) cudaFreeHost(f->h_ipin);
  if (f->h_err) cudaFreeHost(f->h_err);
  if (f->h_mm) cudaFreeHost(f->h_mm);
  for (auto e : f->mm_pin_ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < xb_filter::kPinRing; ++i) {
    if (f->pin_ev[i]) cudaEventDestroy(f->pin_ev[i]);
    if (f->ipin_ev[i]) cudaEventDestroy(f->ipin_ev[i]);
  }
  if (f->own_stream) cudaStreamDestroy(f->stream);
  delete f;
  return XB_OK;
}

// Page-locked host memory for measurement buffers (cudaMallocHost): track lists placed here are copied to the device
// without the staging pass.
extern "C" void* xb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) { fail(XB_E_CUDA, "cudaMallocHost failed"); return nullptr; }
  return p;
}
extern "C" void xb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int xb_set_stream(xb_filter* f, void* s) {
  if (!f) return fail(XB_E_INVALID, "null filter");
  cudaStreamSynchronize(f->stream);
  if (f->own_stream) cudaStreamDestroy(f->stream);
  f->stream = (cudaStream_t)s;
  f->own_stream = false;
  return XB_OK;
}
extern "C" int xb_synchronize(xb_filter* f) {
  // the dataflow-timeout flag rides on the same synchronisation (pinned word, no second blocking copy)
  CK(cudaMemcpyAsync(f->h_err, f->d_err, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  if (f->side_pending) CK(cudaStreamSynchronize(f->side));
  if (*f->h_err) {
    const int e = *f->h_err;
    cudaMemset(f->d_err, 0, sizeof(int));
    if (e & 2)
      return fail(XB_E_RUNTIME, "innovation covariance not positive definite (the covariance has lost definiteness)");
    return fail(XB_E_RUNTIME, "tile Cholesky dependency wait timed out");
  }
  return XB_OK;
}
extern "C" int xb_n_error_states(const xb_filter* f) { return f->N; }
extern "C" int xb_xvec_len(const xb_filter* f) { return f->LX; }
extern "C" int xb_ekf_newest_slot(const xb_filter* f) { return f->tail; }

// ---- covariance transfer helpers -----------------------------------------------------------------------
static int upload_cov(xb_filter* f, const double* cov, int layout, double* d_dst) {
  const size_t nn = (size_t)f->N * f->N;
  if (layout == XB_COL_MAJOR) {
    CK(cudaMemcpyAsync(f->d_T, cov, sizeof(double) * nn, cudaMemcpyHostToDevice, f->stream));
    transpose(f->stream, f->d_T, d_dst, f->N, f->N);
  } else {
    CK(cudaMemcpyAsync(d_dst, cov, sizeof(double) * nn, cudaMemcpyHostToDevice, f->stream));
  }
  return 0;
}
static int download_cov(xb_filter* f, const double* d_src, double* cov, int layout) {
  const size_t nn = (size_t)f->N * f->N;
  if (layout == XB_COL_MAJOR) {
    transpose(f->stream, d_src, f->d_T, f->N, f->N);
    CK(cudaMemcpyAsync(cov, f->d_T, sizeof(double) * nn, cudaMemcpyDeviceToHost, f->stream));
  } else {
    CK(cudaMemcpyAsync(cov, d_src, sizeof(double) * nn, cudaMemcpyDeviceToHost, f->stream));
  }
  CK(cudaStreamSynchronize(f->stream));
  return 0;
}

// ---- StateBuffer (state_buffer.cpp) ----------------------------------------------------------------------
static int next_idx(const xb_filter* f, int i) { return (i + 1) % f->NS; }
static int prev_idx(const xb_filter* f, int i) { return i == 0 ? f->NS - 1 : i - 1; }
static int closest_idx(const xb_filter* f, double t) {  // state_buffer.cpp:26-63
  if (t > f->h_time[f->tail] + f->cfg.time_margin) return -1;
  if (t < f->h_time[f->head] - f->cfg.time_margin) return -1;
  double off = std::fabs(t - f->h_time[f->tail]);
  int idx = prev_idx(f, f->tail);
  int count = 1;
  while (std::fabs(t - f->h_time[idx]) < off && count < f->n_valid) {
    off = std::fabs(t - f->h_time[idx]);
    idx = prev_idx(f, idx);
    ++count;
  }
  return next_idx(f, idx);
}

extern "C" int xb_sm_n_poses(const xb_filter* f) { return f->n_poses; }
extern "C" int xb_sm_n_features(const xb_filter* f) { return f->n_features; }
extern "C" int xb_sm_anchor_idxs(const xb_filter* f, int* out) {
  for (int i = 0; i < f->F; ++i) out[i] = f->anchor[i];
  return XB_OK;
}
extern "C" int xb_sm_set(xb_filter* f, int n_poses, int n_features, const int* anchor_idxs, int filled_before) {
  if (n_poses < 0 || n_poses > f->M || n_features < 0 || n_features > f->F) return fail(XB_E_INVALID, "bad counts");
  f->n_poses = n_poses;
  f->n_features = n_features;
  f->filled_before = filled_before;
  for (int i = 0; i < f->F; ++i) f->anchor[i] = anchor_idxs ? anchor_idxs[i] : -1;
  return XB_OK;
}

// Number of pose slots whose rows/columns of a host covariance are unsymmetric (the unsymmetric clones are always the
// newest ones of the window, so the count is all the device needs)
static int count_asym_clones(const xb_filter* f, const double* cov) {
  const int N = f->N, M = f->M;
  int n = 0;
  for (int s = 0; s < M; ++s) {
    bool asym = false;
    for (int c = 0; c < 6 && !asym; ++c) {
      const int r = XB_CORE + (c < 3 ? 3 * s + c : 3 * M + 3 * s + c - 3);
      for (int j = 0; j < N; ++j)
        if (cov[(size_t)r * N + j] != cov[(size_t)j * N + r]) { asym = true; break; }
    }
    n += asym;
  }
  return n;
}

// ---- Ekf::initializeFromState (ekf.cpp:43-64) -------------------------------------------------------------
extern "C" int xb_ekf_initialize_from_state(xb_filter* f, const double* xvec, const double* cov, int layout) {
  if (!f || !xvec || !cov) return fail(XB_E_INVALID, "null argument");
  CK(cudaSetDevice(f->cfg.device));
  // StateBuffer::resetFromState (state_buffer.cpp:90-102)
  std::fill(f->h_time.begin(), f->h_time.end(), -1.0);
  std::fill(f->slot_gen.begin(), f->slot_gen.end(), -1);
  f->tail = f->head = 0;
  f->n_valid = 1;
  f->cur_gen = 0;
  CK(cudaMemcpyAsync(f->d_xv, xvec, sizeof(double) * f->LX, cudaMemcpyHostToDevice, f->stream));
  int rc = upload_cov(f, cov, layout, f->gen_buf[0]);
  if (rc) return rc;
  launch_extract_strip(f->stream, f->N, f->gen_buf[0], f->d_strip);
  std::fill(f->slot_asym.begin(), f->slot_asym.end(), 0);
  f->slot_asym[0] = count_asym_clones(f, cov);
  if (f->slot_asym[0] > 0) launch_extract_strip2(f->stream, f->N, f->gen_buf[0], f->d_strip2);
  CK(cudaStreamSynchronize(f->stream));
  f->h_time[0] = xvec[XV_TIME];
  for (int e = 0; e < 3; ++e) f->h_am[e] = xvec[XV_AM + e];
  f->slot_gen[0] = 0;
  f->status = 1;
  // VIO::initAtTime clears the state manager (vio.cpp:54-111, state_manager.cpp:22-29)
  f->n_poses = 0;
  f->n_features = 0;
  f->filled_before = 0;
  std::fill(f->anchor.begin(), f->anchor.end(), -1);
  return XB_OK;
}

static PropParams prop_params(const xb_filter* f) {
  PropParams pp;
  for (int e = 0; e < 3; ++e) pp.g[e] = f->cfg.g[e];
  pp.n_w = f->cfg.n_w; pp.n_bw = f->cfg.n_bw; pp.n_a = f->cfg.n_a; pp.n_ba = f->cfg.n_ba;
  return pp;
}

static void propagate_chain(xb_filter* f, int start, int n_steps, const ImuSample& in) {
  ImuSample none{};
  if (n_steps == 1 && !getenv("XB_NO_FUSED_IMU")) {   // the processImu case: one fused launch
    launch_prop_step(f->stream, f->d_xv, f->LX, f->d_strip, f->N, f->NS, start, in, prop_params(f), f->d_FQ);
    if (f->slot_asym[start] > 0)  // P_vi' = P_vi F^T next to P_iv' = F P_iv (propagator.cpp:197-203): rare, second launch
      launch_prop_strips(f->stream, f->d_strip2, f->N, f->NS, start, 1, f->d_FQ, 1);
    return;
  }
  int done = 0;
  while (done < n_steps) {
    const int n = std::min(128, n_steps - done);
    const bool last = done + n == n_steps;
    launch_propagate(f->stream, f->d_xv, f->LX, f->d_strip, f->N, f->NS, (start + done) % f->NS, n, last ? in : none,
                     prop_params(f), f->d_FQ);
    if (f->slot_asym[start] > 0)  // P_vi' = P_vi F^T next to P_iv' = F P_iv (propagator.cpp:197-203)
      launch_prop_strips(f->stream, f->d_strip2, f->N, f->NS, (start + done) % f->NS, n, f->d_FQ, 1);
    done += n;
  }
}

// ---- Ekf::processImu (ekf.cpp:66-140) ----------------------------------------------------------------------
extern "C" int xb_ekf_process_imu(xb_filter* f, double t, unsigned seq, const double w[3], const double a[3],
                                  double* xvec_out) {
  if (!f) return fail(XB_E_INVALID, "null filter");
  if (f->status == 0) return 0;
  const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (f->status == 1) {
    if (!(an < f->cfg.a_m_max)) return 0;
    double imu[8] = {w[0], w[1], w[2], a[0], a[1], a[2], t, (double)seq};
    double* xl = f->d_xv + (size_t)f->tail * f->LX;
    CK(cudaMemcpyAsync(xl + XV_WM, imu, sizeof(double) * 8, cudaMemcpyHostToDevice, f->stream));
    CK(cudaStreamSynchronize(f->stream));
    f->h_time[f->tail] = t;
    for (int e = 0; e < 3; ++e) f->h_am[3 * f->tail + e] = a[e];
    f->last_seq = seq;
    f->status = 2;
    if (xvec_out) return xb_ekf_get_state(f, f->tail, xvec_out) < 0 ? XB_E_CUDA : 1;
    return 1;
  }
  if (t <= f->h_time[f->tail]) return 0;
  f->last_seq = seq;
  const int last = f->tail;
  double as[3] = {a[0], a[1], a[2]};
  if (!(an < f->cfg.a_m_max))
    for (int e = 0; e < 3; ++e) as[e] = f->h_am[3 * last + e];
  // StateBuffer::enqueueInPlace (state_buffer.cpp:76-88)
  f->tail = (f->tail + 1) % f->NS;
  if (f->n_valid < f->NS) ++f->n_valid; else f->head = (f->head + 1) % f->NS;
  ImuSample in;
  in.valid = 1;
  in.t = t;
  in.seq = (double)seq;
  for (int e = 0; e < 3; ++e) { in.w[e] = w[e]; in.a[e] = as[e]; }
  propagate_chain(f, last, 1, in);
  f->h_time[f->tail] = t;
  for (int e = 0; e < 3; ++e) f->h_am[3 * f->tail + e] = as[e];
  f->slot_gen[f->tail] = f->slot_gen[last];
  f->slot_asym[f->tail] = f->slot_asym[last];
  f->slot_serial[f->tail] = ++f->serial;
  if (xvec_out) return xb_ekf_get_state(f, f->tail, xvec_out) < 0 ? XB_E_CUDA : 1;
  return 1;
}

// A run of Ekf::processImu calls (ekf.cpp:66-140) in one go: the per-sample host logic (time must increase, accelerometer
// spikes repeat the last reading, ring enqueue) runs for every sample, the arithmetic of up to XB_IMU_BATCH accepted samples
// is three launches (IMU fields into their slots, all means by one CTA per step, all covariance strips through prefix
// products -- the kernels of Ekf::repropagateFromStateAtIdx) instead of one launch per sample.  Returns the number of
// samples that produced a state; xvec_out (optional) receives the newest one.
extern "C" int xb_ekf_process_imu_batch(xb_filter* f, int n, const double* t, const unsigned* seq, const double* w,
                                        const double* a, double* xvec_out) {
  if (!f || n < 0 || (n > 0 && (!t || !seq || !w || !a))) return fail(XB_E_INVALID, "null argument");
  int done = 0, i = 0;
  if (f->status == 0) return 0;
  while (i < n && f->status == 1) {  // the first accepted sample only initialises the newest state's IMU fields
    const int rc = xb_ekf_process_imu(f, t[i], seq[i], w + 3 * i, a + 3 * i, nullptr);
    if (rc < 0) return rc;
    done += rc;
    ++i;
  }
  while (i < n) {
    ImuBatch b;
    const int start = f->tail;
    int m = 0;
    double t_prev = f->h_time[f->tail];
    for (; i < n && m < XB_IMU_BATCH && m < f->NS - 1; ++i) {
      if (t[i] <= t_prev) continue;  // ekf.cpp:97
      const double an = std::sqrt(a[3 * i] * a[3 * i] + a[3 * i + 1] * a[3 * i + 1] + a[3 * i + 2] * a[3 * i + 2]);
      const int prev = f->tail;
      f->last_seq = seq[i];
      f->tail = (f->tail + 1) % f->NS;   // StateBuffer::enqueueInPlace (state_buffer.cpp:76-88)
      if (f->n_valid < f->NS) ++f->n_valid; else f->head = (f->head + 1) % f->NS;
      for (int e = 0; e < 3; ++e) {
        b.v[m][e] = w[3 * i + e];
        b.v[m][3 + e] = (an < f->cfg.a_m_max) ? a[3 * i + e] : f->h_am[3 * prev + e];  // accelerometer spike (ekf.cpp:84-92)
        f->h_am[3 * f->tail + e] = b.v[m][3 + e];
      }
      b.v[m][6] = t[i];
      b.v[m][7] = (double)seq[i];
      f->h_time[f->tail] = t[i];
      f->slot_gen[f->tail] = f->slot_gen[prev];
      f->slot_asym[f->tail] = f->slot_asym[prev];
      f->slot_serial[f->tail] = ++f->serial;
      t_prev = t[i];
      ++m;
    }
    if (m == 0) continue;
    if (m == 1) {   // a single sample: the fused one-launch step
      ImuSample in;
      in.valid = 1; in.t = b.v[0][6]; in.seq = b.v[0][7];
      for (int e = 0; e < 3; ++e) { in.w[e] = b.v[0][e]; in.a[e] = b.v[0][3 + e]; }
      propagate_chain(f, start, 1, in);
    } else {
      ImuSample none{};
      launch_imu_scatter(f->stream, f->d_xv, f->LX, f->NS, start, m, b);
      propagate_chain(f, start, m, none);
    }
    done += m;
  }
  if (xvec_out && done > 0 && xb_ekf_get_state(f, f->tail, xvec_out) < 0) return XB_E_CUDA;
  return done;
}

extern "C" int xb_ekf_get_state(xb_filter* f, int slot, double* xvec_out) {
  if (slot < 0) slot = f->tail;
  if (slot >= f->NS) return fail(XB_E_INVALID, "bad slot");
  CK(cudaMemcpyAsync(xvec_out, f->d_xv + (size_t)slot * f->LX, sizeof(double) * f->LX, cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return XB_OK;
}
extern "C" int xb_ekf_get_covariance(xb_filter* f, int slot, double* cov_out, int layout) {
  if (slot < 0) slot = f->tail;
  if (slot >= f->NS || f->slot_gen[slot] < 0) return fail(XB_E_INVALID, "slot has no valid covariance");
  // assembled into its own scratch: d_WA / d_WB may hold the work covariance of an update in flight
  launch_assemble(f->stream, f->N, f->d_strip + (size_t)slot * 15 * f->N,
                  f->gen_buf[f->slot_gen[slot]], f->d_getcov,
                  f->slot_asym[slot] > 0 ? f->d_strip2 + (size_t)slot * 15 * f->N : nullptr);
  return download_cov(f, f->d_getcov, cov_out, layout);
}

// ---- measurement upload ---------------------------------------------------------------------------------------
static int stage_list(xb_filter* f, ListDev& l, const xb_track_list& in, char*& pin, const char* name) {
  l.n = in.n_tracks;
  l.n_obs = 0;
  l.Lmax = 0;
  if (in.n_tracks <= 0) { l.n = 0; return 0; }
  if (!in.off || !in.obs) return fail(XB_E_INVALID, std::string(name) + ": null track list");
  if (in.n_tracks > l.cap_tracks) return fail(XB_E_CAPACITY, std::string(name) + ": too many tracks");
  l.n_obs = in.off[in.n_tracks] - in.off[0];
  if (l.n_obs > l.cap_obs) return fail(XB_E_CAPACITY, std::string(name) + ": too many observations");
  l.h_off.assign(in.off, in.off + in.n_tracks + 1);
  for (int t = 0; t < in.n_tracks; ++t) l.Lmax = std::max(l.Lmax, in.off[t + 1] - in.off[t]);
  int* po = (int*)pin;
  const int o0 = in.off[0];
  for (int t = 0; t <= in.n_tracks; ++t) po[t] = in.off[t] - o0;
  for (int t = 0; t <= in.n_tracks; ++t) l.h_off[t] -= o0;
  pin += sizeof(int) * (((size_t)in.n_tracks + 1 + 1) / 2 * 2);
  CK(cudaMemcpyAsync(l.d_off, po, sizeof(int) * ((size_t)in.n_tracks + 1), cudaMemcpyHostToDevice, f->stream));
  // observations: straight from the caller's buffer when it is page-locked (xb_host_alloc / cudaHostRegister), else through
  // the pinned staging ring (one extra host copy, ~45 us for the 0.4 MB of a cfg-2 update)
  const double* src = in.obs + 2 * (size_t)o0;
  cudaPointerAttributes pa;
  const bool pinned = l.n_obs >= 256 && cudaPointerGetAttributes(&pa, src) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  cudaGetLastError();  // an unregistered host pointer is not an error here
  if (!pinned) {
    double* pd = (double*)pin;
    std::memcpy(pd, src, sizeof(double) * 2 * (size_t)l.n_obs);
    pin += sizeof(double) * 2 * (size_t)l.n_obs;
    src = pd;
  }
  CK(cudaMemcpyAsync(l.d_obs, src, sizeof(double) * 2 * (size_t)l.n_obs, cudaMemcpyHostToDevice, f->stream));
  return 0;
}

extern "C" int xb_vio_set_measurement(xb_filter* f, const xb_measurement* m) {
  if (!f || !m) return fail(XB_E_INVALID, "null argument");
  CK(cudaSetDevice(f->cfg.device));
  f->pin_cur = (f->pin_cur + 1) % xb_filter::kPinRing;
  CK(cudaEventSynchronize(f->pin_ev[f->pin_cur]));  // staging region reuse: its last copies have landed
  f->meas_time = m->timestamp;
  char* pin = (char*)f->h_pin + f->pin_bytes * f->pin_cur;
  int rc;
  if ((rc = stage_list(f, f->l_slam, m->slam, pin, "slam"))) return rc;
  if ((rc = stage_list(f, f->l_msckf, m->msckf, pin, "msckf"))) return rc;
  if ((rc = stage_list(f, f->l_short, m->msckf_short, pin, "msckf_short"))) return rc;
  if ((rc = stage_list(f, f->l_newstd, m->new_slam_std, pin, "new_slam_std"))) return rc;
  if ((rc = stage_list(f, f->l_newms, m->new_msckf_slam, pin, "new_msckf_slam"))) return rc;
  if (f->l_msckf.Lmax > 64 || f->l_short.Lmax > 64 || f->l_newms.Lmax > 64)
    return fail(XB_E_CAPACITY, "track longer than 64 observations");
  f->lost.assign(m->lost_slam_idxs, m->lost_slam_idxs + std::max(0, m->n_lost));
  if (f->l_slam.n > 0) {  // chi2(0.9, 2*track_size) per SLAM track (slam_update.cpp:196-197)
    double* pc = (double*)pin;
    for (int j = 0; j < f->l_slam.n; ++j) {
      const size_t dof = 2 * (size_t)(f->l_slam.h_off[j + 1] - f->l_slam.h_off[j]);
      if (dof >= f->chi90_cache.size()) f->chi90_cache.resize(dof + 1, -1.0);
      if (f->chi90_cache[dof] < 0.0) f->chi90_cache[dof] = dof ? xb_chi2_quantile(0.9, (double)dof) : NAN;
      pc[j] = f->chi90_cache[dof];
    }
    CK(cudaMemcpyAsync(f->d_slam_chi2, pc, sizeof(double) * f->l_slam.n, cudaMemcpyHostToDevice, f->stream));
  }
  CK(cudaEventRecord(f->pin_ev[f->pin_cur], f->stream));
  f->sens_range_on = f->sens_sun_on = false;  // setMeasurement replaces the whole VioMeasurement (vio_updater.cpp:122-124)
  return XB_OK;
}

// VioMeasurement::range / sun_angle (include/x/vio/types.h:300-305) of the measurement set last; each is consumed by the
// first constructUpdate that uses it (vio_updater.cpp:381, 402).
extern "C" int xb_vio_set_sensors(xb_filter* f, const xb_range_measurement* range, const xb_sun_angle_measurement* sun) {
  if (!f) return fail(XB_E_INVALID, "null argument");
  f->sens_range_on = f->sens_sun_on = false;
  if (range && range->timestamp > 0.1 && range->n_tr_feat_ids > 0) {  // vio_updater.cpp:358, 369
    if (range->n_tr_feat_ids != 3) return fail(XB_E_INVALID, "range: the facet has three SLAM feature ids");
    for (int j = 0; j < 3; ++j)
      if (range->tr_feat_ids[j] < 0 || range->tr_feat_ids[j] >= f->F)
        return fail(XB_E_INVALID, "range: facet feature id out of range");
    f->sens_range = *range;
    f->sens_range_on = true;
  }
  if (sun && sun->timestamp > -1.0) {  // vio_updater.cpp:390
    f->sens_sun = *sun;
    f->sens_sun_on = true;
  }
  return XB_OK;
}

// ---- work state ---------------------------------------------------------------------------------------------------
extern "C" int xb_work_load(xb_filter* f, int slot) {
  if (slot < 0 || slot >= f->NS || f->slot_gen[slot] < 0) return fail(XB_E_INVALID, "slot has no valid state");
  if (invalidate_early(f)) return XB_E_CUDA;
  StageTimer st_(f, ST_ASSEMBLE);
  CK(cudaMemcpyAsync(f->d_xw, f->d_xv + (size_t)slot * f->LX, sizeof(double) * f->LX, cudaMemcpyDeviceToDevice, f->stream));
  f->virt_strip = f->d_strip + (size_t)slot * 15 * f->N;
  f->virt_gen = f->gen_buf[f->slot_gen[slot]];
  f->virt_strip2 = f->slot_asym[slot] > 0 ? f->d_strip2 + (size_t)slot * 15 * f->N : nullptr;
  f->asym_clones = f->slot_asym[slot];
  f->d_Pw = f->d_WB;
  f->virt = true;
  if (!f->overlap) materialize(f);   // XB_NO_OVERLAP=1 also switches this shortcut off (plain, ordered reference schedule)
  return XB_OK;
}
static void materialize(xb_filter* f) {
  if (!f->virt) return;
  launch_assemble(f->stream, f->N, f->virt_strip, f->virt_gen, f->d_WB, f->virt_strip2);
  f->virt = false;
}
// claim the next covariance generation as destination; slots still pointing at it lose their state
static int claim_generation(xb_filter* f) {
  f->cur_gen = (f->cur_gen + 1) % f->NG;
  for (int s = 0; s < f->NS; ++s)
    if (f->slot_gen[s] == f->cur_gen) { f->slot_gen[s] = -1; }
  return f->cur_gen;
}
static int work_store_impl(xb_filter* f, int slot, bool copy_estimates) {
  if (slot < 0 || slot >= f->NS) return fail(XB_E_INVALID, "bad slot");
  materialize(f);
  StageTimer st_(f, ST_STORE);
  {
    // the work buffer becomes the claimed generation; the buffer it replaces becomes scratch (no copy)
    const int g = claim_generation(f);
    double* old = f->gen_buf[g];
    f->gen_buf[g] = f->d_Pw;
    if (f->d_Pw == f->d_WA) f->d_WA = old;
    else if (f->d_Pw == f->d_WB) f->d_WB = old;
    else if (f->d_Pw == f->d_spare) f->d_spare = old;
    else return fail(XB_E_RUNTIME, "work covariance is not a pool buffer");
  }
  if (copy_estimates)
    CK(cudaMemcpyAsync(f->d_xv + (size_t)slot * f->LX, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToDevice, f->stream));
  launch_extract_strip(f->stream, f->N, f->d_Pw, f->d_strip + (size_t)slot * 15 * f->N);
  if (f->asym_clones > 0) launch_extract_strip2(f->stream, f->N, f->d_Pw, f->d_strip2 + (size_t)slot * 15 * f->N);
  f->slot_asym[slot] = f->asym_clones;
  f->slot_gen[slot] = f->cur_gen;
  f->slot_serial[slot] = ++f->serial;
  return XB_OK;
}
extern "C" int xb_work_store(xb_filter* f, int slot) { return work_store_impl(f, slot, true); }
extern "C" int xb_work_set(xb_filter* f, const double* xvec, const double* cov, int layout) {
  CK(cudaSetDevice(f->cfg.device));
  if (invalidate_early(f)) return XB_E_CUDA;
  if (cov) f->virt = false; else materialize(f);
  if (xvec) CK(cudaMemcpyAsync(f->d_xw, xvec, sizeof(double) * f->LX, cudaMemcpyHostToDevice, f->stream));
  if (cov) {
    int rc = upload_cov(f, cov, layout, f->d_WA);
    if (rc) return rc;
    f->d_Pw = f->d_WA;
    f->asym_clones = count_asym_clones(f, cov);
  }
  CK(cudaStreamSynchronize(f->stream));
  return XB_OK;
}
extern "C" int xb_work_get(xb_filter* f, double* xvec_out, double* cov_out, int layout) {
  if (cov_out) materialize(f);
  if (xvec_out) CK(cudaMemcpyAsync(xvec_out, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToHost, f->stream));
  if (cov_out) {
    if (!f->d_Pw) return fail(XB_E_INVALID, "no work covariance");
    if (f->d_Pw == f->d_WA && layout == XB_ROW_MAJOR) {
      CK(cudaMemcpyAsync(cov_out, f->d_Pw, sizeof(double) * (size_t)f->N * f->N, cudaMemcpyDeviceToHost, f->stream));
    } else {
      int rc = download_cov(f, f->d_Pw, cov_out, layout);
      if (rc) return rc;
    }
  }
  CK(cudaStreamSynchronize(f->stream));
  return XB_OK;
}

// The SLAM-column part of a constructed update was computed early (side stream) from the P / estimates / correction_total
// of that moment.  Anything that changes one of them before apply_constructed discards it; apply then rebuilds it in order.
static int invalidate_early(xb_filter* f) {
  if (f->dd_pending) CK(cudaStreamWaitEvent(f->stream, f->ev_dd, 0));
  f->dd_pending = false;
  if (f->side_pending) CK(cudaStreamWaitEvent(f->stream, f->ev_side, 0));
  f->side_pending = false;
  f->slam_part_done = false;
  return 0;
}

// ---- StateManager::manage (state_manager.cpp:31-149): integer bookkeeping here, arithmetic on the device ----------
extern "C" int xb_sm_manage(xb_filter* f, const int* lost_idxs, int n_lost) {
  const int M = f->M, F = f->F, N = f->N;
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  f->xw_final = false;
  if (invalidate_early(f)) return XB_E_CUDA;
  StageTimer st_(f, ST_MANAGE);
  f->ipin_cur = (f->ipin_cur + 1) % xb_filter::kPinRing;
  CK(cudaEventSynchronize(f->ipin_ev[f->ipin_cur]));  // pinned table region reuse
  int* ip = f->h_ipin + f->ipin_ints * f->ipin_cur;
  int* rowmap = ip;                       // N
  int* ccols = rowmap + N;                // 15 * n_comp
  // feature removal: compaction map over ALL F slots (state_manager.cpp:52-112)
  std::vector<int> src(F), anc(F);
  for (int k = 0; k < F; ++k) { src[k] = k; anc[k] = f->anchor[k]; }
  std::vector<unsigned> del(lost_idxs, lost_idxs + std::max(0, n_lost));
  std::sort(del.begin(), del.end());
  int nf = f->n_features;
  for (size_t i = del.size(); i > 0; --i) {
    const int idx = (int)del[i - 1];
    if (idx < 0 || idx >= f->n_features) return fail(XB_E_INVALID, "lost SLAM feature index out of range");
    if (i > 1 && del[i - 2] == del[i - 1]) return fail(XB_E_INVALID, "duplicate lost SLAM feature index");
    src.erase(src.begin() + idx);
    src.push_back(-1);
    anc.erase(anc.begin() + idx);
    anc.push_back(-1);
    --nf;
  }
  const int slide = f->n_poses == M;
  std::vector<int> reanch;
  if (slide)
    for (int k = 0; k < nf; ++k)
      if (anc[k] == 0) reanch.push_back(k);
  const int n_comp = 6 + 3 * (int)reanch.size();
  const int pos = slide ? M - 1 : f->n_poses;
  const int P0 = XB_CORE, A0 = XB_CORE + 3 * M, F0 = XB_CORE + 6 * M;
  // composite row map
  for (int i = 0; i < XB_CORE; ++i) rowmap[i] = i;
  for (int s = 0; s < M; ++s)
    for (int c = 0; c < 3; ++c) {
      int vp, va;
      if (s == pos) { vp = -2 - c; va = -2 - (3 + c); }
      else if (s < pos) {
        const int so = slide ? s + 1 : s;
        vp = P0 + 3 * so + c;
        va = A0 + 3 * so + c;
      } else {
        // slots beyond the new clone: identity once the window has been filled, else zeroed by J (state_manager.cpp:276-284)
        vp = f->filled_before ? P0 + 3 * s + c : -1;
        va = f->filled_before ? A0 + 3 * s + c : -1;
        if (slide) { vp = -1; va = -1; }
      }
      rowmap[P0 + 3 * s + c] = vp;
      rowmap[A0 + 3 * s + c] = va;
    }
  for (int k = 0; k < F; ++k)
    for (int c = 0; c < 3; ++c) {
      int v;
      const bool active = k < nf;
      if (!active && !f->filled_before) v = -1;
      else v = src[k] >= 0 ? F0 + 3 * src[k] + c : -1;
      rowmap[F0 + 3 * k + c] = v;
    }
  for (int e = 0; e < 15 * n_comp; ++e) ccols[e] = 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { ccols[r * 15 + c] = c; ccols[r * 15 + 3 + c] = 6 + c; }
    for (int c = 0; c < 3; ++c) ccols[(3 + r) * 15 + c] = 6 + c;
  }
  for (size_t i = 0; i < reanch.size(); ++i) {
    const int k = reanch[i];
    for (int r = 0; r < 3; ++r) {
      int* cc = ccols + (6 + 3 * i + r) * 15;
      for (int c = 0; c < 3; ++c) {
        cc[c] = P0 + 3 * (M - 1) + c;
        cc[3 + c] = A0 + 3 * (M - 1) + c;
        cc[6 + c] = P0 + c;
        cc[9 + c] = A0 + c;
        cc[12 + c] = F0 + 3 * src[k] + c;
      }
      rowmap[F0 + 3 * k + r] = -2 - (6 + 3 * (int)i + r);
    }
  }
  const size_t nc_cap = 15 * (size_t)(6 + 3 * std::max(1, F));
  int* fs = ccols + nc_cap;   // fixed layout: mirrors the device table
  int* ra = fs + std::max(1, F);
  for (int k = 0; k < F; ++k) fs[k] = src[k];
  for (size_t i = 0; i < reanch.size(); ++i) ra[i] = reanch[i];
  CK(cudaMemcpyAsync(f->d_rowmap, rowmap, sizeof(int) * ((size_t)N + nc_cap + 2 * (size_t)std::max(1, F)), cudaMemcpyHostToDevice,
                     f->stream));
  CK(cudaEventRecord(f->ipin_ev[f->ipin_cur], f->stream));
  // destination: the scratch buffer that is not the source
  double* src_P = f->d_Pw;
  double* dst_P = (src_P == f->d_WA) ? f->d_WB : f->d_WA;
  // a work covariance that is still in its ring-slot form (strip + generation) is read as such: no assemble pass
  launch_manage_dev(f->stream, M, F, N, f->n_poses, nf, slide, (int)reanch.size(), f->d_featsrc, f->d_reanch, f->d_rowmap,
                    f->d_ccols, f->d_cvals, f->d_mscratch, f->d_xw, f->virt ? nullptr : src_P, dst_P, f->d_Tm, f->d_T2,
                    f->virt_strip, f->virt_gen, f->virt_strip2, f->asym_clones > 0);
  f->virt = false;
  f->d_Pw = dst_P;
  f->asym_clones = std::min(f->asym_clones + 1, M);  // the new clone inherits the unsymmetric core block
  // bookkeeping after the call
  for (int k = 0; k < F; ++k) f->anchor[k] = anc[k];
  f->n_features = nf;
  if (slide) {
    for (size_t i = 0; i < reanch.size(); ++i) f->anchor[reanch[i]] = M - 1;
    for (int k = 0; k < nf; ++k) f->anchor[k] -= 1;  // slideWindow, state_manager.cpp:530-533
    f->n_poses -= 1;
  }
  if (pos + 1 == M) f->filled_before = 1;
  f->n_poses += 1;
  return XB_OK;
}

// ---- VioUpdater::constructUpdate / constructShortMsckfUpdate ----------------------------------------------------------
static TrackParams track_params(xb_filter* f, const ListDev& l, int mode) {
  TrackParams tp{};
  tp.xv = f->d_xw;
  tp.M = f->M;
  tp.n_poses = f->n_poses;
  tp.P = f->d_Pw;
  tp.ldp = f->N;
  tp.off = l.d_off;
  tp.obs = l.d_obs;
  tp.n_tracks = l.n;
  tp.mode = mode;
  tp.Lmax = std::max(2, l.Lmax);
  tp.var_img = f->cfg.sigma_img * f->cfg.sigma_img;
  tp.chi2_95 = f->d_chi95;
  tp.gn_term = 1e-5;   // vio_updater.cpp:283-285
  tp.gn_max_iter = 10;
  tp.prof = mode == 0 ? f->d_track_prof : nullptr;
  tp.asym_clones = std::max(1, f->asym_clones);
  tp.oc = f->cfg.oc_projection;
  if (mode == 0) {
    tp.ivd = f->d_ivd0; tp.gamma = f->d_gamma0; tp.inlier = f->d_inl0; tp.B = f->d_B0; tp.Jout = f->d_J0;
    tp.H1 = nullptr; tp.H2 = nullptr; tp.D = nullptr;
  } else {
    tp.ivd = f->d_ivd1; tp.gamma = f->d_gamma1; tp.inlier = f->d_inl1; tp.B = f->d_B1; tp.Jout = f->d_J1;
    tp.H1 = f->d_H1; tp.H2 = f->d_H2; tp.D = f->d_D1;
  }
  return tp;
}

// Collects, for every own track of the list in order, the matches the reference's erase-loop would hand to it
// (msckf_update.cpp:96-139, restated with its shrinking loop bound), and uploads the group tables.
static int check_ci_weight(double w);
static MmParams mm_params(xb_filter* f, const ListDev& l0, int which) {
  MmParams mp{};
  mp.xv = f->d_xw; mp.M = f->M; mp.n_poses = f->n_poses; mp.N = f->N; mp.P = f->d_Pw;
  mp.off = l0.d_off; mp.obs = l0.d_obs;
  mp.grp = f->d_mm_grp; mp.ent = f->d_mm_ent; mp.pobs = f->d_mm_pobs; mp.chi2 = f->d_mm_chi2;
  mp.n_groups = f->mm_G;
  mp.gathered = f->mm_gather; mp.pp_len = f->mm_pp_len;
  mp.var_img = f->cfg.sigma_img * f->cfg.sigma_img; mp.w_other = f->cfg.ci_msckf_w;
  mp.gn_term = 1e-5; mp.gn_max_iter = 10;
  mp.oc = f->cfg.oc_projection;
  mp.B = f->d_B0; mp.inlier = f->d_inl0;
  mp.ivd = f->d_mm_ivd; mp.F0 = f->d_mm_F0; mp.rec = f->d_mm_rec + (size_t)which * XB_MM_REC * f->mm_max_groups;
  mp.last = f->d_mm_last;
  return mp;
}
static int mm_prepare(xb_filter* f, int which, const ListDev& l0, MmParams& mp) {
  auto& ms = f->mm_matches;
  std::vector<int> cnt(l0.n, 0);
  bool any = false;
  for (const auto& m : ms)
    if (m.which == which && m.trk >= 0 && m.trk < l0.n) { ++cnt[m.trk]; any = true; }
  if (!any) return 0;
  std::vector<int> grp, ent, trkgrp(l0.n, -1);
  std::vector<double> pobs, chi;
  for (int j = 0; j < l0.n; ++j) {
    if (!cnt[j]) continue;
    std::vector<xb_filter::MmMatch> mine;
    int corrected = 0;
    for (int i = 0; i < (int)ms.size(); ++i) {  // size() re-evaluated while the list shrinks, like the reference
      const auto& m = ms[i - corrected];
      if (m.which == which && m.trk == j) {
        mine.push_back(m);
        ms.erase(ms.begin() + (i - corrected));
        ++corrected;
      }
    }
    if (mine.empty()) continue;
    if ((int)mine.size() > XB_MM_KMAX) return fail(XB_E_CAPACITY, "more than 7 matched peers for one MSCKF track");
    const int g = (int)grp.size() / 4;
    if (g >= f->mm_max_groups || (int)ent.size() / 3 + (int)mine.size() > f->mm_max_entries)
      return fail(XB_E_CAPACITY, "too many MSCKF-MSCKF matches");
    int n_tot = l0.h_off[j + 1] - l0.h_off[j];
    grp.push_back(j); grp.push_back((int)mine.size()); grp.push_back((int)ent.size() / 3);
    for (const auto& m : mine) {
      ent.push_back(m.peer); ent.push_back((int)pobs.size() / 2); ent.push_back(m.n_obs);
      pobs.insert(pobs.end(), f->mm_obs_h.begin() + m.obs_off, f->mm_obs_h.begin() + m.obs_off + 2 * (size_t)m.n_obs);
      n_tot += m.n_obs;
    }
    grp.push_back(n_tot);
    if ((size_t)n_tot >= f->chi95_cache.size()) f->chi95_cache.resize((size_t)n_tot + 1, -1.0);
    if (f->chi95_cache[n_tot] < 0.0) f->chi95_cache[n_tot] = xb_chi2_quantile(0.95, 2.0 * n_tot - 3.0);  // msckf_update.cpp:243-247
    chi.push_back(f->chi95_cache[n_tot]);
    trkgrp[j] = g;
  }
  f->mm_G = (int)grp.size() / 4;
  if (f->mm_G == 0) return 0;
  int rc = check_ci_weight(f->cfg.ci_msckf_w);  // ci.cpp:59-62
  if (rc) return rc;
  // through a page-locked region guarded by an event (the tables die with this frame, the copies are asynchronous): no
  // stream synchronisation, so the host keeps enqueueing the update while the device still works on what came before
  if ((int)trkgrp.size() > f->mm_max_tracks) return fail(XB_E_CAPACITY, "too many tracks for the match-group table");
  f->mm_pin_cur ^= 1;
  CK(cudaEventSynchronize(f->mm_pin_ev[f->mm_pin_cur]));
  char* pin = f->h_mm + f->mm_pin_bytes * f->mm_pin_cur;
  auto stage = [&](void* dst, const void* src, size_t bytes) -> int {
    std::memcpy(pin, src, bytes);
    CK(cudaMemcpyAsync(dst, pin, bytes, cudaMemcpyHostToDevice, f->stream));
    pin += (bytes + 15) / 16 * 16;
    return 0;
  };
  if (stage(f->d_mm_grp, grp.data(), sizeof(int) * grp.size()) || stage(f->d_mm_ent, ent.data(), sizeof(int) * ent.size()) ||
      stage(f->d_mm_trkgrp, trkgrp.data(), sizeof(int) * trkgrp.size()) ||
      stage(f->d_mm_pobs, pobs.data(), sizeof(double) * pobs.size()) || stage(f->d_mm_chi2, chi.data(), sizeof(double) * chi.size()))
    return XB_E_CUDA;
  CK(cudaEventRecord(f->mm_pin_ev[f->mm_pin_cur], f->stream));
  mp = mm_params(f, l0, which);
  return f->mm_G;
}

static UpdateDims update_dims(const xb_filter* f, int nslam, int nw);
static int set_omega(xb_filter* f);
static size_t tall_bytes(const UpdateDims& d);
static int early_downdate(xb_filter* f, const UpdateDims& d);
static void slam_phase(xb_filter* f, cudaStream_t st, const UpdateDims& d, int stage_build, int stage_chol, int share,
                       bool with_rows);
extern "C" int xb_updater_apply_ci_lists(xb_filter* f) {
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  materialize(f);
  if (f->mm_G <= 0) return XB_OK;
  f->xw_final = false;
  if (invalidate_early(f)) return XB_E_CUDA;
  const ListDev& l0 = f->last_which == 0 ? f->l_msckf : f->l_short;
  MmParams mp = mm_params(f, l0, f->last_which);
  {
    StageTimer st_(f, ST_MM_APPLY);
    launch_mm_apply(f->stream, mp, f->d_Pw, f->d_xw, f->F, f->d_mm_V, f->mm_max_groups, f->d_mm_D, f->d_mm_K3, f->d_mm_HP3);
  }
  f->mm_G = 0;
  if (f->gram_deferred && f->last_which == 0 && f->last_nslam + f->last_nw > 0) {
    // P is final for the applyUpdate that follows: its SLAM-column half (rows linearised before the CI corrections,
    // S = H P H^T with the corrected P) goes to the side stream, the deferred Gram stage runs next to it on the caller's
    const UpdateDims d = update_dims(f, f->last_nslam, f->last_nw);
    int rc0 = set_omega(f);
    if (rc0) return rc0;
    CK(cudaMemsetAsync(f->d_T, 0, tall_bytes(d), f->stream));
    CK(cudaEventRecord(f->ev_fork, f->stream));
    CK(cudaStreamWaitEvent(f->side, f->ev_fork, 0));
    slam_phase(f, f->side, d, ST_SIDE_SLAM, ST_SIDE_CHOL, f->chol_share, false);
    CK(cudaEventRecord(f->ev_side, f->side));
    if (early_downdate(f, d)) return XB_E_CUDA;
    f->side_pending = true;
    f->slam_part_done = true;
    f->side_used_corr = !f->corr_zero;
  }
  return XB_OK;
}

static UpdateDims update_dims(const xb_filter* f, int nslam, int nw);
static int set_omega(xb_filter* f);

static void launch_slam_rows_on(xb_filter* f, cudaStream_t st, int ns) {
  SlamParams sp{};
  sp.xv = f->d_xw; sp.M = f->M; sp.N = f->N; sp.n_poses = f->n_poses; sp.P = f->d_Pw;
  sp.off = f->l_slam.d_off; sp.obs = f->l_slam.d_obs; sp.anchor = f->d_anchor; sp.chi2 = f->d_slam_chi2;
  sp.n_tracks = ns; sp.var_img = f->cfg.sigma_img * f->cfg.sigma_img;
  sp.cols = f->d_scols; sp.vals = f->d_svals; sp.res = f->d_sres; sp.gamma = f->d_sgamma; sp.inlier = f->d_sinl;
  launch_slam_rows(st, sp);
}
// Range / sun-sensor rows pending for this constructUpdate (vio_updater.cpp:352-403).  Returns the number of rows.
static int sensor_row_count(const xb_filter* f, int which, int ns) {
  if (which != 0) return 0;
  return ((f->sens_range_on && ns > 0) ? 1 : 0) + (f->sens_sun_on ? 2 : 0);
}
static void launch_sensor_rows_on(xb_filter* f, cudaStream_t st, int ns) {
  SensorParams sp{};
  sp.xv = f->d_xw; sp.M = f->M; sp.N = f->N; sp.n_poses = f->n_poses; sp.P = f->d_Pw; sp.anchor = f->d_anchor;
  sp.range_on = f->sens_range_on && ns > 0;
  sp.sun_on = f->sens_sun_on;
  if (!sp.range_on && !sp.sun_on) return;
  const int nw = (sp.range_on ? 1 : 0) + (sp.sun_on ? 2 : 0);
  // rows the reference stacks (every track keeps its 2L - 3 rows, gated out or not: msckf_update.cpp:50-52); more than
  // N + 1 of them -> QR compression, after which EVERY row is weighted sigma_img^2 (vio_updater.cpp:490-508)
  const long rows_total = 2L * f->l_msckf.n_obs - 3L * f->l_msckf.n + 2L * f->l_newms.n_obs - 3L * f->l_newms.n + 2L * ns + nw;
  const bool qr = rows_total > (long)f->N + 1;
  const double var_sun = 10000 * 0.01777777777;  // solar_update.cpp:48
  sp.range = f->sens_range.range; sp.pt_x = f->sens_range.img_pt_n[0]; sp.pt_y = f->sens_range.img_pt_n[1];
  for (int j = 0; j < 3; ++j) sp.tri[j] = f->sens_range.tr_feat_ids[j];
  sp.var_range = f->cfg.sigma_range * f->cfg.sigma_range;
  static const double chi2_1 = xb_chi2_quantile(0.9, 1.0);  // range_update.cpp:249-250
  sp.chi2_1 = chi2_1;
  sp.w_range = qr ? 1.0 : f->cfg.sigma_img / f->cfg.sigma_range;
  sp.w_sun = qr ? 1.0 : f->cfg.sigma_img / std::sqrt(var_sun);
  sp.sun_x = f->sens_sun.x_angle; sp.sun_y = f->sens_sun.y_angle;
  sp.cols = f->d_wcols; sp.vals = f->d_wvals; sp.res = f->d_wres; sp.gamma = f->d_wgamma; sp.inlier = f->d_winl;
  launch_sensor_rows(st, sp);
}
// SLAM rows + everything of the tall buffer on their columns + the first s_pad/32 tile columns of its factorisation + Wsym
static void slam_phase(xb_filter* f, cudaStream_t st, const UpdateDims& d, int stage_build, int stage_chol, int share,
                       bool with_rows) {
  {
    StageTimer st_(f, stage_build, st);
    if (with_rows) {
      launch_slam_rows_on(f, st, d.nslam);
      launch_sensor_rows_on(f, st, d.nslam);
    }
    launch_build_slam_part(st, d, f->d_Pw, f->d_scols, f->d_svals, f->d_sres, f->corr_zero ? nullptr : f->d_corr,
                           f->cfg.sigma_img * f->cfg.sigma_img, f->d_omega, f->d_omega_inv, f->d_T);
  }
  StageTimer st_(f, stage_chol, st);
  tallchol_range(st, f->d_T, d.m_pad, d.m_pad + d.n_pad + 96, d.m_pad, 0, d.s_pad, 1, f->d_flags, f->d_err, 0.0, nullptr, nullptr,
                 share);
  launch_wsym(st, d, f->d_omega_inv, f->d_T, f->d_Bc);
}
static size_t tall_bytes(const UpdateDims& d) { return sizeof(double) * (size_t)(d.m_pad + d.n_pad + 96) * d.m_pad; }

// sym(P) - W1s W1s^T (the SLAM-column part of the covariance downdate, 68 % of its flops at cfg-2) into a spare covariance
// buffer on a second side stream as soon as the SLAM columns are factored (ev_side); apply finishes it in place
static int early_downdate(xb_filter* f, const UpdateDims& d) {
  const int nt64 = (f->N + 63) / 64;
  const bool dd_early = (f->cfg.iekf_iter <= 1 || f->cfg.multi_uav) && f->cfg.downdate_precision == 0 &&
                        nt64 * (nt64 + 1) / 2 < 296 && !getenv("XB_NO_EARLY_DOWNDATE");
  if (!dd_early) return 0;
  f->d_dd = f->d_spare;
  CK(cudaStreamWaitEvent(f->side2, f->ev_side, 0));
  StageTimer st_(f, ST_SIDE_DD, f->side2);
  downdate_f64_range(f->side2, f->d_Pw, f->d_dd, f->N, f->d_T, d.m_pad, d.n_pad, 0, d.s_pad, 1, 0, f->d_omega_inv, f->d_Zb,
                     f->d_Yb, f->d_Qb);
  CK(cudaEventRecord(f->ev_dd, f->side2));
  f->dd_pending = true;
  f->dd_cols = d.s_pad;
  return 0;
}

// Gram stage of constructUpdate: G = [J|r]^T[J|r] - [B|b]^T[B|b] (+ D^T D), then its guarded Cholesky factor
static void run_gram(xb_filter* f, const GramParams& gp) {
  { StageTimer st_(f, ST_GRAM); launch_gram(f->stream, gp, f->overlap ? f->side3 : nullptr, f->ev_g0, f->ev_g1); }
  StageTimer st_(f, ST_CHOLG);
  tallchol_range(f->stream, f->d_Tg, f->gcols_pad, f->grows_pad, f->gcols_pad, 0, f->gcols_pad, 0, f->d_flags_g, f->d_err, 1e-14,
                 f->d_diag0, nullptr, f->side_pending ? f->chol_share : 1);
  // Rg = L_G^T is only materialised for the CUDA-core GEMM fallback (and on demand for xb_debug_read("Rg"))
  if (!gemm_uses_tensor_cores()) transpose(f->stream, f->d_Tg, f->d_Rg, f->gcols_pad, f->gcols_pad);
}

extern "C" int xb_vio_construct_update(xb_filter* f, int which) {
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  materialize(f);
  f->gram_deferred = false;
  const int M = f->M;
  const ListDev& l0 = which == 0 ? f->l_msckf : f->l_short;
  const int n0 = l0.n, n1 = which == 0 ? f->l_newms.n : 0, ns = which == 0 ? f->l_slam.n : 0;
  if (ns > f->n_features) return fail(XB_E_INVALID, "more SLAM tracks than SLAM features in the state");
  if (f->side_pending || f->dd_pending) {  // a constructed update that was never applied: its side work must not outlive this call
    int rci = invalidate_early(f);
    if (rci) return rci;
  }
  const int nw = sensor_row_count(f, which, ns);
  if (nw > 0 && f->cfg.sigma_range <= 0.0 && f->sens_range_on) return fail(XB_E_INVALID, "range measurement with sigma_range <= 0");
  f->slam_part_done = false;
  f->last_which = which;
  f->last_nslam = ns;
  f->last_nw = nw;
  f->constructed_any = (n0 + n1 + ns + nw) > 0;
  if (!f->constructed_any) return XB_OK;
  if (f->n_poses < 1) return fail(XB_E_INVALID, "empty pose window");
  const size_t gbytes = sizeof(double) * (size_t)f->grows_pad * f->gcols_pad;
  f->mm_G = 0;
  MmParams mp{};
  if (f->cfg.multi_uav && n0 > 0 && !f->mm_matches.empty()) {
    int rcm = mm_prepare(f, which, l0, mp);
    if (rcm < 0) return rcm;
  }
  if (ns + nw > 0) {
    // The SLAM rows and everything of the Kalman update that lives on their columns need only P and the estimates: they
    // are forked onto the side stream here and run next to the MSCKF track pipeline below.  Not with MSCKF-MSCKF matches:
    // their CI corrections change P between construct and apply (updater.cpp:84-97).
    const bool early = f->overlap && f->mm_G == 0 && !(f->cfg.multi_uav && which == 1) && f->asym_clones <= 1;
    CK(cudaMemcpyAsync(f->d_anchor, f->anchor.data(), sizeof(int) * f->F, cudaMemcpyHostToDevice, f->stream));
    if (early) {
      const UpdateDims d = update_dims(f, ns, nw);
      int rc0 = set_omega(f);
      if (rc0) return rc0;
      CK(cudaMemsetAsync(f->d_T, 0, tall_bytes(d), f->stream));
      CK(cudaEventRecord(f->ev_fork, f->stream));
      CK(cudaStreamWaitEvent(f->side, f->ev_fork, 0));
      slam_phase(f, f->side, d, ST_SIDE_SLAM, ST_SIDE_CHOL, f->chol_share, true);
      CK(cudaEventRecord(f->ev_side, f->side));
      if (early_downdate(f, d)) return XB_E_CUDA;
      f->side_pending = true;
      f->slam_part_done = true;
      f->side_used_corr = !f->corr_zero;
    } else {
      // gates are part of constructUpdate (inlier masks are observable right after it); the tall-buffer part follows in apply
      StageTimer st_(f, ST_SLAMROWS);
      launch_slam_rows_on(f, f->stream, ns);
      launch_sensor_rows_on(f, f->stream, ns);
    }
    if (which == 0) f->sens_range_on = f->sens_sun_on = false;  // "don't reuse measurement" (vio_updater.cpp:381, 402)
  }
  if (n0 + n1 > 0) {
    {
      StageTimer st_(f, ST_TRACKS);
      TrackParams tp0 = track_params(f, l0, 0);
      if (f->mm_G > 0) {
        StageTimer st2_(f, ST_MM_TRI);
        launch_mm_triangulate(f->stream, mp);
        tp0.mm_grp = f->d_mm_trkgrp; tp0.mm_ivd = f->d_mm_ivd; tp0.mm_F0 = f->d_mm_F0;
      }
      if (n0 > 0 && launch_tracks(f->stream, tp0)) return fail(XB_E_CAPACITY, "track kernel shared memory");
      if (f->mm_G > 0) {
        StageTimer st2_(f, ST_MM_CON);
        if (launch_mm_construct(f->stream, mp)) return fail(XB_E_CAPACITY, "multi-MSCKF kernel shared memory");
      }
      if (n1 > 0 && launch_tracks(f->stream, track_params(f, f->l_newms, 1))) return fail(XB_E_CAPACITY, "track kernel shared memory");
    }
    f->mm_last_G[which] = f->mm_G;
    if (f->cfg.multi_uav && which == 1) return XB_OK;  // the stacked short-MSCKF (h, res) is built and dropped (updater.cpp:58-70)
    GramParams gp{};
    gp.M = M; gp.n_poses = f->n_poses;
    gp.B = f->d_B0; gp.rowsB = 3 * n0; gp.nzB = f->nz; gp.partB = f->d_partB;
    gp.D = f->d_D1; gp.rowsD = 2 * f->l_newms.n_obs * (n1 > 0); gp.nzD = f->nz; gp.partD = f->d_partD;
    gp.off = l0.d_off; gp.inlier = f->d_inl0; gp.n_tracks_msckf = n0; gp.Jout = f->d_J0;
    gp.blocks = f->d_blocks;
    gp.T = f->d_Tg; gp.ld = f->gcols_pad; gp.rows_pad = f->grows_pad; gp.cols_pad = f->gcols_pad; gp.diag0 = f->d_diag0;
    static const bool no_defer = getenv("XB_NO_MM_DEFER") != nullptr;
    if (f->cfg.multi_uav && f->mm_last_G[which] > 0 && f->overlap && ns + nw > 0 && f->asym_clones <= 1 && !no_defer) {
      f->gram_gp = gp;
      f->gram_deferred = true;
      return XB_OK;
    }
    run_gram(f, gp);
  } else {
    CK(cudaMemsetAsync(f->d_Tg, 0, gbytes, f->stream));
    if (!gemm_uses_tensor_cores())
      CK(cudaMemsetAsync(f->d_Rg, 0, sizeof(double) * (size_t)f->gcols_pad * f->gcols_pad, f->stream));
  }
  return XB_OK;
}

static UpdateDims update_dims(const xb_filter* f, int nslam, int nw) {
  UpdateDims d;
  d.M = f->M; d.F = f->F; d.N = f->N;
  d.ms = 6 * f->M;
  d.nslam = nslam;
  d.nw = nw;
  d.wcols = f->d_wcols; d.wvals = f->d_wvals; d.wres = f->d_wres;
  d.ns2 = 2 * nslam + nw;
  d.s_pad = pad32(d.ns2);
  d.ro = d.s_pad;
  d.m = d.ms + d.ns2;
  d.m_pad = d.s_pad + pad32(d.ms);
  d.n_pad = pad32(f->N);
  d.ld = d.m_pad;
  return d;
}

// Omega = 15 core states + the 6 states of the newest clone (where P is not symmetric between updates)
static int set_omega(xb_filter* f) {
  const int slot = std::max(0, f->n_poses - 1);
  if (slot == f->omega_slot) return 0;
  const int n_pad = pad32(f->N);
  launch_set_omega(f->stream, f->d_omega, f->d_omega_inv, f->d_tileflag, slot, f->M, n_pad, n_pad / 32 + 4);
  f->omega_slot = slot;
  return 0;
}

// Exact applyUpdate for a work covariance that is unsymmetric beyond core + newest clone (k_general.cu): only reachable
// after updates without measurement rows.  Hd (m x N), res (m) and rdiag (m, or nullptr for sigma_img^2) on the device.
static int apply_general(xb_filter* f, int m, const double* Hd, const double* res, const double* rdiag, int cov_update,
                         const double* corr) {
  const int N = f->N;
  if ((size_t)m * (m + N + 1) > f->T_doubles || m > N) return fail(XB_E_CAPACITY, "general update: too many rows");
  double* X = (f->d_Pw == f->d_WA) ? f->d_WB : f->d_WA;
  {
    StageTimer st_(f, ST_TALLCHOL);
    general_update(f->stream, N, m, Hd, res, rdiag, f->cfg.sigma_img * f->cfg.sigma_img, corr, f->d_Pw, X, f->d_T, f->d_delta,
                   cov_update);
  }
  {
    StageTimer st_(f, ST_CORRECT);
    launch_apply_delta(f->stream, f->M, f->F, N, f->d_delta, f->d_xw, f->d_corr);
  }
  CK(cudaEventRecord(f->ev_corr, f->stream));
  f->xw_final = true;
  if (cov_update) f->asym_clones = 0;
  return XB_OK;
}

// chol_from > 0: the tile columns [0, chol_from) of the tall buffer are already factored (side stream); finish the rest
static int apply_from_tall(xb_filter* f, int m_pad, int n_pad, int cov_update, double* corr_total, int chol_from = 0) {
  const int N = f->N;
  { StageTimer st_(f, ST_TALLCHOL);
    // chol_from > 0: the leading columns are factored and folded into the rest (Schur complement): the remaining
    // columns are a plain tall factorisation of the sub-buffer starting at (chol_from, chol_from)
    tallchol(f->stream, f->d_T + (size_t)chol_from * m_pad + chol_from, m_pad, m_pad - chol_from + n_pad + 96, m_pad - chol_from,
             f->d_flags, f->d_err, 0.0, nullptr, f->d_trace);
    f->trace_tiles = (m_pad / 32) * (m_pad / 32 + 1) / 2 + ((n_pad + 96) / 32) * (m_pad / 32); }
  // The slab-column part of the downdate needs only W1 (complete once the factorisation is): it runs on a side stream next
  // to the Woodbury / State::correct stage; the Woodbury / Omega tail (which needs that stage's Z, Y, Q) follows as a
  // rank-42 pass of the same kernel with no main slabs.
  const bool dd_split = cov_update && f->overlap && f->dd_pending && f->dd_cols == chol_from && chol_from > 0;
  if (dd_split) {
    CK(cudaEventRecord(f->ev_b0, f->stream));
    CK(cudaStreamWaitEvent(f->side, f->ev_b0, 0));
    CK(cudaStreamWaitEvent(f->side, f->ev_dd, 0));
    downdate_f64_range(f->side, f->d_dd, f->d_dd, N, f->d_T, m_pad, n_pad, chol_from, m_pad, 0, 0, f->d_omega_inv, f->d_Zb,
                       f->d_Yb, f->d_Qb);
    CK(cudaEventRecord(f->ev_b1, f->side));
  }
  {
    StageTimer st_(f, ST_CORRECT);
    launch_correct(f->stream, f->M, f->F, N, f->d_T, m_pad, n_pad, f->d_Pw, f->d_omega, f->d_omega_inv, f->d_om, f->d_Zb,
                   f->d_Yb, f->d_Qb, f->d_Cb, f->d_xw, corr_total, f->d_delta, f->d_err);
  }
  CK(cudaEventRecord(f->ev_corr, f->stream));
  f->xw_final = true;
  if (cov_update) {
    StageTimer st_(f, ST_DOWNDATE);
    if (dd_split) {
      CK(cudaStreamWaitEvent(f->stream, f->ev_b1, 0));
      downdate_f64_range(f->stream, f->d_dd, f->d_dd, N, f->d_T, m_pad, n_pad, m_pad, m_pad, 0, 1, f->d_omega_inv, f->d_Zb,
                         f->d_Yb, f->d_Qb);
      f->d_Pw = f->d_dd;
    } else if (f->dd_pending && f->dd_cols == chol_from && chol_from > 0) {
      // the SLAM-column part is already in d_WA (side2): finish with the slab columns and the Woodbury / Omega terms
      CK(cudaStreamWaitEvent(f->stream, f->ev_dd, 0));
      downdate_f64_range(f->stream, f->d_dd, f->d_dd, N, f->d_T, m_pad, n_pad, chol_from, m_pad, 0, 1, f->d_omega_inv, f->d_Zb,
                         f->d_Yb, f->d_Qb);
      f->d_Pw = f->d_dd;
    } else if (f->cfg.downdate_precision == 1)
      downdate_tc(f->stream, f->d_Pw, N, f->d_T, m_pad, n_pad, f->d_omega_inv, f->d_Zb, f->d_Yb, f->d_Qb, f->d_tcws);
    else
      downdate_f64(f->stream, f->d_Pw, N, f->d_T, m_pad, n_pad, f->d_omega_inv, f->d_Zb, f->d_Yb, f->d_Qb);
  }
  if (f->dd_pending) {  // not consumed (cov_update == 0): later work on the main stream must not overtake side2
    CK(cudaStreamWaitEvent(f->stream, f->ev_dd, 0));
    f->dd_pending = false;
  }
  if (cov_update) f->asym_clones = 0;  // P = (P + P^T)/2 (updater.cpp:133)
  return XB_OK;
}

extern "C" int xb_updater_reset_correction(xb_filter* f) {
  if (f->slam_part_done && f->side_used_corr && invalidate_early(f)) return XB_E_CUDA;
  CK(cudaMemsetAsync(f->d_corr, 0, sizeof(double) * f->N, f->stream));
  f->corr_zero = true;
  return XB_OK;
}

extern "C" int xb_updater_apply_constructed(xb_filter* f, int cov_update) {
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  materialize(f);
  if (!f->constructed_any) return XB_OK;  // h.size() == 0 (updater.cpp:106)
  if (f->gram_deferred) {
    run_gram(f, f->gram_gp);
    f->gram_deferred = false;
  }
  const UpdateDims d = update_dims(f, f->last_nslam, f->last_nw);
  const double var = f->cfg.sigma_img * f->cfg.sigma_img;
  const double* corr = f->corr_zero ? nullptr : f->d_corr;
  if (f->asym_clones > 1) {
    // unsymmetric beyond core + newest clone (the previous updates applied no measurement): exact dense path
    if (invalidate_early(f)) return XB_E_CUDA;
    double* Hd = f->d_Hdense;
    double* res = Hd + (size_t)d.m * f->N;
    launch_gen_densify(f->stream, d, f->d_scols, f->d_svals, f->d_sres, f->d_Tg, f->gcols_pad,
                       f->d_Tg + (size_t)f->gcols_pad * f->gcols_pad, Hd, res);
    f->slam_part_done = false;
    f->corr_zero = false;
    return apply_general(f, d.m, Hd, res, nullptr, cov_update, corr);
  }
  int rc0 = set_omega(f);
  if (rc0) return rc0;
  int chol_from = 0;
  {
    StageTimer st_(f, ST_BUILD);
    const double* zg = f->d_Tg + (size_t)f->gcols_pad * f->gcols_pad;
    const double* Rg = gemm_uses_tensor_cores() ? nullptr : f->d_Rg;
    if (f->slam_part_done) {
      // join: SLAM columns of the tall buffer are built and factored by the side stream
      if (f->side_pending) CK(cudaStreamWaitEvent(f->stream, f->ev_side, 0));
      f->side_pending = false;
    } else {
      CK(cudaMemsetAsync(f->d_T, 0, tall_bytes(d), f->stream));
      if (d.ns2 > 0) slam_phase(f, f->stream, d, ST_SIDE_SLAM, ST_SIDE_CHOL, 1, false);
    }
    chol_from = d.s_pad;
    if (f->overlap) {
      // L21 and the Omega tile are independent of the P H_R^T -> S22 chain: side streams (idle by now), joined before Schur
      CK(cudaEventRecord(f->ev_b0, f->stream));
      CK(cudaStreamWaitEvent(f->side, f->ev_b0, 0));
      CK(cudaStreamWaitEvent(f->side3, f->ev_b0, 0));
      launch_slab_l21(f->side, d, Rg, f->d_Tg, f->gcols_pad, f->d_T, f->d_Bc);
      CK(cudaEventRecord(f->ev_b1, f->side));
      launch_slab_omega(f->side3, d, f->d_Pw, Rg, f->d_Tg, f->gcols_pad, f->d_omega, f->d_T, f->d_Gp);
      CK(cudaEventRecord(f->ev_b2, f->side3));
      launch_slab_s22(f->stream, d, f->d_Pw, Rg, f->d_Tg, f->gcols_pad, zg, corr, var, f->d_T);
      CK(cudaStreamWaitEvent(f->stream, f->ev_b1, 0));
      CK(cudaStreamWaitEvent(f->stream, f->ev_b2, 0));
      launch_slab_schur(f->stream, d, f->d_T);
    } else {
      launch_build_slab_part(f->stream, d, f->d_Pw, Rg, f->d_Tg, f->gcols_pad, zg, f->d_scols, f->d_svals, f->d_sres, corr, var,
                             f->d_omega, f->d_T, f->d_Bc, f->d_Gp);
    }
  }
  f->slam_part_done = false;
  f->corr_zero = false;
  return apply_from_tall(f, d.m_pad, d.n_pad, cov_update, f->d_corr, chol_from);
}

extern "C" int xb_updater_apply_update(xb_filter* f, const double* H, const double* res, const double* r_diag, int m,
                                       double* correction_total, int cov_update) {
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  materialize(f);
  const int N = f->N;
  if (m <= 0) return XB_OK;
  if (invalidate_early(f)) return XB_E_CUDA;
  const int m_pad = pad32(m), n_pad = pad32(N);
  if ((size_t)(m_pad + n_pad + 96) * m_pad > f->T_doubles || (size_t)m * N + 2 * (size_t)m > f->Hdense_doubles)
    return fail(XB_E_CAPACITY, "dense update has too many rows (max N after QR compression)");
  double* dH = f->d_Hdense;
  double* dres = dH + (size_t)m * N;
  double* drd = dres + m;
  CK(cudaMemcpyAsync(dH, H, sizeof(double) * (size_t)m * N, cudaMemcpyHostToDevice, f->stream));
  CK(cudaMemcpyAsync(dres, res, sizeof(double) * m, cudaMemcpyHostToDevice, f->stream));
  CK(cudaMemcpyAsync(drd, r_diag, sizeof(double) * m, cudaMemcpyHostToDevice, f->stream));
  if (correction_total) CK(cudaMemcpyAsync(f->d_corr, correction_total, sizeof(double) * N, cudaMemcpyHostToDevice, f->stream));
  else CK(cudaMemsetAsync(f->d_corr, 0, sizeof(double) * N, f->stream));
  int rc;
  if (f->asym_clones > 1) {
    rc = apply_general(f, m, dH, dres, drd, cov_update, f->d_corr);
  } else {
    int rc0 = set_omega(f);
    if (rc0) return rc0;
    CK(cudaMemsetAsync(f->d_T, 0, sizeof(double) * (size_t)(m_pad + n_pad + 96) * m_pad, f->stream));
    launch_dense_prepare(f->stream, m, m_pad, N, n_pad, f->d_Pw, dH, dres, drd, f->d_corr, f->d_omega, f->d_T);
    rc = apply_from_tall(f, m_pad, n_pad, cov_update, f->d_corr);
  }
  if (rc) return rc;
  if (correction_total) CK(cudaMemcpyAsync(correction_total, f->d_corr, sizeof(double) * N, cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return XB_OK;
}

// Updater::applyCI (updater.cpp:144-161): K = P_j H^T S^-1; delta = K r; P = (I - K H) P_j; symmetrise; correct.
// S is the caller's (CI-fused) innovation covariance; it is tiny (3k x 3k), so S^-1 is formed on the host by
// partial-pivoting Gauss-Jordan exactly like the reference's S.inverse(); everything N-sized runs on the device.
extern "C" int xb_updater_apply_ci(xb_filter* f, const double* H, const double* res, const double* S, int m,
                                   const int* scaled_block_cols, int n_blocks, double w_result) {
  if (!f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  materialize(f);
  const int N = f->N;
  if (m <= 0 || m > 96) return fail(XB_E_INVALID, "applyCI: 0 < rows <= 96");
  if (invalidate_early(f)) return XB_E_CUDA;
  std::vector<double> A((size_t)m * 2 * m, 0.0);
  for (int r = 0; r < m; ++r) {
    for (int c = 0; c < m; ++c) A[(size_t)r * 2 * m + c] = S[(size_t)r * m + c];
    A[(size_t)r * 2 * m + m + r] = 1.0;
  }
  for (int c = 0; c < m; ++c) {
    int best = c;
    for (int r = c + 1; r < m; ++r)
      if (std::fabs(A[(size_t)r * 2 * m + c]) > std::fabs(A[(size_t)best * 2 * m + c])) best = r;
    if (A[(size_t)best * 2 * m + c] == 0.0) return fail(XB_E_RUNTIME, "applyCI: singular S");
    if (best != c)
      for (int k = 0; k < 2 * m; ++k) std::swap(A[(size_t)c * 2 * m + k], A[(size_t)best * 2 * m + k]);
    const double d = A[(size_t)c * 2 * m + c];
    for (int k = 0; k < 2 * m; ++k) A[(size_t)c * 2 * m + k] /= d;
    for (int r = 0; r < m; ++r)
      if (r != c) {
        const double fct = A[(size_t)r * 2 * m + c];
        if (fct != 0.0)
          for (int k = 0; k < 2 * m; ++k) A[(size_t)r * 2 * m + k] -= fct * A[(size_t)c * 2 * m + k];
      }
  }
  std::vector<double> Sinv((size_t)m * m);
  for (int r = 0; r < m; ++r)
    for (int c = 0; c < m; ++c) Sinv[(size_t)r * m + c] = A[(size_t)r * 2 * m + m + c];
  double* dH = f->d_Hdense;
  double* dres = dH + (size_t)m * N;
  int* dcols = (int*)(dres + m + 2);
  double* A1 = f->d_T;                   // N x m
  double* Kd = A1 + (size_t)N * m;       // N x m
  double* HP = Kd + (size_t)N * m;       // m x N
  double* dSinv = HP + (size_t)N * m;    // m x m
  CK(cudaMemcpyAsync(dH, H, sizeof(double) * (size_t)m * N, cudaMemcpyHostToDevice, f->stream));
  CK(cudaMemcpyAsync(dres, res, sizeof(double) * m, cudaMemcpyHostToDevice, f->stream));
  CK(cudaMemcpyAsync(dSinv, Sinv.data(), sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice, f->stream));
  if (n_blocks > 0) {
    if (n_blocks > 256) return fail(XB_E_INVALID, "too many scaled blocks");
    CK(cudaMemcpyAsync(dcols, scaled_block_cols, sizeof(int) * n_blocks, cudaMemcpyHostToDevice, f->stream));
    launch_scale_blocks(f->stream, f->d_Pw, N, dcols, n_blocks, w_result);  // P_j (only diagonal 3x3 blocks)
  }
  gemm_nt(f->stream, N, m, N, 1.0, f->d_Pw, N, dH, N, 0.0, A1, m);    // P_j H^T
  gemm_nn(f->stream, N, m, m, 1.0, A1, m, dSinv, m, 0.0, Kd, m);      // K
  gemm_nn(f->stream, m, N, N, 1.0, dH, N, f->d_Pw, N, 0.0, HP, N);    // H P_j
  gemv(f->stream, N, m, Kd, m, dres, f->d_delta);                      // delta = K r
  launch_apply_delta(f->stream, f->M, f->F, N, f->d_delta, f->d_xw, nullptr);
  launch_ci_cov(f->stream, f->d_Pw, N, Kd, HP, m);
  CK(cudaStreamSynchronize(f->stream));  // host staging lifetime
  return XB_OK;
}

// ---- VioUpdater::postUpdate (vio_updater.cpp:425-449) -----------------------------------------------------------------
extern "C" int xb_vio_post_update(xb_filter* f) {
  const int M = f->M, F = f->F, N = f->N;
  const double var = f->cfg.sigma_img * f->cfg.sigma_img;
  materialize(f);
  StageTimer st_(f, ST_POST);
  if (f->l_newms.n > 0 || f->l_newstd.n > 0) f->xw_final = false;
  if (f->l_newms.n > 0) {
    const int n_new = f->l_newms.n;
    if (f->n_features + n_new > F) return fail(XB_E_CAPACITY, "no free SLAM feature slot (state_manager.cpp:208)");
    FeatInitParams fp{};
    fp.M = M; fp.F = F; fp.N = N; fp.n_poses = f->n_poses; fp.n_features = f->n_features; fp.n_new = n_new;
    fp.H1 = f->d_H1; fp.H2 = f->d_H2; fp.ivd = f->d_ivd1; fp.corr = f->d_corr; fp.var_img = var;
    launch_init_msckf_slam(f->stream, fp, f->d_xw, f->d_Pw, f->d_fscratch);
    for (int i = 0; i < n_new; ++i) f->anchor[f->n_features + i] = f->n_poses - 1;
    f->n_features += n_new;
  }
  if (f->l_newstd.n > 0) {
    const int n_new = f->l_newstd.n;
    if (f->n_features + n_new > F) return fail(XB_E_CAPACITY, "no free SLAM feature slot (state_manager.cpp:208)");
    launch_init_std_slam(f->stream, M, F, N, f->n_features, n_new, f->l_newstd.d_off, f->l_newstd.d_obs, f->cfg.rho_0, var,
                         f->cfg.sigma_rho_0 * f->cfg.sigma_rho_0, f->d_xw, f->d_Pw);
    for (int i = 0; i < n_new; ++i) f->anchor[f->n_features + i] = f->n_poses - 1;
    f->n_features += n_new;
  }
  return XB_OK;
}

// ---- Updater::update (updater.cpp:39-115, single-UAV build) -------------------------------------------------------------
extern "C" int xb_updater_update(xb_filter* f) {
  int rc;
  if ((rc = xb_updater_reset_correction(f)) < 0) return rc;
  if (f->cfg.multi_uav) {  // the -DMULTI_UAV build of the same method (updater.cpp:58-70, 84-97)
    f->mm_last_G[0] = f->mm_last_G[1] = 0;
    if (f->l_short.n > 0) {
      if ((rc = xb_vio_construct_update(f, 1)) < 0) return rc;
      if ((rc = xb_updater_apply_ci_lists(f)) < 0) return rc;  // only the CI lists: no applyUpdate in this build
    }
    if ((rc = xb_sm_manage(f, f->lost.data(), (int)f->lost.size())) < 0) return rc;
    if (f->l_msckf.n || f->l_slam.n || f->l_newstd.n || f->l_newms.n) {
      if ((rc = xb_updater_reset_correction(f)) < 0) return rc;
      if ((rc = xb_vio_construct_update(f, 0)) < 0) return rc;
      if ((rc = xb_updater_apply_ci_lists(f)) < 0) return rc;
      if ((rc = xb_updater_apply_constructed(f, 1)) < 0) return rc;
      if ((rc = xb_vio_post_update(f)) < 0) return rc;
    }
    f->mm_matches.clear();  // preProcess replaces msckf_matches_ on every update (vio_updater.cpp:185)
    f->mm_obs_h.clear();
    return XB_OK;
  }
  if (f->l_short.n > 0) {  // preUpdateShortMsckf (vio_updater.cpp:209-215)
    if ((rc = xb_vio_construct_update(f, 1)) < 0) return rc;
    if ((rc = xb_updater_apply_constructed(f, 1)) < 0) return rc;
  }
  if ((rc = xb_sm_manage(f, f->lost.data(), (int)f->lost.size())) < 0) return rc;  // preUpdate (vio_updater.cpp:200-207)
  const bool requested = f->l_msckf.n || f->l_slam.n || f->l_newstd.n || f->l_newms.n;
  if (requested) {
    if ((rc = xb_updater_reset_correction(f)) < 0) return rc;
    const int iters = f->cfg.iekf_iter;  // zero iterations when iekf_iter_ = 0, like the reference's loop (updater.cpp:99)
    for (int i = 0; i < iters; ++i) {
      if ((rc = xb_vio_construct_update(f, 0)) < 0) return rc;
      if ((rc = xb_updater_apply_constructed(f, i == iters - 1)) < 0) return rc;
    }
    if ((rc = xb_vio_post_update(f)) < 0) return rc;
  }
  return XB_OK;
}

extern "C" int xb_propagate(xb_filter* f, int slot_from, int slot_to) {
  if (slot_to != (slot_from + 1) % f->NS) return fail(XB_E_INVALID, "slot_to must follow slot_from in the ring");
  ImuSample none{};
  propagate_chain(f, slot_from, 1, none);
  f->slot_gen[slot_to] = f->slot_gen[slot_from];
  f->slot_asym[slot_to] = f->slot_asym[slot_from];
  return XB_OK;
}

static int repropagate_from(xb_filter* f, int idx) {  // ekf.cpp:227-255
  int n = 0;
  for (int c = idx; c != f->tail; c = next_idx(f, c)) ++n;
  StageTimer st_(f, ST_PROPAGATE);
  ImuSample none{};
  propagate_chain(f, idx, n, none);
  for (int c = idx, k = 0; k < n; ++k) {
    c = next_idx(f, c);
    f->slot_gen[c] = f->slot_gen[idx];
    f->slot_asym[c] = f->slot_asym[idx];
    f->slot_serial[c] = ++f->serial;
  }
  return n;
}

// ---- Ekf::processUpdateMeasurement (ekf.cpp:179-213) ---------------------------------------------------------------------
// The two halves of Ekf::processUpdateMeasurement around Updater::update, for a host-side template method that drives
// the stages itself (include/x/xb200_binding.hpp): begin = closestIdx + copy of the buffered state into the work state
// (ekf.cpp:184-196), end = write-back + re-propagation of the newer states (ekf.cpp:201-205, 227-255).
extern "C" int xb_ekf_update_begin(xb_filter* f, double timestamp, double* xvec_out) {
  if (!f) return fail(XB_E_INVALID, "null filter");
  if (f->status == 0) return 0;
  const int idx = closest_idx(f, timestamp);
  if (idx < 0) return 0;
  if (f->slot_gen[idx] < 0) return fail(XB_E_STALE, "the buffered state's covariance generation was recycled: raise xb_config.n_generations");
  int rc;
  if ((rc = xb_work_load(f, idx)) < 0) return rc;
  f->last_update_slot = idx;
  f->xw_final = false;
  if (xvec_out) {
    CK(cudaMemcpyAsync(xvec_out, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToHost, f->stream));
    CK(cudaStreamSynchronize(f->stream));
  }
  return 1;
}
static int update_end(xb_filter* f, int idx, double* xvec_out);
extern "C" int xb_ekf_update_end(xb_filter* f, double* xvec_out) {
  if (!f || f->last_update_slot < 0) return fail(XB_E_INVALID, "no update in progress");
  return update_end(f, f->last_update_slot, xvec_out);
}
extern "C" int xb_ekf_process_update(xb_filter* f, double* xvec_out) {
  int rc = xb_ekf_update_begin(f, f ? f->meas_time : 0.0, nullptr);
  if (rc <= 0) return rc;
  if ((rc = xb_updater_update(f)) < 0) return rc;
  return update_end(f, f->last_update_slot, xvec_out);
}
static int update_end(xb_filter* f, int idx, double* xvec_out) {
  int rc;
  int n_re = 0;
  for (int c = idx; c != f->tail; c = next_idx(f, c)) ++n_re;
  if (f->overlap && f->xw_final && n_re > 0 && n_re <= 128) {
    // The estimates have been final since ev_corr (no feature initialisation followed), so the means of the re-propagation
    // (a serial quaternion/velocity chain on one CTA) run on a side stream next to the covariance downdate; the strips
    // wait for both.
    CK(cudaStreamWaitEvent(f->side2, f->ev_corr, 0));
    {
      StageTimer st_(f, ST_SIDE_MEANS, f->side2);
      CK(cudaMemcpyAsync(f->d_xv + (size_t)idx * f->LX, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToDevice, f->side2));
      ImuSample none{};
      launch_prop_means(f->side2, f->d_xv, f->LX, f->NS, idx, n_re, none, prop_params(f), f->d_FQ2);
    }
    CK(cudaEventRecord(f->ev_means, f->side2));
    if ((rc = work_store_impl(f, idx, false)) < 0) return rc;
    CK(cudaStreamWaitEvent(f->stream, f->ev_means, 0));
    {
      StageTimer st_(f, ST_PROPAGATE);
      launch_prop_strips(f->stream, f->d_strip, f->N, f->NS, idx, n_re, f->d_FQ2);
      if (f->slot_asym[idx] > 0) launch_prop_strips(f->stream, f->d_strip2, f->N, f->NS, idx, n_re, f->d_FQ2, 1);
    }
    for (int c = idx, k = 0; k < n_re; ++k) {
      c = next_idx(f, c);
      f->slot_gen[c] = f->slot_gen[idx];
      f->slot_asym[c] = f->slot_asym[idx];
      f->slot_serial[c] = ++f->serial;
    }
  } else {
    if ((rc = xb_work_store(f, idx)) < 0) return rc;
    repropagate_from(f, idx);
  }
  if (xvec_out) {
    CK(cudaMemcpyAsync(xvec_out, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToHost, f->stream));
    if ((rc = xb_synchronize(f)) < 0) return rc;
  }
  return 1;
}

// ---- covariance-intersection fusion -----------------------------------------------------------------------------------
extern "C" int xb_ci_payload_len(const xb_filter* f) { return f->ci_payload_len; }

extern "C" int xb_ci_pack(xb_filter* f, int slot, double* dev_payload) {
  if (!f || !dev_payload) return fail(XB_E_INVALID, "null argument");
  if (slot < 0) slot = f->tail;
  if (slot >= f->NS || f->slot_gen[slot] < 0) return fail(XB_E_INVALID, "slot has no valid state");
  CK(cudaMemcpyAsync(f->d_anchor, f->anchor.data(), sizeof(int) * std::max(1, f->F), cudaMemcpyHostToDevice, f->stream));
  // only pose / feature columns are read, so the generation buffer (P_vv) is the slot's covariance here
  launch_ci_pack(f->stream, f->d_xv + (size_t)slot * f->LX, f->gen_buf[f->slot_gen[slot]], f->N, f->M,
                 f->F, f->n_poses, f->n_features, f->d_anchor, dev_payload);
  return XB_OK;
}

static int check_ci_weight(double w) {  // ci.cpp:98-101
  if (w > 1.0 || w == 0.0 || w < -1.0) return fail(XB_E_RUNTIME, "The CI weights must be lower than 1.0 and larger than 0.0");
  // w in [-1, 0) asks the reference for NLopt-optimised weights (ci.cpp:64-75,103-116).  There is no optimiser here:
  // the weights are the ones the reference itself falls back to when its optimiser fails -- the pair form uses -w
  // (ci.cpp:112-113), the k-agent form keeps w as written (ci.cpp:70-75) -- see ci_pair_weight / k_mm_construct.
  return 0;
}

// Updater::collaborativeUpdate (updater.cpp:22-36) on the work state, peers given as gathered payload slots
static int collaborative_update_packed(xb_filter* f, const double* dev_gathered, int n_agents, const xb_slam_match* matches,
                                       int n_matches) {
  if (n_matches <= 0) return XB_OK;  // preUpdateCI (vio_updater.cpp:76-79)
  if (invalidate_early(f)) return XB_E_CUDA;
  materialize(f);
  if (n_matches > f->ci_max_matches) return fail(XB_E_CAPACITY, "too many SLAM-SLAM matches");
  int rc = check_ci_weight(f->cfg.ci_slam_w);
  if (rc) return rc;
  std::vector<int> hm(3 * (size_t)n_matches);
  for (int j = 0; j < n_matches; ++j) {
    hm[3 * j] = matches[j].peer;
    hm[3 * j + 1] = matches[j].current_feature_id;
    hm[3 * j + 2] = matches[j].received_feature_id;
    if (matches[j].peer < 0 || matches[j].peer >= n_agents || matches[j].received_feature_id < 0 ||
        matches[j].received_feature_id >= f->F)
      return fail(XB_E_INVALID, "bad SLAM match");
    if (matches[j].current_feature_id < 0 || matches[j].current_feature_id >= f->n_features || f->anchor[matches[j].current_feature_id] < 0)
      return fail(XB_E_RUNTIME, "anchor_idx < 0");  // multi_slam_update.cpp:83-85
  }
  CK(cudaMemcpyAsync(f->d_ci_matches, hm.data(), sizeof(int) * hm.size(), cudaMemcpyHostToDevice, f->stream));
  CK(cudaMemcpyAsync(f->d_anchor, f->anchor.data(), sizeof(int) * std::max(1, f->F), cudaMemcpyHostToDevice, f->stream));
  const double var_lm = f->cfg.sigma_landmark * f->cfg.sigma_landmark;
  launch_ci_slam(f->stream, f->d_xw, f->d_Pw, f->N, f->M, f->F, f->n_poses, f->n_features, f->d_anchor, dev_gathered,
                 f->ci_payload_len, f->d_ci_matches, n_matches, var_lm, f->cfg.ci_slam_w < 0.0 ? -f->cfg.ci_slam_w : f->cfg.ci_slam_w,
                 xb_chi2_quantile(0.9, 3.0),
                 f->d_ci_rec, f->d_ci_last, f->d_ci_K, f->d_ci_delta, f->d_ci_HP);
  f->ci_last_n = n_matches;
  CK(cudaStreamSynchronize(f->stream));  // hm lifetime
  return XB_OK;
}

extern "C" int xb_ekf_process_others_packed(xb_filter* f, double t, const double* dev_gathered, int n_agents,
                                            const xb_slam_match* matches, int n_matches, double* xvec_out) {
  if (!f) return fail(XB_E_INVALID, "null filter");
  if (f->status == 0) return 0;
  if (n_agents > f->ci_max_agents) return fail(XB_E_CAPACITY, "too many agents");
  const int idx = closest_idx(f, t);
  if (idx < 0) return 0;
  if (f->slot_gen[idx] < 0) return fail(XB_E_STALE, "the buffered state's covariance generation was recycled: raise xb_config.n_generations");
  int rc;
  if ((rc = xb_work_load(f, idx)) < 0) return rc;
  f->last_update_slot = idx;
  if ((rc = collaborative_update_packed(f, dev_gathered, n_agents, matches, n_matches)) < 0) return rc;
  if ((rc = xb_work_store(f, idx)) < 0) return rc;
  repropagate_from(f, idx);
  if (xvec_out) {
    CK(cudaMemcpyAsync(xvec_out, f->d_xw, sizeof(double) * f->LX, cudaMemcpyDeviceToHost, f->stream));
    if ((rc = xb_synchronize(f)) < 0) return rc;
  }
  return 1;
}

// Reference-format entry: peers as SimpleState (full covariance).  The host only gathers, per match, the peer's 9x9
// covariance block and forms the 13-double payload entry; everything N-sized runs on the device.
static int peers_to_payload(xb_filter* f, const xb_peer_state* peers, int n_peers, const xb_slam_match* matches, int n_matches) {
  if (!f || (n_peers > 0 && !peers)) return fail(XB_E_INVALID, "null argument");
  if (n_peers > f->ci_max_agents) return fail(XB_E_CAPACITY, "too many peers");
  const int PL = f->ci_payload_len;
  std::vector<double> pay((size_t)std::max(1, n_peers) * PL, 0.0);
  for (int j = 0; j < n_matches; ++j) {
    const int pi = matches[j].peer, rf = matches[j].received_feature_id;
    if (pi < 0 || pi >= n_peers) return fail(XB_E_INVALID, "match refers to an unknown peer");
    const xb_peer_state& ps = peers[pi];
    if (rf < 0 || rf >= ps.n_features_max || rf >= f->F) return fail(XB_E_INVALID, "bad received feature id");
    const int Mp = ps.n_poses_max, Np = XB_NERR(Mp, ps.n_features_max);
    const int an = ps.anchor_idxs[rf];
    if (an < 0) return fail(XB_E_RUNTIME, "anchor_idx < 0");
    if (an >= Mp) return fail(XB_E_INVALID, "peer anchor index outside its pose window");
    const double a = ps.features[3 * rf], b = ps.features[3 * rf + 1], r = ps.features[3 * rf + 2];
    if (r == 0.0) return fail(XB_E_RUNTIME, "rho = 0");
    double Ra[9], sk[9], m3[9], t3[3], A1[9], A2[9], h9[27];
    xb_rot(ps.orientations + 4 * an, Ra);
    const double ab1[3] = {a, b, 1.0};
    xb_mv33(Ra, ab1, t3);
    double* o = pay.data() + (size_t)pi * PL + 8 + 13 * rf;
    o[0] = 1.0;
    for (int e = 0; e < 3; ++e) o[1 + e] = (1.0 / r) * t3[e] + (ps.positions[3 * an + e] + ps.translation[e]);  // simple_state.cpp:51-65
    xb_skew(ab1, sk);
    xb_mm33(Ra, sk, A1);
    xb_mat_ivd(a, b, r, m3);
    xb_mm33(Ra, m3, A2);
    int c9[9];
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < 3; ++c) {
        h9[i * 9 + c] = -((i == c) ? 1.0 : 0.0);          // other_h_j carries the opposite sign (multi_slam_update.cpp:180-203)
        h9[i * 9 + 3 + c] = (1.0 / r) * A1[i * 3 + c];
        h9[i * 9 + 6 + c] = -(1.0 / r) * A2[i * 3 + c];
      }
    for (int c = 0; c < 3; ++c) { c9[c] = XB_CORE + 3 * an + c; c9[3 + c] = XB_CORE + 3 * Mp + 3 * an + c; c9[6 + c] = XB_CORE + (2 * Mp + rf) * 3 + c; }
    auto Pat = [&](int rr, int cc) { return ps.cov_layout == XB_COL_MAJOR ? ps.cov[(size_t)cc * Np + rr] : ps.cov[(size_t)rr * Np + cc]; };
    double T[27];
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < 9; ++c) {
        double s = 0.0;
        for (int d = 0; d < 9; ++d) s += h9[i * 9 + d] * Pat(c9[d], c9[c]);
        T[i * 9 + c] = s;
      }
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int c = 0; c < 9; ++c) s += T[i * 9 + c] * h9[k * 9 + c];
        o[4 + i * 3 + k] = s;
      }
  }
  CK(cudaMemcpyAsync(f->d_ci_gather, pay.data(), sizeof(double) * pay.size(), cudaMemcpyHostToDevice, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return XB_OK;
}
extern "C" int xb_ekf_process_others(xb_filter* f, double t, const xb_peer_state* peers, int n_peers,
                                     const xb_slam_match* matches, int n_matches, double* xvec_out) {
  int rc = peers_to_payload(f, peers, n_peers, matches, n_matches);
  if (rc < 0) return rc;
  return xb_ekf_process_others_packed(f, t, f->d_ci_gather, std::max(1, n_peers), matches, n_matches, xvec_out);
}
// Updater::collaborativeUpdate (updater.cpp:22-36) on the WORK state (stage-level twin of xb_ekf_process_others)
extern "C" int xb_updater_collaborative_update(xb_filter* f, const xb_peer_state* peers, int n_peers,
                                               const xb_slam_match* matches, int n_matches) {
  if (!f || !f->d_Pw) return fail(XB_E_INVALID, "no work state loaded");
  int rc = peers_to_payload(f, peers, n_peers, matches, n_matches);
  if (rc < 0) return rc;
  return collaborative_update_packed(f, f->d_ci_gather, std::max(1, n_peers), matches, n_matches);
}
// Ring-buffer bookkeeping a host-side x::State needs to refer to a buffered covariance safely: time stamp and
// covariance generation of a slot (-1: the slot's covariance is gone), and the slot the last update was stored in.
extern "C" int xb_ekf_slot_info(const xb_filter* f, int slot, double* time_out, int* generation_out) {
  if (!f || slot < 0 || slot >= f->NS) return fail(XB_E_INVALID, "bad slot");
  if (time_out) *time_out = f->h_time[slot];
  if (generation_out) *generation_out = f->slot_gen[slot] < 0 ? -1 : f->slot_serial[slot];
  return XB_OK;
}
extern "C" int xb_ekf_last_update_slot(const xb_filter* f) { return f ? f->last_update_slot : -1; }

extern "C" int xb_ci_last_gates(xb_filter* f, double* out, int max_matches) {
  const int n = std::min(max_matches, f->ci_last_n);
  std::vector<double> rec(64 * (size_t)std::max(1, n));
  CK(cudaStreamSynchronize(f->stream));
  CK(cudaMemcpy(rec.data(), f->d_ci_rec, sizeof(double) * 64 * (size_t)n, cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; ++j) { out[2 * j] = rec[64 * (size_t)j]; out[2 * j + 1] = rec[64 * (size_t)j + 1]; }
  return n;
}

static int mm_store_matches(xb_filter* f, const xb_msckf_match* matches, int n_matches, int n_agents) {
  f->mm_matches.clear();
  f->mm_obs_h.clear();
  for (int j = 0; j < n_matches; ++j) {
    const xb_msckf_match& m = matches[j];
    if (m.peer < 0 || m.peer >= n_agents) return fail(XB_E_INVALID, "match refers to an unknown peer");
    if (m.n_obs < 1 || m.n_obs > f->M || !m.obs) return fail(XB_E_INVALID, "peer track length must be in [1, n_poses_max]");
    if (m.which != 0 && m.which != 1) return fail(XB_E_INVALID, "which must be 0 (msckf) or 1 (msckf_short)");
    f->mm_matches.push_back({m.peer, m.which, m.id_current_track, m.n_obs, f->mm_obs_h.size()});
    f->mm_obs_h.insert(f->mm_obs_h.end(), m.obs, m.obs + 2 * (size_t)m.n_obs);
  }
  return XB_OK;
}

extern "C" int xb_vio_set_msckf_matches_packed(xb_filter* f, const double* dev_gathered, int n_agents,
                                               const xb_msckf_match* matches, int n_matches) {
  if (!f || (n_matches > 0 && (!matches || !dev_gathered))) return fail(XB_E_INVALID, "null argument");
  if (n_matches > 0 && !f->cfg.multi_uav) return fail(XB_E_INVALID, "MSCKF-MSCKF matches need xb_config.multi_uav = 1");
  f->mm_gather = dev_gathered;
  return mm_store_matches(f, matches, std::max(0, n_matches), n_agents);
}

// Reference-format entry: peers as SimpleState.  The host cuts the pose payload (window + 6M x 6M covariance block)
// out of each snapshot; camera positions carry SimpleState::translation_ (simple_state.cpp:51-65).
extern "C" int xb_vio_set_msckf_matches(xb_filter* f, const xb_peer_state* peers, int n_peers,
                                        const xb_msckf_match* matches, int n_matches) {
  if (!f) return fail(XB_E_INVALID, "null filter");
  if (n_matches <= 0) { f->mm_matches.clear(); f->mm_obs_h.clear(); return XB_OK; }
  if (!peers || !matches) return fail(XB_E_INVALID, "null argument");
  if (!f->cfg.multi_uav) return fail(XB_E_INVALID, "MSCKF-MSCKF matches need xb_config.multi_uav = 1");
  if (n_peers > f->ci_max_agents) return fail(XB_E_CAPACITY, "too many peers");
  const int M = f->M, PL = f->mm_pp_len, n6 = 6 * M;
  std::vector<double> pay((size_t)n_peers * PL, 0.0);
  for (int p = 0; p < n_peers; ++p) {
    const xb_peer_state& ps = peers[p];
    if (ps.n_poses_max != M) return fail(XB_E_INVALID, "peers must use the same n_poses_max");
    if (!ps.positions || !ps.orientations || !ps.cov) return fail(XB_E_INVALID, "peer snapshot incomplete");
    const int Np = XB_NERR(ps.n_poses_max, ps.n_features_max);
    double* o = pay.data() + (size_t)p * PL;
    o[0] = 1.0; o[2] = M;
    for (int i = 0; i < 3 * M; ++i) o[8 + i] = ps.positions[i] + ps.translation[i % 3];
    for (int i = 0; i < 4 * M; ++i) o[8 + 3 * M + i] = ps.orientations[i];
    for (int r = 0; r < n6; ++r)
      for (int c = 0; c < n6; ++c)
        o[8 + 7 * M + (size_t)r * n6 + c] = ps.cov_layout == XB_COL_MAJOR ? ps.cov[(size_t)(XB_CORE + c) * Np + XB_CORE + r]
                                                                          : ps.cov[(size_t)(XB_CORE + r) * Np + XB_CORE + c];
  }
  CK(cudaMemcpyAsync(f->d_mm_gather, pay.data(), sizeof(double) * pay.size(), cudaMemcpyHostToDevice, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  f->mm_gather = f->d_mm_gather;
  return mm_store_matches(f, matches, n_matches, n_peers);
}

extern "C" int xb_ci_pose_payload_len(const xb_filter* f) { return f->mm_pp_len; }

extern "C" int xb_ci_pack_poses(xb_filter* f, int slot, double* dev_payload) {
  if (!f || !dev_payload) return fail(XB_E_INVALID, "null argument");
  if (slot < 0) slot = f->tail;
  if (slot >= f->NS || f->slot_gen[slot] < 0) return fail(XB_E_INVALID, "slot has no valid state");
  // only pose columns are read, so the generation buffer (P_vv) is the slot's covariance here
  launch_pack_poses(f->stream, f->d_xv + (size_t)slot * f->LX, f->gen_buf[f->slot_gen[slot]], f->N, f->M,
                    dev_payload);
  return XB_OK;
}

extern "C" int xb_mm_last_gates(xb_filter* f, int which, double* out, int max_groups) {
  if (which < 0 || which > 1) return fail(XB_E_INVALID, "which must be 0 or 1");
  const int n = std::min(max_groups, f->mm_last_G[which]);
  if (n <= 0) return 0;
  std::vector<double> r((size_t)XB_MM_REC * n);
  CK(cudaStreamSynchronize(f->stream));
  CK(cudaMemcpy(r.data(), f->d_mm_rec + (size_t)which * XB_MM_REC * f->mm_max_groups, sizeof(double) * r.size(), cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; ++j)
    for (int e = 0; e < 3; ++e) out[3 * j + e] = r[(size_t)XB_MM_REC * j + e];
  return n;
}

// ---- introspection ----------------------------------------------------------------------------------------------------------
// Test / measurement hook for the dense contraction kernels: C = beta C + alpha A B^T (A: M x K, B: N x K, row-major, host
// buffers).  op 0: gemm_nt as the update calls it (TMA-staged tiles when the shape qualifies), 1: the cp.async kernel,
// 2: symmetric downdate C <- (C + C^T)/2 - A A^T through the TMA kernel (M == N, B ignored).  Returns 1 when the TMA kernel
// ran (op 0, 2), 0 otherwise; ms_out (optional) receives the average device time of `reps` launches.
extern "C" int xb_debug_gemm(int op, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double alpha,
                             double beta, double* C, int ldc, int reps, double* ms_out) {
  double *dA = nullptr, *dB = nullptr, *dC = nullptr, *dC0 = nullptr;
  const size_t ba = sizeof(double) * (size_t)M * lda, bb = sizeof(double) * (size_t)N * ldb, bc = sizeof(double) * (size_t)M * ldc;
  if (cudaMalloc(&dA, ba) || cudaMalloc(&dB, std::max<size_t>(bb, 8)) || cudaMalloc(&dC, bc) || cudaMalloc(&dC0, bc))
    return fail(XB_E_CUDA, "xb_debug_gemm: allocation");
  CK(cudaMemcpy(dA, A, ba, cudaMemcpyHostToDevice));
  if (op != 2) CK(cudaMemcpy(dB, B, bb, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC0, C, bc, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int used_tma = 0;
  float ms_tot = 0.f;
  for (int r = 0; r < std::max(1, reps); ++r) {
    CK(cudaMemcpy(dC, dC0, bc, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0, 0));
    if (op == 0) {
      used_tma = gemm_nt_tma(0, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc) ? 1 : 0;
      if (!used_tma) gemm_nt_cpasync(0, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc);
    } else if (op == 1) {
      gemm_nt_cpasync(0, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc);
    } else {
      used_tma = downdate_sym_tma(0, M, K, dA, lda, dC, dC, ldc) ? 1 : 0;
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 || reps <= 1) ms_tot += ms;
  }
  if (ms_out) *ms_out = ms_tot / std::max(1, reps > 1 ? reps - 1 : 1);
  CK(cudaMemcpy(C, dC, bc, cudaMemcpyDeviceToHost));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dC0);
  return used_tma;
}

extern "C" int xb_debug_read(xb_filter* f, const char* name, double* out, int max_doubles) {
  const std::string n(name);
  const double* src = nullptr;
  size_t cnt = 0;
  const int W = 6 * f->M + 1;
  if (n == "gamma0") { src = f->d_gamma0; cnt = f->last_which ? f->l_short.n : f->l_msckf.n; }
  else if (n == "ivd0") { src = f->d_ivd0; cnt = 3 * (size_t)(f->last_which ? f->l_short.n : f->l_msckf.n); }
  else if (n == "gamma1") { src = f->d_gamma1; cnt = f->l_newms.n; }
  else if (n == "ivd1") { src = f->d_ivd1; cnt = 3 * (size_t)f->l_newms.n; }
  else if (n == "H1") { src = f->d_H1; cnt = (size_t)f->l_newms.n * 3 * W; }
  else if (n == "H2") { src = f->d_H2; cnt = 9 * (size_t)f->l_newms.n; }
  else if (n == "B0") { src = f->d_B0; cnt = (size_t)(f->last_which ? f->l_short.n : f->l_msckf.n) * 3 * W; }
  else if (n == "slam_gamma") { src = f->d_sgamma; cnt = f->l_slam.n; }
  else if (n == "range_gamma") { src = f->d_wgamma; cnt = 1; }
  else if (n == "wide_vals") { src = f->d_wvals; cnt = XB_WMAX * XB_WNZ; }
  else if (n == "wide_res") { src = f->d_wres; cnt = XB_WMAX; }
  else if (n == "slam_vals") { src = f->d_svals; cnt = 30 * (size_t)f->l_slam.n; }
  else if (n == "slam_res") { src = f->d_sres; cnt = 2 * (size_t)f->l_slam.n; }
  else if (n == "Tg") { src = f->d_Tg; cnt = (size_t)f->grows_pad * f->gcols_pad; }
  else if (n == "Rg") {
    if (gemm_uses_tensor_cores()) transpose(f->stream, f->d_Tg, f->d_Rg, f->gcols_pad, f->gcols_pad);  // not kept otherwise
    src = f->d_Rg; cnt = (size_t)f->gcols_pad * f->gcols_pad;
  }
  else if (n == "corr") { src = f->d_corr; cnt = f->N; }
  else if (n == "FQ") { src = f->d_FQ; cnt = 470; }
  else if (n == "delta") { src = f->d_delta; cnt = f->N; }
  else if (n == "om") { src = f->d_om; cnt = 21 * 21 + 32; }
  else if (n == "T") { src = f->d_T; cnt = f->T_doubles; }
  else if (n == "track_prof") {
    if (!f->d_track_prof) return fail(XB_E_INVALID, "set XB_TRACK_PROF=1 before xb_create");
    const int nt = f->last_which ? f->l_short.n : f->l_msckf.n;
    std::vector<long long> tr(12 * (size_t)nt);
    CK(cudaStreamSynchronize(f->stream));
    CK(cudaMemcpy(tr.data(), f->d_track_prof, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost));
    cnt = std::min((size_t)max_doubles, tr.size());
    for (size_t i = 0; i < cnt; ++i) out[i] = (double)(tr[i] - tr[i / 12 * 12]);
    return (int)cnt;
  }
  else if (n == "chol_trace") {
    if (!f->d_trace) return fail(XB_E_INVALID, "set XB_CHOL_TRACE=1 before xb_create");
    std::vector<long long> tr(10 * (size_t)f->trace_tiles);
    CK(cudaStreamSynchronize(f->stream));
  CK(cudaStreamSynchronize(f->side));
    CK(cudaMemcpy(tr.data(), f->d_trace, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost));
    cnt = std::min((size_t)max_doubles, tr.size());
    for (size_t i = 0; i < cnt; ++i) out[i] = (double)(tr[i] - ((i % 10) >= 2 ? tr[2] : 0));
    return (int)cnt;
  }
  else return fail(XB_E_INVALID, "unknown debug buffer " + n);
  if ((size_t)max_doubles < cnt) cnt = max_doubles;
  CK(cudaStreamSynchronize(f->stream));
  CK(cudaStreamSynchronize(f->side));
  CK(cudaMemcpy(out, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
  return (int)cnt;
}
extern "C" int xb_debug_read_int(xb_filter* f, const char* name, int* out, int max_ints) {
  const std::string n(name);
  const int* src = nullptr;
  size_t cnt = 0;
  if (n == "inlier0") { src = f->d_inl0; cnt = f->last_which ? f->l_short.n : f->l_msckf.n; }
  else if (n == "inlier1") { src = f->d_inl1; cnt = f->l_newms.n; }
  else if (n == "slam_inlier") { src = f->d_sinl; cnt = f->l_slam.n; }
  else if (n == "range_inlier") { src = f->d_winl; cnt = 1; }
  else if (n == "wide_cols") { src = f->d_wcols; cnt = XB_WMAX * XB_WNZ; }
  else if (n == "slam_cols") { src = f->d_scols; cnt = 15 * (size_t)f->l_slam.n; }
  else if (n == "slot_gen") {
    cnt = std::min((size_t)max_ints, (size_t)f->NS);
    for (size_t i = 0; i < cnt; ++i) out[i] = f->slot_gen[i];
    return (int)cnt;
  } else return fail(XB_E_INVALID, "unknown debug buffer " + n);
  if ((size_t)max_ints < cnt) cnt = max_ints;
  CK(cudaStreamSynchronize(f->stream));
  CK(cudaStreamSynchronize(f->side));
  CK(cudaMemcpy(out, src, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
  return (int)cnt;
}
