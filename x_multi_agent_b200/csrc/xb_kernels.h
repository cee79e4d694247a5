// Internal interface between the host orchestration (xb_api.cu) and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "xb_common.cuh"

namespace xb {

void count_launch();  // bumps the per-thread kernel-launch counter (xb_kernel_launches)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------------
// An update is ~45 small dependent launches; with the stream-serialisation attribute a kernel's CTAs are scheduled while its
// predecessor in the stream is still running and block in griddepcontrol.wait until that grid has completed and flushed,
// which removes the launch latency from every kernel-to-kernel edge.  Every kernel launched through XB_LAUNCH starts with
// XB_PDL_SHORT() (let the dependents in early, then wait) or XB_PDL_LONG() (wait only: the dependents of a long-running kernel
// must not sit resident next to the side-stream kernels for its whole duration).  No global memory is touched before the wait,
// so the stream order of reads and writes is unchanged.  XB_NO_PDL=1 launches without the attribute (A/B measurements).
#ifdef __CUDACC__
#define XB_PDL_SHORT() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#define XB_PDL_LONG() asm volatile("griddepcontrol.wait;" ::: "memory")
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void xb_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define XB_LAUNCH(kern, grid, block, smem, stream, ...) xb::xb_launch(kern, grid, block, smem, stream, ##__VA_ARGS__)
#endif

// ---- propagation --------------------------------------------------------------------------------
struct ImuSample { int valid; double t, seq, w[3], a[3]; };
struct PropParams { double g[3]; double n_w, n_bw, n_a, n_ba; };
#define XB_IMU_BATCH 32
struct ImuBatch { double v[XB_IMU_BATCH][8]; };   // per sample: w_m[3], a_m[3], time, seq (the xvec order 23..30)
void launch_imu_scatter(cudaStream_t s, double* xv, int LX, int NS, int start, int n, const ImuBatch& b);
void launch_propagate(cudaStream_t s, double* xv, int LX, double* strip, int N, int NS, int start, int n_steps,
                      const ImuSample& in, const PropParams& pp, double* FQ);
// one IMU step (means + row strips) in a single launch
void launch_prop_step(cudaStream_t s, double* xv, int LX, double* strip, int N, int NS, int start, const ImuSample& in,
                      const PropParams& pp, double* FQ);
// the two halves of launch_propagate: estimates + F_d/Q_d per step (one CTA), then the covariance strips
void launch_prop_means(cudaStream_t s, double* xv, int LX, int NS, int start, int n_steps, const ImuSample& in,
                       const PropParams& pp, double* FQ);
// second = 1: the column strips P_vi^T (same recurrence; their core block carries Q_d^T)
void launch_prop_strips(cudaStream_t s, double* strip, int N, int NS, int start, int n_steps, const double* FQ, int second = 0);

// ---- dense linear algebra -----------------------------------------------------------------------
// TMA-staged DMMA tiles for large shapes (k_gemm_tma.cu); return false when the shape / alignment does not qualify
bool gemm_nt_tma(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                 double* C, int ldc);
void gemm_nt_cpasync(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
                     double beta, double* C, int ldc);
bool downdate_sym_tma(cudaStream_t s, int n, int K, const double* W, int ldw, const double* Pin, double* Pout, int ldp,
                      const int* omega_inv = nullptr, const double* Zb = nullptr, const double* Yb = nullptr,
                      const double* Qb = nullptr);
void gemm_nt(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc);
void gemm_nn(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc);
void gemm_nt_splitk(cudaStream_t s, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                    int ldc, size_t strideC, int nz);
void tallchol(cudaStream_t s, double* T, int ld, int rows_pad, int cols_pad, int* flags, int* err, double piv_tol,
              const double* diag0 = nullptr, long long* trace = nullptr);
// Partial factorisation (k_linalg.cu, CholRange): phase 0 = plain; phase 1 = tile columns [0, jend) with the tile rows
// [jend, cols_pad) left out (their entries do not exist yet; the caller finishes with a Schur complement + plain call).
void tallchol_range(cudaStream_t s, double* T, int ld, int rows_pad, int cols_pad, int jstart_cols, int jend_cols, int phase,
                    int* flags, int* err, double piv_tol, const double* diag0, long long* trace, int share);
bool gemm_uses_tensor_cores();
void gemm_tn(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
             double* C, int ldc);
void gemm_tn_splitk(cudaStream_t s, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                    size_t strideC, int nz);
void downdate_f64(cudaStream_t s, double* P, int n, const double* T, int m_pad, int n_pad, const int* omega_inv,
                  const double* Zb, const double* Yb, const double* Qb);
// K-range form (small N): Pout = [sym](Pin) - W1[:, kbeg:kend] W1[:, kbeg:kend]^T [+ Woodbury / Omega tail]
void downdate_f64_range(cudaStream_t s, const double* Pin, double* Pout, int n, const double* T, int m_pad, int n_pad, int kbeg,
                        int kend, int do_sym, int do_tail, const int* omega_inv, const double* Zb, const double* Yb,
                        const double* Qb);
void symmetrise(cudaStream_t s, double* P, int n);
void gemv(cudaStream_t s, int rows, int cols, const double* A, int lda, const double* x, double* y);
void transpose(cudaStream_t s, const double* in, double* out, int rows, int cols);

// ---- per-track measurement construction ---------------------------------------------------------
struct TrackParams {
  const double* xv;   // work estimates (poses are read from p_array / q_array)
  int M, n_poses;
  const double* P;    // work covariance, N x N
  int ldp;
  const int* off;     // CSR offsets (device)
  const double* obs;  // 2 doubles per observation
  int n_tracks;
  int mode;           // 0 = MSCKF, 1 = MSCKF-SLAM
  int Lmax;           // longest track of the launch
  double var_img;
  const double* chi2_95;  // table: index = dof
  double gn_term;
  int gn_max_iter;
  // outputs
  double* ivd;     // [K][3]
  double* gamma;   // [K]
  int* inlier;     // [K]
  double* B;       // [K][3][6M+1]   U^T [J | r], zeroed for outliers
  double* Jout;    // [obs][14]      Jp(2x3) Ja(2x3) r(2)
  double* H1;      // [K][3][6M+1]   mode 1: unmasked U^T [h | r]  (H1 and r1 of Li 2012)
  double* H2;      // [K][9]         mode 1: U^T Hf
  double* D;       // [2*obs][6M+1]  mode 1: Pi [h | r]
  // MULTI_UAV (mode 0 only; all null otherwise): per-track group index (-1: not matched), the jointly
  // triangulated inverse-depth estimate per group, and the output A_up^T Hf per group
  const int* mm_grp;
  const double* mm_ivd;
  double* mm_F0;
  int asym_clones;   // the newest `asym_clones` clones carry unsymmetric covariance blocks (1 in the regular case)
  int oc;            // apply the reference's OC projection (xb_config.oc_projection)
  long long* prof;   // optional [K][12] clock64 stamps at the phase boundaries (XB_TRACK_PROF=1, tools/track_prof.py)
};
int launch_tracks(cudaStream_t s, const TrackParams& tp);

struct SlamParams {
  const double* xv;
  int M, N, n_poses;
  const double* P;
  const int* off;
  const double* obs;
  const int* anchor;
  const double* chi2;  // per track: quantile(0.9, 2*track_size)
  int n_tracks;
  double var_img;
  int* cols;      // [F][15]
  double* vals;   // [F][2][15]
  double* res;    // [F][2]
  double* gamma;  // [F]
  int* inlier;    // [F]
};
void launch_slam_rows(cudaStream_t s, const SlamParams& sp);

// Range (laser range finder) and sun-sensor rows of VioUpdater::constructUpdate (vio_updater.cpp:352-403;
// range_update.cpp:61-265, solar_update.cpp:39-94): up to XB_WMAX "wide" sparse rows of XB_WNZ entries each, stored after
// the SLAM row pairs of the compressed measurement.  Every row is scaled to the variance sigma_img^2 of the image rows:
// by sigma_img / sigma_row when the reference keeps the row's own variance (rows <= N + 1: no QR compression), by 1 when
// the reference's QR compression re-weights every row with sigma_img^2 (vio_updater.cpp:507-508).
#define XB_WMAX 3
#define XB_WNZ 36
struct SensorParams {
  const double* xv;
  int M, N, n_poses;
  const double* P;
  const int* anchor;
  int range_on, sun_on;
  double range, pt_x, pt_y;   // RangeMeasurement::range, img_pt_n
  int tri[3];                 // TrackManager::featureTriangleAtPoint: the facet's SLAM feature ids
  double var_range, chi2_1;   // sigma_range^2, quantile(0.9, 1)
  double w_range, w_sun;      // row scales (see above)
  double sun_x, sun_y;        // SunAngleMeasurement, degrees
  int* cols;      // [XB_WMAX][XB_WNZ]
  double* vals;   // [XB_WMAX][XB_WNZ]
  double* res;    // [XB_WMAX]
  double* gamma;  // [1] range gate
  int* inlier;    // [1]
};
void launch_sensor_rows(cudaStream_t s, const SensorParams& sp);

struct GramParams {
  int M, n_poses;
  const double* B; int rowsB; int nzB; double* partB;
  const double* D; int rowsD; int nzD; double* partD;
  const int* off; const int* inlier; int n_tracks_msckf; const double* Jout;
  double* blocks;  // [M][28]
  double* T; int ld, rows_pad, cols_pad;
  double* diag0;   // [cols_pad] original diagonal of G (relative pivot test of the semi-definite factorisation)
};
void launch_gram(cudaStream_t s, const GramParams& gp, cudaStream_t s_jtj = nullptr, cudaEvent_t ev_fork = nullptr,
                 cudaEvent_t ev_join = nullptr);

// ---- update assembly / state correction / state management (k_update.cu, k_manage.cu) -----------
struct UpdateDims {
  int M, F, N;
  int ms;       // slab rows = 6M
  int nslam;    // SLAM tracks (2 rows each)
  int nw;       // wide rows (range, sun sensor) after the SLAM row pairs: rows [2*nslam, 2*nslam + nw)
  int ns2;      // 2*nslam + nw: real rows of the sparse part
  int s_pad;    // sparse-part columns rounded up to 32 (0 without such rows)
  int ro;       // first slab column (= s_pad)
  int m;        // real rows: ms + ns2
  int m_pad;    // s_pad + ms rounded up to 32
  int n_pad;    // N rounded up to 32
  int ld;       // leading dimension of the tall buffer (= m_pad)
  const int* wcols;     // [nw][XB_WNZ]
  const double* wvals;  // [nw][XB_WNZ]
  const double* wres;   // [nw]
};
// Tall-buffer pieces on the SLAM columns (no Rg needed) and on the slab columns (k_update.cu)
void launch_build_slam_part(cudaStream_t s, const UpdateDims& d, const double* P, const int* scols, const double* svals,
                            const double* sres, const double* corr_total, double var, const int* omega, const int* omega_inv,
                            double* T);
void launch_slab_l21(cudaStream_t s, const UpdateDims& d, const double* Rg, const double* Lg, int ldr, double* T, const double* Bc);
void launch_slab_omega(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                       const int* omega, double* T, double* Gp);
void launch_slab_s22(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                     const double* zg, const double* corr_total, double var, double* T);
void launch_slab_schur(cudaStream_t s, const UpdateDims& d, double* T);
// Wsym = (W1s + W2s)/2 on the pose rows -> Bc (after the SLAM columns are factored)
void launch_wsym(cudaStream_t s, const UpdateDims& d, const int* omega_inv, const double* T, double* Bc);
void launch_build_slab_part(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                            const double* zg, const int* scols, const double* svals, const double* sres,
                            const double* corr_total, double var, const int* omega, double* T, const double* Bc, double* Gp);
// dense-H variant (Updater::applyUpdate with a caller-supplied H)
void launch_dense_prepare(cudaStream_t s, int m, int m_pad, int N, int n_pad, const double* P, const double* H,
                          const double* res, const double* rdiag, const double* corr_total, const int* omega, double* T);
// Woodbury factors, delta = K r - corr_total ; State::correct ; corr_total += delta
void launch_correct(cudaStream_t s, int M, int F, int N, double* T, int m_pad, int n_pad, const double* P,
                    const int* omega, const int* omega_inv, double* om, double* Zb, double* Yb, double* Qb, double* Cb,
                    double* xv, double* corr_total, double* delta_out, int* err = nullptr);
void launch_set_omega(cudaStream_t s, int* omega, int* omega_inv, int* tileflag, int slot, int M, int n_pad, int nflag);
void launch_apply_delta(cudaStream_t s, int M, int F, int N, const double* delta, double* xv, double* corr_total);
void launch_ci_cov(cudaStream_t s, double* P, int N, const double* K, const double* HP, int m);
// 3xTF32 tcgen05 tensor-core covariance downdate (k_downdate_tc.cu)
size_t downdate_tc_workspace_bytes(int n, int m_pad);
void downdate_tc(cudaStream_t s, double* P, int n, const double* T, int m_pad, int n_pad, const int* omega_inv,
                 const double* Zb, const double* Yb, const double* Qb, void* ws);

// StateManager::manage arithmetic: P' = A P A^T as one gather pass + thin products (k_manage.cu)
void launch_manage_dev(cudaStream_t s, int M, int F, int N, int n_poses, int n_features, int slide, int n_reanch,
                       const int* d_feat_src, const int* d_reanch, const int* d_rowmap, const int* d_ccols,
                       double* d_cvals, double* d_scratch, double* xv, const double* Pold, double* Pnew, double* Tm,
                       double* T2,
                       const double* strip = nullptr, const double* gen = nullptr, const double* strip2 = nullptr,
                       int general = 0);
// P(full) <- strip rows/cols + P_vv of a generation buffer (strip2: the column strip P_vi^T when it differs from strip)
void launch_assemble(cudaStream_t s, int N, const double* strip, const double* Pgen, double* Pwork,
                     const double* strip2 = nullptr);
void launch_extract_strip(cudaStream_t s, int N, const double* Pwork, double* strip);
void launch_extract_strip2(cudaStream_t s, int N, const double* Pwork, double* strip2);
// general (unsymmetric-prior) Updater::applyUpdate (k_general.cu)
void launch_gen_densify(cudaStream_t s, const UpdateDims& d, const int* scols, const double* svals, const double* sres,
                        const double* Lg, int ldr, const double* zg, double* Hd, double* res);
void general_update(cudaStream_t s, int N, int m, const double* Hd, const double* res, const double* rdiag, double var,
                    const double* corr_total, double* P, double* X, double* A, double* delta, int cov_update);
// MSCKF-SLAM / standard SLAM feature initialisation (state_manager.cpp:151-227)
struct FeatInitParams {
  int M, F, N, n_poses, n_features, n_new;
  const double* H1;   // [n_new][3][6M+1]
  const double* H2;   // [n_new][9]
  const double* ivd;  // [n_new][3]
  const double* corr; // [N]
  double var_img;
};
void launch_init_msckf_slam(cudaStream_t s, const FeatInitParams& fp, double* xv, double* P, double* scratch);
void launch_init_std_slam(cudaStream_t s, int M, int F, int N, int n_features, int n_new, const int* off,
                          const double* obs, double rho0, double var_img, double var_rho0, double* xv, double* P);
void launch_scale_blocks(cudaStream_t s, double* P, int N, const int* cols, int n_blocks, double w);
void launch_add_diag(cudaStream_t s, double* A, int ld, int n, double v);

// ---- covariance-intersection fusion (k_ci.cu) -----------------------------------------------------------
void launch_ci_pack(cudaStream_t s, const double* xv, const double* P, int N, int M, int F, int n_poses, int n_features,
                    const int* anchor, double* out);
void launch_ci_slam(cudaStream_t s, double* xv, double* P, int N, int M, int F, int n_poses, int n_features,
                    const int* anchor, const double* gathered, int payload_len, const int* matches, int n_matches,
                    double var_lm, double w_other, double chi2_90_3, double* rec, int* last_inlier, double* Kall,
                    double* delta, double* HP);


// ---- multi-agent MSCKF-MSCKF block (k_multi_msckf.cu; msckf_update.cpp:65-281, MULTI_UAV build) ----------
#define XB_MM_KMAX 7   // matched peers per own track
#define XB_MM_REC 32   // doubles per group record
struct MmParams {
  const double* xv; int M, n_poses, N;
  const double* P;            // work covariance (prior of the construct step)
  const int* off; const double* obs;   // own track list (CSR)
  const int* grp;             // [G][4]: own track, k peers, first entry, total joint observations
  const int* ent;             // [E][3]: peer slot in `gathered`, first observation in pobs, n_obs
  const double* pobs;         // peers' observations, 2 doubles each
  const double* chi2;         // [G] quantile(0.95, 2 n_tot - 3)
  int n_groups;
  const double* gathered; int pp_len;  // peers' pose payloads: [8 hdr | 3M camera positions | 4M quats | 6M x 6M cov]
  double var_img, w_other, gn_term; int gn_max_iter;
  int oc;                     // apply the reference's OC projection (xb_config.oc_projection)
  const double* B; const int* inlier;  // k_tracks outputs of the own tracks
  double* ivd;                // [G][3]
  double* F0;                 // [G][9]
  double* rec;                // [G][XB_MM_REC]: inlier, gamma, chi2, w_result, a(3), C3(9), trk, i1, L, k
  int* last;                  // last inlier group (-1: none)
};
void launch_mm_triangulate(cudaStream_t s, const MmParams& mp);
int launch_mm_construct(cudaStream_t s, const MmParams& mp);
// applyCI over the group list: sequential state corrections, covariance of the last inlier entry (updater.cpp:144-161)
void launch_mm_apply(cudaStream_t s, const MmParams& mp, double* P, double* xv, int F, double* V, int ldv, double* D,
                     double* K3, double* HP3);
void launch_pack_poses(cudaStream_t s, const double* xv, const double* P, int N, int M, double* out);

}  // namespace xb
