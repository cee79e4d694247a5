// General (unsymmetric-prior) form of Updater::applyUpdate.  reference: src/x/ekf/updater.cpp:117-141.
//
// The reference never symmetrises its covariance between updates: an update WITHOUT measurement rows runs
// StateManager::manage only (updater.cpp:106), so after k such updates in a row the antisymmetric part of P has spread
// from the core block into the k newest clones (state_manager.cpp:273-349 copies the unsymmetric core block into every
// new clone; propagator.cpp:197-203 propagates P_iv and P_vi separately).  The structured path of k_update.cu
// represents that part on Omega = core + newest clone (rank <= 21 Woodbury); this file is the exact path for the rare
// wider case (start-up frames / featureless frames): the reference's own formulas on the dense compressed measurement,
//   X = P Hc^T,  Y = Hc P,  S = Hc X + R,  Z = S^-1 [Y | r_eff],  delta = X z_r - corr_total,  P <- sym(P - X Z_Y),
// with S^-1 applied by Gauss-Jordan elimination of the augmented matrix (S has a positive definite symmetric part, so
// no pivoting is needed).  Throughput is irrelevant here; every step is a plain kernel or one of the DMMA GEMMs.
#include "xb_kernels.h"

namespace xb {

// Hd (m x N, row-major): rows [0, ns2) the sparse SLAM rows, rows [ns2, ns2 + ms) Rg on the pose columns (Rg = Lg^T);
// res likewise (SLAM residuals, then z of the Gram factorisation)
__global__ void k_gen_densify(UpdateDims d, const int* __restrict__ scols, const double* __restrict__ svals,
                              const double* __restrict__ sres, const double* __restrict__ Lg, int ldr,
                              const double* __restrict__ zg, double* __restrict__ Hd, double* __restrict__ res) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (r >= d.m || c >= d.N) return;
  double v = 0.0;
  if (r >= 2 * d.nslam && r < d.ns2) {  // wide row (range / sun sensor)
    const int w = r - 2 * d.nslam;
    for (int e = 0; e < XB_WNZ; ++e)
      if (d.wcols[XB_WNZ * w + e] == c) v += d.wvals[XB_WNZ * w + e];
    if (c == 0) res[r] = d.wres[w];
  } else if (r < d.ns2) {
    const int j = r >> 1, h = r & 1;
    for (int e = 0; e < 15; ++e)
      if (scols[15 * j + e] == c) v += svals[30 * j + 15 * h + e];
    if (c == 0) res[r] = sres[r];
  } else {
    const int a = r - d.ns2;
    if (c >= XB_CORE && c < XB_CORE + d.ms) v = Lg[(size_t)(c - XB_CORE) * ldr + a];  // Rg[a][b] = Lg[b][a]
    if (c == 0) res[r] = zg[a];
  }
  Hd[(size_t)r * d.N + c] = v;
}
void launch_gen_densify(cudaStream_t s, const UpdateDims& d, const int* scols, const double* svals, const double* sres,
                        const double* Lg, int ldr, const double* zg, double* Hd, double* res) {
  dim3 g((d.N + 127) / 128, d.m);
  k_gen_densify<<<g, 128, 0, s>>>(d, scols, svals, sres, Lg, ldr, zg, Hd, res);
  count_launch();
}

// A[:, m + N] = r_eff = res + Hd corr_total ; diagonal of the S block += rdiag (or var)
__global__ void k_gen_finish(int m, int N, int lda, const double* __restrict__ Hd, const double* __restrict__ res,
                             const double* __restrict__ rdiag, double var, const double* __restrict__ corr,
                             double* __restrict__ A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double v = res[r];
  if (corr)
    for (int b = 0; b < N; ++b) v = fma(Hd[(size_t)r * N + b], corr[b], v);
  A[(size_t)r * lda + m + N] = v;
  A[(size_t)r * lda + r] += rdiag ? rdiag[r] : var;
}

// One Gauss-Jordan step without pivoting on the augmented matrix A (m x lda): rows i != k lose their entry in column k.
// Column k itself and the columns before it are never read again, so only columns > k are touched; the pivot row is
// left unscaled (k_gen_scale divides at the end), which keeps every launch free of read/write races.
__global__ void __launch_bounds__(256) k_gen_gj_step(int m, int cols, int lda, int k, double* __restrict__ A) {
  const int j = k + 1 + blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i == k || j >= cols) return;
  const double f = A[(size_t)i * lda + k] / A[(size_t)k * lda + k];
  if (f != 0.0) A[(size_t)i * lda + j] = fma(-f, A[(size_t)k * lda + j], A[(size_t)i * lda + j]);
}
__global__ void k_gen_scale(int m, int cols, int lda, double* __restrict__ A) {
  const int j = m + blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= cols) return;
  A[(size_t)i * lda + j] /= A[(size_t)i * lda + i];
}
// delta_i = X_i . z_r - corr_i    (one warp per state)
__global__ void __launch_bounds__(128) k_gen_delta(int N, int m, int lda, const double* __restrict__ X,
                                                   const double* __restrict__ A, const double* __restrict__ corr,
                                                   double* __restrict__ delta) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  double s = 0.0;
  for (int c = lane; c < m; c += 32) s = fma(X[(size_t)i * m + c], A[(size_t)c * lda + lda - 1], s);
  s = xb_warp_sum(s);
  if (lane == 0) delta[i] = s - (corr ? corr[i] : 0.0);
}

// Hd: m x N dense measurement Jacobian, res: m, rdiag: m or nullptr (then var on the diagonal).
// X: scratch N x m; A: scratch m x (m + N + 1).  P is updated in place when cov_update != 0; delta (N) is written.
void general_update(cudaStream_t s, int N, int m, const double* Hd, const double* res, const double* rdiag, double var,
                    const double* corr_total, double* P, double* X, double* A, double* delta, int cov_update) {
  const int lda = m + N + 1;
  gemm_nt(s, N, m, N, 1.0, P, N, Hd, N, 0.0, X, m);          // X = P Hc^T
  gemm_nn(s, m, m, N, 1.0, Hd, N, X, m, 0.0, A, lda);         // S = Hc X
  gemm_nn(s, m, N, N, 1.0, Hd, N, P, N, 0.0, A + m, lda);     // Y = Hc P
  k_gen_finish<<<(m + 127) / 128, 128, 0, s>>>(m, N, lda, Hd, res, rdiag, var, corr_total, A);
  count_launch();
  for (int k = 0; k < m; ++k) {
    dim3 g((lda - k - 1 + 255) / 256, m);
    k_gen_gj_step<<<g, 256, 0, s>>>(m, lda, lda, k, A);
    count_launch();
  }
  {
    dim3 g((N + 1 + 255) / 256, m);
    k_gen_scale<<<g, 256, 0, s>>>(m, lda, lda, A);
    count_launch();
  }
  k_gen_delta<<<(N * 32 + 127) / 128, 128, 0, s>>>(N, m, lda, X, A, corr_total, delta);
  count_launch();
  if (cov_update) {
    gemm_nn(s, N, N, m, -1.0, X, m, A + m, lda, 1.0, P, N);   // P - K (Hc P)
    symmetrise(s, P, N);                                       // updater.cpp:133
  }
}

}  // namespace xb
