// Assembly of the Kalman-update operands from the device-resident compressed measurement
//   Hc = [ Rg (6M x 6M, on the pose columns) ; SLAM rows (2 per feature, <=15 columns each) ]
// and the state correction.  reference: src/x/vio/vio_updater.cpp:405-423,487-512 (stack + QR compress),
// src/x/ekf/updater.cpp:117-141 (applyUpdate), src/x/ekf/state.cpp:197-249 (State::correct).
//
// Tall buffer T (leading dimension ld = m_pad):
//   rows [0, m_pad)                S = Hc P Hc^T + var I   (identity on the padding)
//   rows [m_pad, m_pad + n_pad)    A1 = P Hc^T
//   row  m_pad + n_pad             r_eff^T = (res + Hc * correction_total)^T          (aux tile, 32 rows)
//   rows m_pad + n_pad + 32 + k    A2[Omega_k, :] = (P^T Hc^T)[Omega_k, :], k < 21       (Omega tile, 32 rows)
//   rows m_pad + n_pad + 64 + k    V^T[k] = Hc[:, Omega_k]^T, k < 21                     (V tile, 32 rows)
//
// Omega = the 15 core states + the 6 states of the newest clone: between two updates the reference's P is
// NOT symmetric there (its Q_d is not symmetric, propagator.cpp:207-840, and augmentCovariance copies the
// core block into the clone, state_manager.cpp:273-349).  With E = (P - P^T)/2 (support Omega x Omega) the
// reference's  K = P H^T (H P H^T + R)^-1,  P <- sym((I - K H) P)  is evaluated exactly as
//   S_s = sym(H P H^T) + R = L L^T,  W1 = A1 L^-T,  W2 = A2 L^-T (differs from W1 on Omega rows only),
//   Vt = L^-1 V,  G = Vt^T Vt,  C = E (I + G E)^-1                                   (Woodbury, rank <= 21)
//   K A2^T = W1 W2^T - (W1 Vt) C (W2 Vt)^T ,   K r = W1 z - (W1 Vt) C (Vt^T z).
#include "xb_kernels.h"

namespace xb {

#define NOM 21  // |Omega| = 15 core + 6 clone states
#define CBZ 4   // split-K partials of the thin Woodbury GEMM

// Column order of the compressed measurement inside the tall buffer (UpdateDims): the SLAM rows come first,
//   columns [0, 2 nslam) SLAM rows | padding up to s_pad | columns [ro, ro + 6M) slab rows Rg | padding up to m_pad
// so that everything belonging to the SLAM columns -- including the first s_pad/32 tile columns of the Cholesky
// factorisation -- can be computed before Rg exists (xb_api.cu runs it next to the MSCKF track pipeline).
//
// PHt[i, 2j + r] = sum_e P[i, col_e] * val[j][r][e]
__global__ void k_pht_slam(UpdateDims d, const double* __restrict__ P, const int* __restrict__ scols,
                           const double* __restrict__ svals, const int* __restrict__ omega_inv, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= d.N || j >= d.nslam) return;
  // P[i][col] is read as P[col][i] (consecutive threads -> consecutive addresses; the row-wise form is a stride-N gather
  // that costs 4x the sectors, 1.6 ms at cfg-5) wherever P is symmetric, i.e. everywhere except Omega x Omega
  const bool i_om = omega_inv[i] >= 0;
  int col[15];
  double pv[15];
#pragma unroll
  for (int e = 0; e < 15; ++e) col[e] = scols[15 * j + e];
#pragma unroll
  for (int e = 0; e < 15; ++e)
    pv[e] = (i_om && omega_inv[col[e]] >= 0) ? P[(size_t)i * d.N + col[e]] : P[(size_t)col[e] * d.N + i];
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int e = 0; e < 15; ++e) {
    a0 = fma(pv[e], svals[30 * j + e], a0);
    a1 = fma(pv[e], svals[30 * j + 15 + e], a1);
  }
  double* row = T + (size_t)(d.m_pad + i) * d.ld + 2 * j;
  row[0] = a0;
  row[1] = a1;
}

// Wide rows (range / sun sensor, XB_WNZ entries each): PHt[i, 2 nslam + w] = sum_e P[i, col_e] * val[w][e]
__global__ void k_pht_wide(UpdateDims d, const double* __restrict__ P, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
  if (i >= d.N) return;
  double a = 0.0;
  for (int e = 0; e < XB_WNZ; ++e) {
    const double v = d.wvals[XB_WNZ * w + e];
    if (v != 0.0) a = fma(P[(size_t)i * d.N + d.wcols[XB_WNZ * w + e]], v, a);
  }
  T[(size_t)(d.m_pad + i) * d.ld + 2 * d.nslam + w] = a;
}
// S[2 nslam + w, c] = sum_e val[w][e] * PHt[col_e, c] for every sparse-part column c
__global__ void k_s_wide(UpdateDims d, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
  if (c >= d.ns2) return;
  double a = 0.0;
  for (int e = 0; e < XB_WNZ; ++e) {
    const double v = d.wvals[XB_WNZ * w + e];
    if (v != 0.0) a = fma(v, T[(size_t)(d.m_pad + d.wcols[XB_WNZ * w + e]) * d.ld + c], a);
  }
  T[(size_t)(2 * d.nslam + w) * d.ld + c] = a;
}

// After the SLAM tile columns are factored (S11 = L11 L11^T, W1s = (P Hs^T) L11^-T, W2s on the Omega rows), the
// sub-diagonal block of the factor needs no triangular solve:
//   L21 = S21 L11^-T = Rg (sym(P) Hs^T)[pose rows] L11^-T = Rg * Wsym,   Wsym = (W1s + W2s)/2 on the pose rows
// (W2s differs from W1s on the newest clone's 6 rows only).  This kernel gathers Wsym (6M x s_pad) into Bc, and
// is multiplied by launch_build_slab_part.
__global__ void k_wsym(UpdateDims d, const int* __restrict__ omega_inv, const double* __restrict__ T, double* __restrict__ Bc) {
  XB_PDL_SHORT();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= d.s_pad || b >= d.ms) return;
  double v = T[(size_t)(d.m_pad + XB_CORE + b) * d.ld + c];
  const int k = omega_inv[XB_CORE + b];
  if (k >= 0) v = 0.5 * (v + T[(size_t)(d.m_pad + d.n_pad + 32 + k) * d.ld + c]);
  Bc[(size_t)b * d.s_pad + c] = v;
}
void launch_wsym(cudaStream_t s, const UpdateDims& d, const int* omega_inv, const double* T, double* Bc) {
  if (d.ns2 <= 0) return;
  dim3 g((d.s_pad + 127) / 128, d.ms);
  XB_LAUNCH(k_wsym, g, 128, 0, s, d, omega_inv, T, Bc);
  count_launch();
}

// S[2j + r, c] = sum_e val[j][r][e] * PHt[col_e, c]   for the columns c0 + [0, nc)
__global__ void k_s_slam(UpdateDims d, int c0, int nc, const int* __restrict__ scols, const double* __restrict__ svals,
                         double* __restrict__ T) {
  XB_PDL_SHORT();
  const int cl = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (cl >= nc || j >= d.nslam) return;
  const int c = c0 + cl;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int e = 0; e < 15; ++e) {
    const double p = T[(size_t)(d.m_pad + scols[15 * j + e]) * d.ld + c];
    a0 = fma(svals[30 * j + e], p, a0);
    a1 = fma(svals[30 * j + 15 + e], p, a1);
  }
  T[(size_t)(2 * j) * d.ld + c] = a0;
  T[(size_t)(2 * j + 1) * d.ld + c] = a1;
}

// diagonal (+var on real rows, 1 on padding) and the r_eff row, for the rows/columns a in c0 + [0, nc)
__global__ void k_s_finish(UpdateDims d, int c0, int nc, const double* __restrict__ Lg, int ldr, const double* __restrict__ zg,
                           const int* __restrict__ scols, const double* __restrict__ svals, const double* __restrict__ sres,
                           const double* __restrict__ corr, double var, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int al = blockIdx.x * blockDim.x + threadIdx.x;
  if (al >= nc) return;
  const int a = c0 + al;
  double* reff = T + (size_t)(d.m_pad + d.n_pad) * d.ld;
  const bool slam = a < d.ns2, slab = a >= d.ro && a < d.ro + d.ms;
  if (!slam && !slab) {
    T[(size_t)a * d.ld + a] = 1.0;
    reff[a] = 0.0;
    return;
  }
  T[(size_t)a * d.ld + a] += var;
  double r;
  if (slab) {
    const int ar = a - d.ro;
    r = zg[ar];
    if (corr)
      for (int b = 0; b < d.ms; ++b) r = fma(Lg[(size_t)b * ldr + ar], corr[XB_CORE + b], r);  // Rg[a][b] = Lg[b][a]
  } else if (a >= 2 * d.nslam) {  // wide row (range / sun sensor)
    const int w = a - 2 * d.nslam;
    r = d.wres[w];
    if (corr)
      for (int e = 0; e < XB_WNZ; ++e) r = fma(d.wvals[XB_WNZ * w + e], corr[d.wcols[XB_WNZ * w + e]], r);
  } else {
    const int j = a >> 1, h = a & 1;
    r = sres[2 * j + h];
    if (corr)
      for (int e = 0; e < 15; ++e) r = fma(svals[30 * j + 15 * h + e], corr[scols[15 * j + e]], r);
  }
  reff[a] = r;
}

void launch_sym_lower(cudaStream_t s, double* T, int ld, int r0, int r1, int c0);
void launch_omega_rows(cudaStream_t s, const UpdateDims& d, int c0, int nc, const double* P, const double* Lg, int ldr,
                       const int* scols, const double* svals, const int* omega, double* T);

// Everything of the tall buffer that lives on the SLAM columns [0, s_pad): PHt, S[slam, slam], diagonal, r_eff, Omega/V rows.
void launch_build_slam_part(cudaStream_t s, const UpdateDims& d, const double* P, const int* scols, const double* svals,
                            const double* sres, const double* corr_total, double var, const int* omega, const int* omega_inv,
                            double* T) {
  if (d.ns2 <= 0) return;
  if (d.nslam > 0) {
    dim3 g((d.N + 127) / 128, d.nslam);
    XB_LAUNCH(k_pht_slam, g, 128, 0, s, d, P, scols, svals, omega_inv, T);
    count_launch();
  }
  if (d.nw > 0) {
    dim3 g((d.N + 127) / 128, d.nw);
    XB_LAUNCH(k_pht_wide, g, 128, 0, s, d, P, T);
    count_launch();
  }
  if (d.nslam > 0) {
    dim3 g((d.ns2 + 127) / 128, d.nslam);
    XB_LAUNCH(k_s_slam, g, 128, 0, s, d, 0, d.ns2, scols, svals, T);
    count_launch();
  }
  if (d.nw > 0) {
    dim3 g((d.ns2 + 127) / 128, d.nw);
    XB_LAUNCH(k_s_wide, g, 128, 0, s, d, T);
    count_launch();
  }
  launch_sym_lower(s, T, d.ld, 0, d.ns2, 0);
  XB_LAUNCH(k_s_finish, (d.s_pad + 127) / 128, 128, 0, s, d, 0, d.s_pad, nullptr, 0, nullptr, scols, svals, sres, corr_total, var, T);
  count_launch();
  launch_omega_rows(s, d, 0, d.ns2, P, nullptr, 0, scols, svals, omega, T);
}

// lower(S22) <- sym, diagonal (+var on real rows, 1 on padding) and r_eff of the slab block (k_sym_lower + k_s_finish fused)
__global__ void k_slab_sym_finish(UpdateDims d, const double* __restrict__ Lg, int ldr, const double* __restrict__ zg,
                                  const double* __restrict__ corr, double var, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int c = d.ro + blockIdx.x * blockDim.x + threadIdx.x, r = d.ro + blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= d.m_pad || c > r) return;
  const bool real = r < d.ro + d.ms;
  if (c < r) {
    if (real) T[(size_t)r * d.ld + c] = 0.5 * (T[(size_t)r * d.ld + c] + T[(size_t)c * d.ld + r]);
    return;
  }
  double* reff = T + (size_t)(d.m_pad + d.n_pad) * d.ld;
  if (!real) {
    T[(size_t)r * d.ld + r] = 1.0;
    reff[r] = 0.0;
    return;
  }
  T[(size_t)r * d.ld + r] += var;
  const int ar = r - d.ro;
  double rr = zg[ar];
  if (corr)
    for (int b = 0; b < d.ms; ++b) rr = fma(Lg[(size_t)b * ldr + ar], corr[XB_CORE + b], rr);  // Rg[a][b] = Lg[b][a]
  reff[r] = rr;
}

// Gp[k][b] = P[15 + b, Omega_k] (32 x 6M, row-major) and the V tile on the slab columns: V^T[k][ro + a] = Rg[a][Omega_k - 15]
__global__ void k_omega_gather(UpdateDims d, const double* __restrict__ P, const double* __restrict__ Lg, int ldr,
                               const int* __restrict__ omega, double* __restrict__ Gp, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int b = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (b >= d.ms) return;
  const int ok = omega[k];
  Gp[(size_t)k * d.ms + b] = P[(size_t)(XB_CORE + b) * d.N + ok];
  double v = 0.0;
  if (ok >= XB_CORE && ok < XB_CORE + d.ms) v = Lg[(size_t)(ok - XB_CORE) * ldr + b];  // Rg[a][b'] = Lg[b'][a]
  T[(size_t)(d.m_pad + d.n_pad + 64 + k) * d.ld + d.ro + b] = v;
}

// The slab columns [ro, m_pad) once Rg exists: PHt[:, slab] = P[:, pose] Rg^T, S22 = Rg * PHt[pose rows, slab], and the
// finished factor block L21 = Rg * Wsym on the SLAM columns (k_wsym).
// The four pieces are independent up to the Schur complement, which needs all of them: xb_api.cu runs L21 and the Omega
// tile on side streams next to the P H_R^T -> S22 chain.
// Rg (the transposed Gram factor) is only materialised for the CUDA-core GEMM fallback; the tensor-core kernel takes the
// factor Lg = Rg^T as a k-major / [K x N] operand directly (Rg == nullptr).
void launch_slab_l21(cudaStream_t s, const UpdateDims& d, const double* Rg, const double* Lg, int ldr, double* T, const double* Bc) {
  if (d.ns2 <= 0) return;
  if (Rg) gemm_nn(s, d.ms, d.ns2, d.ms, 1.0, Rg, ldr, Bc, d.s_pad, 0.0, T + (size_t)d.ro * d.ld, d.ld);
  else gemm_tn(s, d.ms, d.ns2, d.ms, 1.0, Lg, ldr, Bc, d.s_pad, 0.0, T + (size_t)d.ro * d.ld, d.ld);
}
void launch_slab_omega(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                       const int* omega, double* T, double* Gp) {
  // Omega tile on the slab columns: A2[Omega_k, a] = sum_b P[15+b, Omega_k] Rg[a][b] = (Gp Rg^T)[k][a] with the gathered
  // Gp[k][b] = P[15+b, Omega_k] (k_omega_gather, which also writes the V rows Rg[:, Omega_k]^T); a 21 x 6M x 6M tensor-core GEMM
  dim3 g((d.ms + 127) / 128, NOM);
  XB_LAUNCH(k_omega_gather, g, 128, 0, s, d, P, Lg, ldr, omega, Gp, T);
  count_launch();
  double* dst = T + (size_t)(d.m_pad + d.n_pad + 32) * d.ld + d.ro;
  if (Rg) gemm_nt(s, NOM, d.ms, d.ms, 1.0, Gp, d.ms, Rg, ldr, 0.0, dst, d.ld);
  else gemm_nn(s, NOM, d.ms, d.ms, 1.0, Gp, d.ms, Lg, ldr, 0.0, dst, d.ld);
}
void launch_slab_s22(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                     const double* zg, const double* corr_total, double var, double* T) {
  double* PHt = T + (size_t)d.m_pad * d.ld;
  double* S22 = T + (size_t)d.ro * d.ld + d.ro;
  if (Rg) {
    gemm_nt(s, d.N, d.ms, d.ms, 1.0, P + XB_CORE, d.N, Rg, ldr, 0.0, PHt + d.ro, d.ld);
    gemm_nn(s, d.ms, d.ms, d.ms, 1.0, Rg, ldr, PHt + (size_t)XB_CORE * d.ld + d.ro, d.ld, 0.0, S22, d.ld);
  } else {
    gemm_nn(s, d.N, d.ms, d.ms, 1.0, P + XB_CORE, d.N, Lg, ldr, 0.0, PHt + d.ro, d.ld);
    gemm_tn(s, d.ms, d.ms, d.ms, 1.0, Lg, ldr, PHt + (size_t)XB_CORE * d.ld + d.ro, d.ld, 0.0, S22, d.ld);
  }
  {  // symmetrise the slab block, add the measurement variance, identity on the padding, r_eff: one launch
    dim3 b(32, 8), g((d.m_pad - d.ro + 31) / 32, (d.m_pad - d.ro + 7) / 8);
    XB_LAUNCH(k_slab_sym_finish, g, b, 0, s, d, Lg, ldr, zg, corr_total, var, T);
    count_launch();
  }
}
void launch_slab_schur(cudaStream_t s, const UpdateDims& d, double* T) {
  if (d.ns2 <= 0) return;
  // Schur complement of the factored SLAM columns on every row from the slab rows down (S22, P H^T, r_eff, Omega, V):
  //   T[ro:, ro:] -= T[ro:, 0:s_pad] * L21^T     -- after it the slab columns are a plain tall factorisation of their own
  const int rows = d.m_pad - d.ro + d.n_pad + 96;
  gemm_nt(s, rows, d.m_pad - d.ro, d.s_pad, -1.0, T + (size_t)d.ro * d.ld, d.ld, T + (size_t)d.ro * d.ld, d.ld, 1.0,
          T + (size_t)d.ro * d.ld + d.ro, d.ld);
}
void launch_build_slab_part(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, const double* Lg, int ldr,
                            const double* zg, const int* scols, const double* svals, const double* sres,
                            const double* corr_total, double var, const int* omega, double* T, const double* Bc, double* Gp) {
  launch_slab_l21(s, d, Rg, Lg, ldr, T, Bc);
  launch_slab_s22(s, d, P, Rg, Lg, ldr, zg, corr_total, var, T);
  launch_slab_omega(s, d, P, Rg, Lg, ldr, omega, T, Gp);
  launch_slab_schur(s, d, T);
}

// dense-H path: S += diag(rdiag) (+ identity padding), r_eff = res + H corr
__global__ void k_dense_finish(int m, int m_pad, int n_pad, int N, const double* __restrict__ H, const double* __restrict__ res,
                               const double* __restrict__ rdiag, const double* __restrict__ corr, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m_pad) return;
  double* reff = T + (size_t)(m_pad + n_pad) * m_pad;
  if (a >= m) {
    T[(size_t)a * m_pad + a] = 1.0;
    reff[a] = 0.0;
    return;
  }
  T[(size_t)a * m_pad + a] += rdiag[a];
  double r = res[a];
  if (corr)
    for (int b = 0; b < N; ++b) r = fma(H[(size_t)a * N + b], corr[b], r);
  reff[a] = r;
}
// ---- Omega rows for the VIO (structured Hc) path ------------------------------------------------------
__global__ void k_omega_rows(UpdateDims d, int c0, int nc, const double* __restrict__ P, const double* __restrict__ Lg, int ldr,
                             const int* __restrict__ scols, const double* __restrict__ svals, const int* __restrict__ omega,
                             double* __restrict__ T) {
  XB_PDL_SHORT();
  const int al = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (al >= nc) return;
  const int a = c0 + al;
  const int ok = omega[k];
  double v = 0.0, a2 = 0.0;
  if (a >= d.ro) {
    const int ar = a - d.ro;
    if (ok >= XB_CORE && ok < XB_CORE + d.ms) v = Lg[(size_t)(ok - XB_CORE) * ldr + ar];  // Rg[a][b] = Lg[b][a]
    for (int b = 0; b < d.ms; ++b) a2 = fma(P[(size_t)(XB_CORE + b) * d.N + ok], Lg[(size_t)b * ldr + ar], a2);
  } else if (a >= 2 * d.nslam) {  // wide row (range / sun sensor)
    const int w = a - 2 * d.nslam;
    if (w < d.nw)
      for (int e = 0; e < XB_WNZ; ++e) {
        const double hv = d.wvals[XB_WNZ * w + e];
        const int col = d.wcols[XB_WNZ * w + e];
        if (hv != 0.0) {
          if (col == ok) v += hv;
          a2 = fma(P[(size_t)col * d.N + ok], hv, a2);
        }
      }
  } else {
    const int j = a >> 1, h = a & 1;
    int col[15];
    double hv[15], pv[15];
#pragma unroll
    for (int e = 0; e < 15; ++e) { col[e] = scols[15 * j + e]; hv[e] = svals[30 * j + 15 * h + e]; }
#pragma unroll
    for (int e = 0; e < 15; ++e) pv[e] = P[(size_t)col[e] * d.N + ok];  // 15 independent loads in flight
#pragma unroll
    for (int e = 0; e < 15; ++e) {
      if (col[e] == ok) v += hv[e];
      a2 = fma(pv[e], hv[e], a2);
    }
  }
  T[(size_t)(d.m_pad + d.n_pad + 32 + k) * d.ld + a] = a2;
  T[(size_t)(d.m_pad + d.n_pad + 64 + k) * d.ld + a] = v;
}
// dense-H variant
__global__ void k_omega_rows_dense(int m, int m_pad, int n_pad, int N, const double* __restrict__ P,
                                   const double* __restrict__ H, const int* __restrict__ omega, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int a = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (a >= m) return;
  const int ok = omega[k];
  double a2 = 0.0;
  for (int j = 0; j < N; ++j) a2 = fma(P[(size_t)j * N + ok], H[(size_t)a * N + j], a2);
  T[(size_t)(m_pad + n_pad + 32 + k) * m_pad + a] = a2;
  T[(size_t)(m_pad + n_pad + 64 + k) * m_pad + a] = H[(size_t)a * N + ok];
}
// lower(S) <- lower((S + S^T)/2) for the rows [r0, r1), columns >= c0
__global__ void k_sym_lower(double* __restrict__ T, int ld, int r0, int r1, int c0) {
  XB_PDL_SHORT();
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x, r = r0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (r < r1 && c < r) T[(size_t)r * ld + c] = 0.5 * (T[(size_t)r * ld + c] + T[(size_t)c * ld + r]);
}

// After the tile Cholesky and the thin GEMM  Cb = [W1 ; aux ; dW ; Vt] [z ; dW ; Vt]^T  (rows n_pad+64+k of Cb hold
// Vt_k . {z, dW, Vt}):  G = Vt Vt^T, q = Vt z, E from P, C = E (I + G E)^-1.   om = [C 21x21 | q 21]
__global__ void __launch_bounds__(256) k_omega_small(int N, int n_pad, const double* __restrict__ Cb,
                                                     const double* __restrict__ P, const int* __restrict__ omega,
                                                     double* __restrict__ om, int* __restrict__ err) {
  XB_PDL_SHORT();
  __shared__ double G[NOM][NOM], E[NOM][NOM], A[2][NOM][2 * NOM + 1], q[NOM];
  __shared__ int perm[NOM];
  const int t = threadIdx.x, lane = t & 31;
  for (int e = t; e < NOM * NOM; e += blockDim.x) {
    const int k = e / NOM, l = e % NOM;
    double g = 0.0;
    for (int z = 0; z < CBZ; ++z) g += Cb[(size_t)z * (n_pad + 96) * 96 + (size_t)(n_pad + 64 + k) * 96 + 64 + l];
    G[k][l] = g;
    E[k][l] = 0.5 * (P[(size_t)omega[k] * N + omega[l]] - P[(size_t)omega[l] * N + omega[k]]);
  }
  if (t < NOM) {
    double g = 0.0;
    for (int z = 0; z < CBZ; ++z) g += Cb[(size_t)z * (n_pad + 96) * 96 + (size_t)(n_pad + 64 + t) * 96];
    q[t] = g;
  }
  __syncthreads();
  // A = [I + G E | I]   (G E is of order one: the antisymmetric part of the covariance is as large as the covariance itself
  // in the velocity/attitude cross blocks, so there is no series shortcut -- measured max |G E| = 0.6 at cfg-2)
  for (int e = t; e < NOM * 2 * NOM; e += blockDim.x) {
    const int r = e / (2 * NOM), c = e % (2 * NOM);
    double v;
    if (c < NOM) {
      v = (r == c) ? 1.0 : 0.0;
      for (int x = 0; x < NOM; ++x) v = fma(G[r][x], E[x][c], v);
    } else {
      v = (c - NOM == r) ? 1.0 : 0.0;
    }
    A[0][r][c] = v;
  }
  // Gauss-Jordan with (implicit) partial pivoting, all 256 threads, ONE barrier per pivot: the pivot row stays where it is
  // (perm[c] remembers it, rows already used are masked out of the search), every warp finds the pivot redundantly from
  // shared memory, and the elimination writes into the other copy of the matrix, so no element is read after it has been
  // overwritten.  The pivot row is normalised at the end.  Same pivots as the explicit row-swapping form (first maximum of
  // |A[r][c]| over the unused rows); that form needed five barriers per pivot (23 us for the kernel).
  int er[4], ex[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = t + u * 256;
    er[u] = e < NOM * 2 * NOM ? e / (2 * NOM) : -1;
    ex[u] = e % (2 * NOM);
  }
  unsigned used = 0;  // same value in every thread
  int cur = 0;
  __syncthreads();
  for (int c = 0; c < NOM; ++c) {
    // first maximum of |A[r][c]| over the unused rows: non-negative doubles order like their bit patterns, so two
    // warp-wide integer max reductions (high word, then low word among the ties) + a ballot replace the 15 dependent
    // shuffles of a (value, index) butterfly
    const bool cand = lane < NOM && !((used >> lane) & 1u);
    const double av = cand ? fabs(A[cur][lane][c]) : 0.0;
    const unsigned long long key = cand ? (unsigned long long)__double_as_longlong(av) + 1ull : 0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned win = __ballot_sync(0xffffffffu, cand && hi == mh && lo == ml);
    const int bi = win ? __ffs(win) - 1 : -1;
    const double bv = __shfl_sync(0xffffffffu, av, bi < 0 ? 0 : bi);
    // a non-finite column (the factorisation met a non-positive pivot: S was not positive definite) must not select a
    // row outside the matrix; the error word makes xb_synchronize report it
    int pv = bi;
    if (!(bv == bv) || pv < 0 || pv >= NOM || ((used >> pv) & 1u)) {
      pv = 0;
      while (pv < NOM - 1 && ((used >> pv) & 1u)) ++pv;
    }
    if (t == 0) {
      perm[c] = pv;
      if (!(bv > 0.0) && err) atomicOr(err, 2);
    }
    used |= 1u << pv;
    const double rp = __drcp_rn(A[cur][pv][c]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = er[u], x = ex[u];
      if (r < 0) continue;
      const double a = A[cur][r][x];
      A[cur ^ 1][r][x] = (r == pv) ? a : ((x == c) ? 0.0 : fma(-(A[cur][r][c] * rp), A[cur][pv][x], a));
    }
    cur ^= 1;
    __syncthreads();
  }
  // inv row x = (pivot row perm[x]) / pivot, then C = E * inv
  for (int e = t; e < NOM * NOM; e += blockDim.x) {
    const int x = e / NOM, c = e % NOM, px = perm[x];
    A[cur ^ 1][x][c] = A[cur][px][NOM + c] / A[cur][px][x];
  }
  __syncthreads();
  for (int e = t; e < NOM * NOM; e += blockDim.x) {
    const int r = e / NOM, c = e % NOM;
    double v = 0.0;
    for (int x = 0; x < NOM; ++x) v = fma(E[r][x], A[cur ^ 1][x][c], v);
    om[e] = v;
  }
  if (t < NOM) om[NOM * NOM + t] = q[t];
}

// Omega tile <- W2_Omega - W1[Omega rows]  (dW: the rows on which W2 differs from W1)
__global__ void k_omega_delta(int m_pad, int n_pad, const int* __restrict__ omega, double* __restrict__ T) {
  XB_PDL_SHORT();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (c >= m_pad) return;
  double* dst = T + (size_t)(m_pad + n_pad + 32 + k) * m_pad + c;
  *dst -= T[(size_t)(m_pad + omega[k]) * m_pad + c];
}
// One warp per state row i, fed by Cb:  Y1 = W1_i Vt^T (cols 64..), Q_i = W1_i dW^T (cols 32..), W1_i.z (col 0);
// Omega rows add dW_k(i) Vt^T (rows n_pad+32+k of Cb):  Y2 = W2_i Vt^T, Z1 = Y1 C,
//   delta_i = W1_i z - Z1 . q - corr_i      (updater.cpp:127-129)
__global__ void __launch_bounds__(128) k_omega_finish(int N, int n_pad, const double* __restrict__ Cb,
                                                      const double* __restrict__ om, const int* __restrict__ omega_inv,
                                                      const double* __restrict__ corr, double* __restrict__ delta,
                                                      double* __restrict__ Zb, double* __restrict__ Yb,
                                                      double* __restrict__ Qb) {
  XB_PDL_SHORT();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const size_t zs = (size_t)(n_pad + 96) * 96;
  const double* ci = Cb + (size_t)i * 96;
  const int ki = omega_inv[i];
  double y1 = 0.0, ql = 0.0, wz = 0.0, y2 = 0.0;
  for (int z = 0; z < CBZ; ++z) {  // fixed-order sum of the split-K partials
    if (lane < NOM) { y1 += ci[z * zs + 64 + lane]; ql += ci[z * zs + 32 + lane]; }
    wz += ci[z * zs];
    if (ki >= 0 && lane < NOM) y2 += Cb[z * zs + (size_t)(n_pad + 32 + ki) * 96 + 64 + lane];
  }
  y2 += y1;
  double zl = 0.0;
#pragma unroll
  for (int k = 0; k < NOM; ++k) {
    const double yk = __shfl_sync(0xffffffffu, y1, k);
    if (lane < NOM) zl = fma(yk, om[k * NOM + lane], zl);
  }
  double zq = (lane < NOM) ? zl * om[NOM * NOM + lane] : 0.0;
  zq = xb_warp_sum(zq);
  Zb[(size_t)i * 32 + lane] = (lane < NOM) ? zl : 0.0;
  Yb[(size_t)i * 32 + lane] = (lane < NOM) ? y2 : 0.0;
  Qb[(size_t)i * 32 + lane] = ql;
  if (lane == 0) delta[i] = wz - zq - (corr ? corr[i] : 0.0);
}

// State::correct (state.cpp:197-249) + correction_total += correction (updater.cpp:140)
__global__ void k_correct(int M, int F, int N, const double* __restrict__ delta, double* __restrict__ xv,
                          double* __restrict__ corr) {
  XB_PDL_SHORT();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N && corr) corr[t] += delta[t];
  if (t < 3) {
    xv[XV_P + t] += delta[t];
    xv[XV_V + t] += delta[3 + t];
    xv[XV_BW + t] += delta[9 + t];
    xv[XV_BA + t] += delta[12 + t];
  }
  if (t < 3 * M) xv[XV_ARR + t] += delta[XB_CORE + t];
  if (t < 3 * F) xv[XV_ARR + 7 * M + t] += delta[XB_CORE + 6 * M + t];
  if (t <= M) {  // t == M: core quaternion; t < M: window pose t
    double* q = (t == M) ? xv + XV_Q : xv + XV_ARR + 3 * M + 4 * t;
    const double* d = (t == M) ? delta + 6 : delta + XB_CORE + 3 * M + 3 * t;
    double dq[4], qo[4];
    xb_small_angle_quat(d, dq);
    xb_qmul(q, dq, qo);
    const double n = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
    if (n > 0.0) { qo[0] /= n; qo[1] /= n; qo[2] /= n; qo[3] /= n; }
    q[0] = qo[0]; q[1] = qo[1]; q[2] = qo[2]; q[3] = qo[3];
  }
}

void launch_dense_prepare(cudaStream_t s, int m, int m_pad, int N, int n_pad, const double* P, const double* H,
                          const double* res, const double* rdiag, const double* corr_total, const int* omega, double* T) {
  gemm_nt(s, N, m, N, 1.0, P, N, H, N, 0.0, T + (size_t)m_pad * m_pad, m_pad);          // P H^T
  gemm_nn(s, m, m, N, 1.0, H, N, T + (size_t)m_pad * m_pad, m_pad, 0.0, T, m_pad);      // H (P H^T)
  launch_sym_lower(s, T, m_pad, 0, m, 0);
  {
    dim3 g((m + 127) / 128, NOM);
    XB_LAUNCH(k_omega_rows_dense, g, 128, 0, s, m, m_pad, n_pad, N, P, H, omega, T);
    count_launch();
  }
  XB_LAUNCH(k_dense_finish, (m_pad + 127) / 128, 128, 0, s, m, m_pad, n_pad, N, H, res, rdiag, corr_total, T);
  count_launch();
}

void launch_omega_rows(cudaStream_t s, const UpdateDims& d, int c0, int nc, const double* P, const double* Lg, int ldr,
                       const int* scols, const double* svals, const int* omega, double* T) {
  if (nc <= 0) return;
  dim3 g((nc + 127) / 128, NOM);
  XB_LAUNCH(k_omega_rows, g, 128, 0, s, d, c0, nc, P, Lg, ldr, scols, svals, omega, T);
  count_launch();
}
void launch_sym_lower(cudaStream_t s, double* T, int ld, int r0, int r1, int c0) {
  if (r1 <= r0) return;
  dim3 b(32, 8), g((r1 - c0 + 31) / 32, (r1 - r0 + 7) / 8);
  XB_LAUNCH(k_sym_lower, g, b, 0, s, T, ld, r0, r1, c0);
  count_launch();
}
void launch_correct(cudaStream_t s, int M, int F, int N, double* T, int m_pad, int n_pad, const double* P,
                    const int* omega, const int* omega_inv, double* om, double* Zb, double* Yb, double* Qb, double* Cb,
                    double* xv, double* corr_total, double* delta_out, int* err) {
  {
    dim3 g((m_pad + 127) / 128, NOM);
    XB_LAUNCH(k_omega_delta, g, 128, 0, s, m_pad, n_pad, omega, T);
    count_launch();
  }
  // Cb[(n_pad + 96) x 96] = [W1 ; aux ; dW ; Vt] * [aux ; dW ; Vt]^T : every dot product the Woodbury step needs
  gemm_nt_splitk(s, n_pad + 96, 96, m_pad, T + (size_t)m_pad * m_pad, m_pad, T + (size_t)(m_pad + n_pad) * m_pad, m_pad, Cb, 96,
                 (size_t)(n_pad + 96) * 96, CBZ);
  XB_LAUNCH(k_omega_small, 1, 256, 0, s, N, n_pad, Cb, P, omega, om, err);
  count_launch();
  XB_LAUNCH(k_omega_finish, (N * 32 + 127) / 128, 128, 0, s, N, n_pad, Cb, om, omega_inv, corr_total, delta_out, Zb, Yb, Qb);
  count_launch();
  launch_apply_delta(s, M, F, N, delta_out, xv, corr_total);
}
// P <- sym(P_j - K (H P_j))   (updater.cpp:153-156), m small
__global__ void __launch_bounds__(256) k_ci_cov(double* __restrict__ P, int N, const double* __restrict__ K,
                                                const double* __restrict__ HP, int m) {
  XB_PDL_SHORT();
  const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= N || j >= N || i > j) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < m; ++k) {
    a = fma(K[(size_t)i * m + k], HP[(size_t)k * N + j], a);
    b = fma(K[(size_t)j * m + k], HP[(size_t)k * N + i], b);
  }
  const double v = 0.5 * ((P[(size_t)i * N + j] - a) + (P[(size_t)j * N + i] - b));
  P[(size_t)i * N + j] = v;
  P[(size_t)j * N + i] = v;
}
void launch_ci_cov(cudaStream_t s, double* P, int N, const double* K, const double* HP, int m) {
  dim3 g((N + 15) / 16, (N + 15) / 16);
  XB_LAUNCH(k_ci_cov, g, 256, 0, s, P, N, K, HP, m);
  count_launch();
}
void launch_apply_delta(cudaStream_t s, int M, int F, int N, const double* delta, double* xv, double* corr_total) {
  XB_LAUNCH(k_correct, (N + 127) / 128, 128, 0, s, M, F, N, delta, xv, corr_total);
  count_launch();
}

// Omega = 15 core states + the 6 states of the newest clone: index table, inverse table and the flags of the 32-row
// tiles that contain an Omega row -- written on the stream (no host staging, no synchronisation).
__global__ void k_set_omega(int* __restrict__ omega, int* __restrict__ omega_inv, int* __restrict__ tileflag, int slot, int M,
                            int n_pad, int nflag) {
  XB_PDL_SHORT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto om = [&](int k) { return k < 15 ? k : (k < 18 ? XB_CORE + 3 * slot + (k - 15) : XB_CORE + 3 * M + 3 * slot + (k - 18)); };
  if (i < 32) omega[i] = i < 21 ? om(i) : 0;
  if (i < n_pad) {
    int inv = -1;
    for (int k = 0; k < 21; ++k) if (om(k) == i) inv = k;
    omega_inv[i] = inv;
  }
  if (i < nflag) {
    int fl = 0;
    for (int k = 0; k < 21; ++k) if (om(k) / 32 == i) fl = 1;
    tileflag[i] = fl;
  }
}
void launch_set_omega(cudaStream_t s, int* omega, int* omega_inv, int* tileflag, int slot, int M, int n_pad, int nflag) {
  XB_LAUNCH(k_set_omega, (n_pad + 255) / 256, 256, 0, s, omega, omega_inv, tileflag, slot, M, n_pad, nflag);
  count_launch();
}

}  // namespace xb
