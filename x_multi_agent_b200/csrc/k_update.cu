// Assembly of the Kalman-update operands from the device-resident compressed measurement
//   Hc = [ Rg (6M x 6M, on the pose columns) ; SLAM rows (2 per feature, <=15 columns each) ]
// and the state correction.  reference: src/x/vio/vio_updater.cpp:405-423,487-512 (stack + QR compress),
// src/x/ekf/updater.cpp:117-141 (applyUpdate), src/x/ekf/state.cpp:197-249 (State::correct).
//
// Tall buffer T (leading dimension ld = m_pad):
//   rows [0, m_pad)                S = Hc P Hc^T + var I   (identity on the padding)
//   rows [m_pad, m_pad + n_pad)    P Hc^T
//   row  m_pad + n_pad             r_eff^T = (res + Hc * correction_total)^T
#include "xb_kernels.h"

namespace xb {

// PHt[i, ms + 2j + r] = sum_e P[i, col_e] * val[j][r][e]
__global__ void k_pht_slam(UpdateDims d, const double* __restrict__ P, const int* __restrict__ scols,
                           const double* __restrict__ svals, double* __restrict__ T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= d.N || j >= d.nslam) return;
  double a0 = 0.0, a1 = 0.0;
  const double* Pi = P + (size_t)i * d.N;
#pragma unroll
  for (int e = 0; e < 15; ++e) {
    const double p = Pi[scols[15 * j + e]];
    a0 = fma(p, svals[30 * j + e], a0);
    a1 = fma(p, svals[30 * j + 15 + e], a1);
  }
  double* row = T + (size_t)(d.m_pad + i) * d.ld + d.ms + 2 * j;
  row[0] = a0;
  row[1] = a1;
}

// S[ms + 2j + r, c] = sum_e val[j][r][e] * PHt[col_e, c]
__global__ void k_s_slam(UpdateDims d, const int* __restrict__ scols, const double* __restrict__ svals, double* __restrict__ T) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (c >= d.m || j >= d.nslam) return;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int e = 0; e < 15; ++e) {
    const double p = T[(size_t)(d.m_pad + scols[15 * j + e]) * d.ld + c];
    a0 = fma(svals[30 * j + e], p, a0);
    a1 = fma(svals[30 * j + 15 + e], p, a1);
  }
  T[(size_t)(d.ms + 2 * j) * d.ld + c] = a0;
  T[(size_t)(d.ms + 2 * j + 1) * d.ld + c] = a1;
}

// diagonal (+var on real rows, 1 on padding) and the r_eff row
__global__ void k_s_finish(UpdateDims d, const double* __restrict__ Rg, int ldr, const double* __restrict__ zg,
                           const int* __restrict__ scols, const double* __restrict__ svals, const double* __restrict__ sres,
                           const double* __restrict__ corr, double var, double* __restrict__ T) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= d.m_pad) return;
  double* reff = T + (size_t)(d.m_pad + d.n_pad) * d.ld;
  if (a >= d.m) {
    T[(size_t)a * d.ld + a] = 1.0;
    reff[a] = 0.0;
    return;
  }
  T[(size_t)a * d.ld + a] += var;
  double r;
  if (a < d.ms) {
    r = zg[a];
    if (corr)
      for (int b = 0; b < d.ms; ++b) r = fma(Rg[(size_t)a * ldr + b], corr[XB_CORE + b], r);
  } else {
    const int j = (a - d.ms) >> 1, h = (a - d.ms) & 1;
    r = sres[2 * j + h];
    if (corr)
      for (int e = 0; e < 15; ++e) r = fma(svals[30 * j + 15 * h + e], corr[scols[15 * j + e]], r);
  }
  reff[a] = r;
}

void launch_build_pht(cudaStream_t s, const UpdateDims& d, const double* P, const double* Rg, int ldr, const int* scols,
                      const double* svals, double* T) {
  // slab columns: PHt[:, 0:ms] = P[:, pose] * Rg^T
  gemm_nt(s, d.N, d.ms, d.ms, 1.0, P + XB_CORE, d.N, Rg, ldr, 0.0, T + (size_t)d.m_pad * d.ld, d.ld);
  if (d.nslam > 0) {
    dim3 g((d.N + 127) / 128, d.nslam);
    k_pht_slam<<<g, 128, 0, s>>>(d, P, scols, svals, T);
    count_launch();
  }
}

void launch_build_s(cudaStream_t s, const UpdateDims& d, const double* Rg, int ldr, const double* zg, const int* scols,
                    const double* svals, const double* sres, const double* corr_total, double var, double* T) {
  // slab rows: S[0:ms, 0:m] = Rg * PHt[pose rows, 0:m]
  gemm_nn(s, d.ms, d.m, d.ms, 1.0, Rg, ldr, T + (size_t)(d.m_pad + XB_CORE) * d.ld, d.ld, 0.0, T, d.ld);
  if (d.nslam > 0) {
    dim3 g((d.m + 127) / 128, d.nslam);
    k_s_slam<<<g, 128, 0, s>>>(d, scols, svals, T);
    count_launch();
  }
  k_s_finish<<<(d.m_pad + 127) / 128, 128, 0, s>>>(d, Rg, ldr, zg, scols, svals, sres, corr_total, var, T);
  count_launch();
}

// dense-H path: S += diag(rdiag) (+ identity padding), r_eff = res + H corr
__global__ void k_dense_finish(int m, int m_pad, int n_pad, int N, const double* __restrict__ H, const double* __restrict__ res,
                               const double* __restrict__ rdiag, const double* __restrict__ corr, double* __restrict__ T) {
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
void launch_dense_prepare(cudaStream_t s, int m, int m_pad, int N, int n_pad, const double* P, const double* H,
                          const double* res, const double* rdiag, const double* corr_total, double* T) {
  gemm_nt(s, N, m, N, 1.0, P, N, H, N, 0.0, T + (size_t)m_pad * m_pad, m_pad);          // P H^T
  gemm_nn(s, m, m, N, 1.0, H, N, T + (size_t)m_pad * m_pad, m_pad, 0.0, T, m_pad);      // H (P H^T)
  k_dense_finish<<<(m_pad + 127) / 128, 128, 0, s>>>(m, m_pad, n_pad, N, H, res, rdiag, corr_total, T);
  count_launch();
}

// delta = W z - corr_total   (updater.cpp:127-129), one warp per state row
__global__ void k_delta(int N, int m_pad, int n_pad, const double* __restrict__ T, const double* __restrict__ corr,
                        double* __restrict__ delta) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= N) return;
  const double* Wr = T + (size_t)(m_pad + w) * m_pad;
  const double* z = T + (size_t)(m_pad + n_pad) * m_pad;
  double s = 0.0;
  for (int k = lane; k < m_pad; k += 32) s = fma(Wr[k], z[k], s);
  s = xb_warp_sum(s);
  if (lane == 0) delta[w] = s - (corr ? corr[w] : 0.0);
}

// State::correct (state.cpp:197-249) + correction_total += correction (updater.cpp:140)
__global__ void k_correct(int M, int F, int N, const double* __restrict__ delta, double* __restrict__ xv,
                          double* __restrict__ corr) {
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

void launch_correct(cudaStream_t s, int M, int F, int N, const double* T, int m_pad, int n_pad, double* xv,
                    double* corr_total, double* delta_out) {
  k_delta<<<(N * 32 + 127) / 128, 128, 0, s>>>(N, m_pad, n_pad, T, corr_total, delta_out);
  count_launch();
  launch_apply_delta(s, M, F, N, delta_out, xv, corr_total);
}
void launch_apply_delta(cudaStream_t s, int M, int F, int N, const double* delta, double* xv, double* corr_total) {
  k_correct<<<(N + 127) / 128, 128, 0, s>>>(M, F, N, delta, xv, corr_total);
  count_launch();
}

}  // namespace xb
