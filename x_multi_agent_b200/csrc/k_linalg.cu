// Dense fp64 building blocks of the Kalman update (reference: src/x/ekf/updater.cpp:117-141).
//
//   S = H P H^T + R ; K = P H^T S^-1 ; P <- (I - K H) P ; P <- (P + P^T)/2
// is evaluated as   [L ; W ; z^T] = tallchol([S ; P H^T ; r^T])   (W = P H^T L^-T, z = L^-1 r)
//                   P <- (P + P^T)/2 - W W^T ,  delta = W z
// which is the same arithmetic with the explicit inverse replaced by a Cholesky factor.
#include "xb_kernels.h"

namespace xb {

// ------------------------------------------------------------------------------------------------
// GEMM: C[M x N] = alpha * A[M x K] * op(B) + beta * C, row-major, arbitrary leading dimensions.
//   TRANS_B = true : B is [N x K]  (C = A B^T)      TRANS_B = false : B is [K x N]
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ------------------------------------------------------------------------------------------------
template <bool TRANS_B>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, double alpha, const double* __restrict__ A,
                                              int lda, const double* __restrict__ B, int ldb, double beta,
                                              double* __restrict__ C, int ldc) {
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < K; k0 += 16) {
    {  // A tile: 64 rows x 16 k
      const int r = t >> 2, kk = (t & 3) * 4;
      const int gr = m0 + r;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gk = k0 + kk + u;
        As[kk + u][r] = (gr < M && gk < K) ? A[(size_t)gr * lda + gk] : 0.0;
      }
    }
    if (TRANS_B) {
      const int r = t >> 2, kk = (t & 3) * 4;
      const int gr = n0 + r;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gk = k0 + kk + u;
        Bs[kk + u][r] = (gr < N && gk < K) ? B[(size_t)gr * ldb + gk] : 0.0;
      }
    } else {
      const int kk = t >> 4, c = (t & 15) * 4;
      const int gk = k0 + kk;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gc = n0 + c + u;
        Bs[kk][c + u] = (gk < K && gc < N) ? B[(size_t)gk * ldb + gc] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = m0 + ty * 4 + i;
    if (gr >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = n0 + tx * 4 + j;
      if (gc >= N) continue;
      double* p = &C[(size_t)gr * ldc + gc];
      *p = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*p);
    }
  }
}

void gemm_nt(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  k_gemm<true><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  count_launch();
}
void gemm_nn(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  k_gemm<false><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  count_launch();
}

// ------------------------------------------------------------------------------------------------
// Tall tile Cholesky (single launch, dataflow over 32x32 tiles).
//   T is [rows_pad x ld] row-major with ct = cols_pad/32 tile columns and rt = rows_pad/32 tile rows.
//   The top ct x ct tiles hold a symmetric (semi-)definite matrix S (lower part is read); on exit the
//   lower part holds L (S = L L^T, strictly-upper part of the diagonal tiles zeroed) and every tile row
//   below holds X L^-T for the rows X stored there (TRSM), e.g. W = (P H^T) L^-T and z^T = r^T L^-T.
//   One CTA per tile in column-major tile order; tile (i,j) accumulates  T(i,j) - sum_{k<j} T(i,k) T(j,k)^T
//   as soon as the producers publish their flags, then factors (i==j) or solves against L(j,j).
//   Pivots <= piv_tol * (original diagonal) are treated as zero (column zeroed): semi-definite Gram input.
// ------------------------------------------------------------------------------------------------
#define TC 32
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ bool tc_wait(const int* flag, int* err) {
  if (threadIdx.x == 0) {
    long long spins = 0;
    while (ld_acquire(flag) == 0) {
      __nanosleep(20);
      if (++spins > (1ll << 26)) { atomicExch(err, 1); break; }
    }
  }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(64) k_tallchol(double* __restrict__ T, int ld, int rt, int ct, int* __restrict__ flags,
                                                 int* __restrict__ err, double piv_tol) {
  // map block index -> tile (i, j), column-major over the trapezoid
  int b = blockIdx.x;
  int j = 0;
  {
    int rem = b;
    while (rem >= rt - j) { rem -= rt - j; ++j; }
    b = rem;
  }
  const int i = j + b;
  const int t = threadIdx.x;
  const int ty = t >> 3, tx = t & 7;  // 8x8 threads, 4x4 each

  __shared__ double As[TC][TC + 1];
  __shared__ double Bs[TC][TC + 1];
  __shared__ double Cs[TC][TC + 1];
  __shared__ double dorig[TC];

  double acc[4][4];
  const double* Tij = T + (size_t)i * TC * ld + (size_t)j * TC;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = __ldcg(&Tij[(size_t)(ty * 4 + a) * ld + tx * 4 + c]);
  if (i == j && t < TC) dorig[t] = __ldcg(&Tij[(size_t)t * ld + t]);

  for (int k = 0; k < j; ++k) {
    tc_wait(&flags[i * ct + k], err);
    if (i != j) tc_wait(&flags[j * ct + k], err);
    const double* Aik = T + (size_t)i * TC * ld + (size_t)k * TC;
    const double* Bjk = T + (size_t)j * TC * ld + (size_t)k * TC;
    for (int e = t; e < TC * TC; e += 64) {
      const int r = e >> 5, c = e & 31;
      As[c][r] = __ldcg(&Aik[(size_t)r * ld + c]);  // stored k-major
      Bs[c][r] = (i == j) ? As[c][r] : __ldcg(&Bjk[(size_t)r * ld + c]);
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TC; ++kk) {
      double a[4], bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = As[kk][ty * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) bb[u] = Bs[kk][tx * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(-a[u], bb[v], acc[u][v]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) Cs[ty * 4 + a][tx * 4 + c] = acc[a][c];
  __syncthreads();

  double* Tw = T + (size_t)i * TC * ld + (size_t)j * TC;
  if (i == j) {
    // in-place lower Cholesky of Cs
    for (int c = 0; c < TC; ++c) {
      const double piv = Cs[c][c];
      const bool ok = piv > piv_tol * fabs(dorig[c]) && piv > 0.0;
      const double d = ok ? sqrt(piv) : 0.0;
      __syncthreads();
      if (t < TC) {
        if (t == c) Cs[c][c] = d;
        else if (t > c) Cs[t][c] = ok ? Cs[t][c] / d : 0.0;
        else Cs[t][c] = 0.0;  // strictly upper part
      }
      __syncthreads();
      for (int e = t; e < TC * TC; e += 64) {
        const int r = e >> 5, k2 = e & 31;
        if (k2 > c && k2 <= r) Cs[r][k2] = fma(-Cs[r][c], Cs[k2][c], Cs[r][k2]);
      }
      __syncthreads();
    }
  } else {
    // X L^T = C  (L = L(j,j)), column by column
    tc_wait(&flags[j * ct + j], err);
    const double* Ljj = T + (size_t)j * TC * ld + (size_t)j * TC;
    for (int e = t; e < TC * TC; e += 64) {
      const int r = e >> 5, c = e & 31;
      Bs[r][c] = __ldcg(&Ljj[(size_t)r * ld + c]);
    }
    __syncthreads();
    for (int c = 0; c < TC; ++c) {
      const double d = Bs[c][c];
      if (t < TC) Cs[t][c] = (d != 0.0) ? Cs[t][c] / d : 0.0;
      __syncthreads();
      for (int e = t; e < TC * TC; e += 64) {
        const int r = e >> 5, k2 = e & 31;
        if (k2 > c) Cs[r][k2] = fma(-Cs[r][c], Bs[k2][c], Cs[r][k2]);
      }
      __syncthreads();
    }
  }
  for (int e = t; e < TC * TC; e += 64) {
    const int r = e >> 5, c = e & 31;
    Tw[(size_t)r * ld + c] = Cs[r][c];
  }
  __threadfence();
  __syncthreads();
  if (t == 0) st_release(&flags[i * ct + j], 1);
}

void tallchol(cudaStream_t s, double* T, int ld, int rows_pad, int cols_pad, int* flags, int* err, double piv_tol) {
  const int rt = rows_pad / TC, ct = cols_pad / TC;
  const int ntiles = ct * (ct + 1) / 2 + (rt - ct) * ct;
  cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)rt * ct, s);
  k_tallchol<<<ntiles, 64, 0, s>>>(T, ld, rt, ct, flags, err, piv_tol);
  count_launch();
}

// ------------------------------------------------------------------------------------------------
// Covariance downdate (fp64 CUDA-core variant), reference: updater.cpp:131-136 ((I-KH)P, then symmetrise):
//   P_ij <- (P_ij + P_ji)/2 - (W1_i.W2_j + W2_i.W1_j)/2 + (Z_i.Y_j + Y_i.Z_j)/2
// W1 = rows m_pad.. of the tall buffer; W2 = W1 except on the Omega rows (core + newest clone), which come
// from the Omega tile; Z/Y are the 32-wide rank-21 Woodbury factors (k_update.cu).  One CTA per 32x32
// tile pair (I <= J): it owns both P(I,J) and P(J,I), so the update is in place.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_downdate(double* __restrict__ P, int n, const double* __restrict__ T, int m_pad,
                                                 int n_pad, const int* __restrict__ omega_inv,
                                                 const int* __restrict__ tileflag, const double* __restrict__ Zb,
                                                 const double* __restrict__ Yb) {
  const int nt = (n + TC - 1) / TC;
  int b = blockIdx.x, I = 0;
  while (b >= nt - I) { b -= nt - I; ++I; }
  const int J = I + b;
  const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
  __shared__ double As[TC][TC + 1], Bs[TC][TC + 1], As2[TC][TC + 1], Bs2[TC][TC + 1];
  const double* W1 = T + (size_t)m_pad * m_pad;
  const double* W2o = T + (size_t)(m_pad + n_pad + 32) * m_pad;
  const bool two = tileflag[I] | tileflag[J];
  double acc[4][4], acc2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c] = 0.0; acc2[a][c] = 0.0; }
  // K loop over the m_pad columns of W, plus one final 32-wide block holding the Woodbury factors
  for (int k0 = 0; k0 <= m_pad; k0 += TC) {
    const bool last = k0 == m_pad;
    for (int e = t; e < TC * TC; e += 64) {
      const int r = e >> 5, c = e & 31;
      const int gi = I * TC + r, gj = J * TC + r;
      if (!last) {
        const double a1 = gi < n ? W1[(size_t)gi * m_pad + k0 + c] : 0.0;
        const double b1 = gj < n ? W1[(size_t)gj * m_pad + k0 + c] : 0.0;
        As[c][r] = a1;
        Bs[c][r] = b1;
        if (two) {
          const int oi = gi < n ? omega_inv[gi] : -1, oj = gj < n ? omega_inv[gj] : -1;
          As2[c][r] = oi >= 0 ? W2o[(size_t)oi * m_pad + k0 + c] : a1;
          Bs2[c][r] = oj >= 0 ? W2o[(size_t)oj * m_pad + k0 + c] : b1;
        }
      } else {  // acc -= Z_i.Y_j , acc2 -= Y_i.Z_j
        As[c][r] = gi < n ? -Zb[(size_t)gi * 32 + c] : 0.0;
        Bs2[c][r] = gj < n ? Yb[(size_t)gj * 32 + c] : 0.0;
        As2[c][r] = gi < n ? Yb[(size_t)gi * 32 + c] : 0.0;
        Bs[c][r] = gj < n ? -Zb[(size_t)gj * 32 + c] : 0.0;
      }
    }
    __syncthreads();
    if (two || last) {
#pragma unroll 4
      for (int kk = 0; kk < TC; ++kk) {
        double a[4], bb[4], a2[4], b2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; a2[u] = As2[kk][ty * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { bb[u] = Bs[kk][tx * 4 + u]; b2[u] = Bs2[kk][tx * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc[u][v] = fma(a[u], b2[v], acc[u][v]);    // W1_i . W2_j
            acc2[u][v] = fma(a2[u], bb[v], acc2[u][v]); // W2_i . W1_j
          }
      }
    } else {
#pragma unroll 8
      for (int kk = 0; kk < TC; ++kk) {
        double a[4], bb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = As[kk][ty * 4 + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) bb[u] = Bs[kk][tx * 4 + u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], bb[v], acc[u][v]);
      }
    }
    __syncthreads();
    if (!two && k0 + TC == m_pad) {  // symmetric part accumulated once: mirror it before the two-sided tail
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[a][c] = acc[a][c];
    }
  }
  // stage P(I,J) and P(J,I)^T, then write both
  for (int e = t; e < TC * TC; e += 64) {
    const int r = e >> 5, c = e & 31;
    const int gi = I * TC + r, gj = J * TC + c;
    const bool in = gi < n && gj < n;
    As[r][c] = in ? P[(size_t)gi * n + gj] : 0.0;
    Bs[r][c] = in ? P[(size_t)gj * n + gi] : 0.0;  // (J,I) block transposed
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int r = ty * 4 + a, cc = tx * 4 + c;
      const int gi = I * TC + r, gj = J * TC + cc;
      if (gi < n && gj < n) {
        const double v = 0.5 * (As[r][cc] + Bs[r][cc]) - 0.5 * (acc[a][c] + acc2[a][c]);
        P[(size_t)gi * n + gj] = v;
        if (I != J) P[(size_t)gj * n + gi] = v;
      }
    }
}

void downdate_f64(cudaStream_t s, double* P, int n, const double* T, int m_pad, int n_pad, const int* omega_inv,
                  const int* tileflag, const double* Zb, const double* Yb) {
  const int nt = (n + TC - 1) / TC;
  k_downdate<<<nt * (nt + 1) / 2, 64, 0, s>>>(P, n, T, m_pad, n_pad, omega_inv, tileflag, Zb, Yb);
  count_launch();
}

// P <- (P + P^T)/2 only (cov_update path with an empty measurement never reaches here; used by applyCI glue).
__global__ void k_symmetrise(double* __restrict__ P, int n) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i < n && j < n && i < j) {
    const double v = 0.5 * (P[(size_t)i * n + j] + P[(size_t)j * n + i]);
    P[(size_t)i * n + j] = v;
    P[(size_t)j * n + i] = v;
  }
}
void symmetrise(cudaStream_t s, double* P, int n) {
  dim3 g((n + 15) / 16, (n + 15) / 16), b(16, 16);
  k_symmetrise<<<g, b, 0, s>>>(P, n);
  count_launch();
}

// y[r] = sum_k A[r, k] * x[k]   (one warp per row)
__global__ void k_gemv(int rows, int cols, const double* __restrict__ A, int lda, const double* __restrict__ x,
                       double* __restrict__ y) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= rows) return;
  double s = 0.0;
  for (int k = lane; k < cols; k += 32) s = fma(A[(size_t)w * lda + k], x[k], s);
  s = xb_warp_sum(s);
  if (lane == 0) y[w] = s;
}
void gemv(cudaStream_t s, int rows, int cols, const double* A, int lda, const double* x, double* y) {
  if (rows <= 0) return;
  k_gemv<<<(rows * 32 + 127) / 128, 128, 0, s>>>(rows, cols, A, lda, x, y);
  count_launch();
}

// out (cols x rows) = in (rows x cols)^T
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
  __shared__ double tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int k = 0; k < 32; k += 8)
    if (x < cols && y + k < rows) tile[threadIdx.y + k][threadIdx.x] = in[(size_t)(y + k) * cols + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;
  y = blockIdx.x * 32 + threadIdx.y;
  for (int k = 0; k < 32; k += 8)
    if (x < rows && y + k < cols) out[(size_t)(y + k) * rows + x] = tile[threadIdx.x][threadIdx.y + k];
}
void transpose(cudaStream_t s, const double* in, double* out, int rows, int cols) {
  dim3 g((cols + 31) / 32, (rows + 31) / 32), b(32, 8);
  k_transpose<<<g, b, 0, s>>>(in, out, rows, cols);
  count_launch();
}

}  // namespace xb
