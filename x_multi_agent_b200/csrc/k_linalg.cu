// Dense fp64 building blocks of the Kalman update (reference: src/x/ekf/updater.cpp:117-141).
//
//   S = H P H^T + R ; K = P H^T S^-1 ; P <- (I - K H) P ; P <- (P + P^T)/2
// is evaluated as   [L ; W ; z^T] = tallchol([S ; P H^T ; r^T])   (W = P H^T L^-T, z = L^-1 r)
//                   P <- (P + P^T)/2 - W W^T ,  delta = W z
// which is the same arithmetic with the explicit inverse replaced by a Cholesky factor.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include <cstdio>
#include <cstdlib>

#include "xb_kernels.h"

namespace xb {

// ------------------------------------------------------------------------------------------------
// GEMM: C[M x N] = alpha * A[M x K] * op(B) + beta * C, row-major, arbitrary leading dimensions.
//   TRANS_B = true : B is [N x K]  (C = A B^T)      TRANS_B = false : B is [K x N]
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// ------------------------------------------------------------------------------------------------
template <bool TRANS_B>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int Kfull, double alpha, const double* __restrict__ A,
                                              int lda, const double* __restrict__ B, int ldb, double beta,
                                              double* __restrict__ C, int ldc, int kchunk, size_t strideC) {
  XB_PDL_SHORT();
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  // split-K: blockIdx.z handles k in [z*kchunk, min(K, (z+1)*kchunk)) and writes its own partial C + z*strideC
  const int kbeg = blockIdx.z * kchunk;
  const int K = min(Kfull, kbeg + kchunk);
  C += (size_t)blockIdx.z * strideC;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // register-prefetched K slabs: the global loads of slab k+1 are in flight while slab k is multiplied
  double ra[4], rb[4];
  const int ar = t >> 2, ak = (t & 3) * 4;      // A tile (and B^T tile): row, first k
  const int bk = t >> 4, bc = (t & 15) * 4;     // B tile (non-transposed): k, first column
  auto gload = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gk = k0 + ak + u;
      ra[u] = (m0 + ar < M && gk < K) ? A[(size_t)(m0 + ar) * lda + gk] : 0.0;
      if (TRANS_B) rb[u] = (n0 + ar < N && gk < K) ? B[(size_t)(n0 + ar) * ldb + gk] : 0.0;
      else rb[u] = (k0 + bk < K && n0 + bc + u < N) ? B[(size_t)(k0 + bk) * ldb + n0 + bc + u] : 0.0;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      As[ak + u][ar] = ra[u];
      if (TRANS_B) Bs[ak + u][ar] = rb[u]; else Bs[bk][bc + u] = rb[u];
    }
  };
  gload(kbeg);
  sstore();
  __syncthreads();
  for (int k0 = kbeg; k0 < K; k0 += 16) {
    if (k0 + 16 < K) gload(k0 + 16);
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
    if (k0 + 16 < K) sstore();
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

// Small-problem variant: 32x32 output tile, 128 threads, 2x4 register micro-tile, K in 32-wide slabs.  The GEMMs of this
// path are tiny (a few hundred rows/columns): with 64x64 tiles most of them occupy 9..54 of the 148 SMs and every CTA
// walks the whole K loop alone; quarter-size tiles spread the same work over 4x more SMs.
template <bool TRANS_B>
__global__ void __launch_bounds__(128) k_gemm32(int M, int N, int Kfull, double alpha, const double* __restrict__ A,
                                                int lda, const double* __restrict__ B, int ldb, double beta,
                                                double* __restrict__ C, int ldc, int kchunk, size_t strideC) {
  XB_PDL_SHORT();
  __shared__ __align__(16) double As[32][32 + 2];  // [k][row]
  __shared__ __align__(16) double Bs[32][32 + 4];  // [k][col]
  const int kbeg = blockIdx.z * kchunk;
  const int K = min(Kfull, kbeg + kchunk);
  C += (size_t)blockIdx.z * strideC;
  const int t = threadIdx.x;
  const int ty = t >> 3, tx = t & 7;  // rows 2*ty.., cols 4*tx..
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  double acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double ra[8], rb[8];
  const int ar = t >> 2, ak = (t & 3) * 8;   // A tile (and B^T tile): row, first k (8 consecutive k)
  const int bk = t >> 2, bc = (t & 3) * 8;   // B tile (non-transposed): k row, first column (8 consecutive columns)
  auto gload = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int gk = k0 + ak + u;
      ra[u] = (m0 + ar < M && gk < K) ? A[(size_t)(m0 + ar) * lda + gk] : 0.0;
      if (TRANS_B) rb[u] = (n0 + ar < N && gk < K) ? B[(size_t)(n0 + ar) * ldb + gk] : 0.0;
      else rb[u] = (k0 + bk < K && n0 + bc + u < N) ? B[(size_t)(k0 + bk) * ldb + n0 + bc + u] : 0.0;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      As[ak + u][ar] = ra[u];
      if (TRANS_B) Bs[ak + u][ar] = rb[u]; else Bs[bk][bc + u] = rb[u];
    }
  };
  gload(kbeg);
  sstore();
  __syncthreads();
  for (int k0 = kbeg; k0 < K; k0 += 32) {
    if (k0 + 32 < K) gload(k0 + 32);
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[kk][ty * 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4 + 2]);
      acc[0][0] = fma(a01.x, b01.x, acc[0][0]); acc[0][1] = fma(a01.x, b01.y, acc[0][1]);
      acc[0][2] = fma(a01.x, b23.x, acc[0][2]); acc[0][3] = fma(a01.x, b23.y, acc[0][3]);
      acc[1][0] = fma(a01.y, b01.x, acc[1][0]); acc[1][1] = fma(a01.y, b01.y, acc[1][1]);
      acc[1][2] = fma(a01.y, b23.x, acc[1][2]); acc[1][3] = fma(a01.y, b23.y, acc[1][3]);
    }
    __syncthreads();
    if (k0 + 32 < K) sstore();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gr = m0 + ty * 2 + i;
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

// ------------------------------------------------------------------------------------------------
// fp64 tensor-core GEMM: mma.sync.m8n8k4.f64 (DMMA).  Measured on B200 (tools/bench_dmma.cu): DMMA and DFMA share the
// same 36.9 TFLOP/s peak, but one DMMA retires 256 FMAs from two operand registers per lane, so a warp needs ~7x fewer
// issue slots and shared-memory loads per flop -- which is what bounds the small, low-occupancy GEMMs of this path
// (ncu on the DFMA tiles: fp64 pipe 12 % active, 66 % of the stalls on shared-memory operands).
//   CTA tile 32 x 64, four warps (2 x 2), warp tile 16 x 32 = 2 x 4 DMMA accumulators; K in 16-wide slabs through a
//   3-stage cp.async ring (8-byte copies: operands such as P + 15 with an odd leading dimension are only 8-byte aligned;
//   out-of-range elements are zero-filled by the copy itself).
//   Fragment layout (PTX ISA, m8n8k4 .f64): A row g = lane/4, col t = lane%4; B row(k) t, col(n) g; C row g, cols 2t, 2t+1.
// Same interface as k_gemm (split-K through blockIdx.z writes partials at C + z*strideC).
// ------------------------------------------------------------------------------------------------
#define MG_BM 32
#define MG_BN 64
#define MG_BK 16
#define MG_ST 3
#define MG_LDK (MG_BK + 4)   // [row][k] tiles: row stride 20 doubles = 4 mod 16 -> conflict-free fragment loads
#define MG_LDN (MG_BN + 4)   // [k][n] tile of a non-transposed B: row stride 68 doubles = 4 mod 16
#define MG_LDM (MG_BM + 4)   // [k][m] tile of a k-major A: row stride 36 doubles = 4 mod 16
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <bool TRANS_B, bool TRANS_A = false>  // TRANS_A: A is given k-major, [K x M] row-major (C = A^T op(B))
__global__ void __launch_bounds__(128) k_gemm_mma(int M, int N, int Kfull, double alpha, const double* __restrict__ A, int lda,
                                                  const double* __restrict__ B, int ldb, double beta, double* __restrict__ C,
                                                  int ldc, int kchunk, size_t strideC) {
  XB_PDL_SHORT();
  __shared__ double As[MG_ST][TRANS_A ? MG_BK * MG_LDM : MG_BM * MG_LDK];
  __shared__ double Bs[MG_ST][TRANS_B ? MG_BN * MG_LDK : MG_BK * MG_LDN];
  const int kbeg = blockIdx.z * kchunk;
  const int K = min(Kfull, kbeg + kchunk);
  C += (size_t)blockIdx.z * strideC;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int m0 = blockIdx.y * MG_BM, n0 = blockIdx.x * MG_BN;
  const int wm0 = (warp >> 1) * 16, wn0 = (warp & 1) * 32;
  const int nslab = (K - kbeg + MG_BK - 1) / MG_BK;
  auto issue = [&](int slab) {
    if (slab < nslab) {
      const int k0 = kbeg + slab * MG_BK;
      double* as = As[slab % MG_ST];
      double* bs = Bs[slab % MG_ST];
      // A tile: 32 rows x 16 k = 512 elements, 4 per thread (k fastest: 16 consecutive threads cover one row)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (TRANS_A) {
          const int e = t + 128 * u, k = e >> 5, r = e & 31;
          const bool v = m0 + r < M && k0 + k < K;
          cp_async8(as + k * MG_LDM + r, v ? A + (size_t)(k0 + k) * lda + m0 + r : A, v);
        } else {
          const int e = t + 128 * u, r = e >> 4, k = e & 15;
          const bool v = m0 + r < M && k0 + k < K;
          cp_async8(as + r * MG_LDK + k, v ? A + (size_t)(m0 + r) * lda + k0 + k : A, v);
        }
      }
      if (TRANS_B) {  // B is [N x K]: 64 rows x 16 k
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = t + 128 * u, r = e >> 4, k = e & 15;
          const bool v = n0 + r < N && k0 + k < K;
          cp_async8(bs + r * MG_LDK + k, v ? B + (size_t)(n0 + r) * ldb + k0 + k : B, v);
        }
      } else {  // B is [K x N]: 16 k rows x 64 columns (n fastest)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = t + 128 * u, k = e >> 6, c = e & 63;
          const bool v = k0 + k < K && n0 + c < N;
          cp_async8(bs + k * MG_LDN + c, v ? B + (size_t)(k0 + k) * ldb + n0 + c : B, v);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll
  for (int s_ = 0; s_ < MG_ST - 1; ++s_) issue(s_);
  for (int slab = 0; slab < nslab; ++slab) {
    asm volatile("cp.async.wait_group %0;" ::"n"(MG_ST - 2) : "memory");
    __syncthreads();
    issue(slab + MG_ST - 1);
    const double* as = As[slab % MG_ST];
    const double* bs = Bs[slab % MG_ST];
#pragma unroll
    for (int kk = 0; kk < MG_BK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        a[i] = TRANS_A ? as[(kk + tg) * MG_LDM + wm0 + 8 * i + g] : as[(wm0 + 8 * i + g) * MG_LDK + kk + tg];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = TRANS_B ? bs[(wn0 + 8 * j + g) * MG_LDK + kk + tg] : bs[(kk + tg) * MG_LDN + wn0 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gr = m0 + wm0 + 8 * i + g;
    if (gr >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = n0 + wn0 + 8 * j + 2 * tg + h;
        if (gc >= N) continue;
        double* p = &C[(size_t)gr * ldc + gc];
        *p = (beta == 0.0) ? alpha * acc[i][j][h] : alpha * acc[i][j][h] + beta * (*p);
      }
  }
}

// XB_GEMM=dfma selects the CUDA-core tiles (k_gemm / k_gemm32) for A/B measurements; default is the DMMA kernel
static const bool g_use_mma = [] { const char* e = getenv("XB_GEMM"); return !(e && e[0] == 'd'); }();
static bool small_gemm(int M, int N, int nz) { return (size_t)((M + 63) / 64) * ((N + 63) / 64) * nz < 128; }

void gemm_nt(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  if (g_use_mma && gemm_nt_tma(s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc)) return;  // large shapes: TMA-staged tiles
  gemm_nt_cpasync(s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
void gemm_nt_cpasync(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
                     double beta, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  if (g_use_mma) {
    dim3 grid((N + MG_BN - 1) / MG_BN, (M + MG_BM - 1) / MG_BM);
    XB_LAUNCH((k_gemm_mma<true>), grid, 128, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  } else if (small_gemm(M, N, 1)) {
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    XB_LAUNCH((k_gemm32<true>), grid, 128, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    XB_LAUNCH((k_gemm<true>), grid, 256, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  }
  count_launch();
}
// C_z = A[:, kz] B[:, kz]^T for nz K-chunks (partials at C + z*strideC; the consumer sums them in a fixed order)
void gemm_nt_splitk(cudaStream_t s, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                    int ldc, size_t strideC, int nz) {
  if (g_use_mma) {
    const int kchunk = ((K + nz - 1) / nz + MG_BK - 1) / MG_BK * MG_BK;
    dim3 grid((N + MG_BN - 1) / MG_BN, (M + MG_BM - 1) / MG_BM, nz);
    XB_LAUNCH((k_gemm_mma<true>), grid, 128, 0, s, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc, kchunk, strideC);
  } else if (small_gemm(M, N, nz)) {
    const int kchunk = ((K + nz - 1) / nz + 31) / 32 * 32;
    dim3 grid((N + 31) / 32, (M + 31) / 32, nz);
    XB_LAUNCH((k_gemm32<true>), grid, 128, 0, s, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc, kchunk, strideC);
  } else {
    const int kchunk = ((K + nz - 1) / nz + 15) / 16 * 16;
    dim3 grid((N + 63) / 64, (M + 63) / 64, nz);
    XB_LAUNCH((k_gemm<true>), grid, 256, 0, s, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc, kchunk, strideC);
  }
  count_launch();
}
// C_z = A[kz, 0:M]^T B[kz, 0:N] for nz K-chunks of row-major [K x M] / [K x N] operands (Gram matrices of tall slabs)
void gemm_tn_splitk(cudaStream_t s, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                    size_t strideC, int nz) {
  const int kchunk = ((K + nz - 1) / nz + MG_BK - 1) / MG_BK * MG_BK;
  dim3 grid((N + MG_BN - 1) / MG_BN, (M + MG_BM - 1) / MG_BM, nz);
  XB_LAUNCH((k_gemm_mma<false, true>), grid, 128, 0, s, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc, kchunk, strideC);
  count_launch();
}
bool gemm_uses_tensor_cores() { return g_use_mma; }
// C = alpha * A^T * B + beta * C with A given k-major ([K x M] row-major) -- tensor-core kernel only
void gemm_tn(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
             double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + MG_BN - 1) / MG_BN, (M + MG_BM - 1) / MG_BM);
  XB_LAUNCH((k_gemm_mma<false, true>), grid, 128, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  count_launch();
}
void gemm_nn(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc) {
  if (M <= 0 || N <= 0) return;
  if (g_use_mma) {
    dim3 grid((N + MG_BN - 1) / MG_BN, (M + MG_BM - 1) / MG_BM);
    XB_LAUNCH((k_gemm_mma<false>), grid, 128, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  } else if (small_gemm(M, N, 1)) {
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    XB_LAUNCH((k_gemm32<false>), grid, 128, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    XB_LAUNCH((k_gemm<false>), grid, 256, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, K, 0);
  }
  count_launch();
}

// ------------------------------------------------------------------------------------------------
// Tall tile Cholesky (single launch, dataflow over 32x32 tiles).
//   T is [rows_pad x ld] row-major with ct = cols_pad/32 tile columns and rt = rows_pad/32 > ct tile rows.
//   The top ct x ct tiles hold a symmetric (semi-)definite matrix S (lower part is read); on exit the
//   lower part holds L (S = L L^T, strictly-upper part of the diagonal tiles zeroed) and every tile row
//   below holds X L^-T for the rows X stored there (TRSM), e.g. W = (P H^T) L^-T and z^T = r^T L^-T.
//   Pivots <= piv_tol * (original diagonal) are treated as zero (column zeroed): semi-definite Gram input.
//
//   The factorisation of an m x m matrix has a serial chain of m pivots; everything here is organised around
//   that chain (measured on B200: a 32x32 fp64 potrf by one warp = 8.1k cycles warm, 22-45k cycles with a
//   cold instruction cache, tools/bench_potrf.cu):
//     block 0 ("critical-path CTA") is persistent and walks the diagonal: for column j it owns tiles
//       D=(j,j) and E=(j+1,j); it applies the last panel update from shared memory, factors D (one warp,
//       rows in registers, left-looking), solves E against it, and publishes both -- consecutive columns
//       hand over through shared memory, not through L2, and its code stays hot in the instruction cache;
//     blocks 1.. ("workers", persistent, round-robin over a dependency-ordered task list) do everything
//       with slack: the TRSM tiles (i,j), i >= j+2 (left-looking accumulation, then solve against L(j,j)),
//       and the partial sums of D/E over the panels k <= j-2 ("pre" tasks).
//   Tiles are handed over through global memory + release/acquire flags; all tile reads bypass L1 (ld.cg).
// ------------------------------------------------------------------------------------------------
#define TC 32
__device__ __forceinline__ void dmma884_t(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// one thread polls with acquire loads; the CTA / group barrier that follows extends the ordering to the other threads.
// Publishing side: tile stores by all threads, barrier, then ONE st.release by one thread (release is cumulative over the
// barrier: no separate __threadfence, which costs a second ~0.5 us round trip per hand-over).
// A flag is "set" when it holds the epoch of the current launch (a process-wide launch counter): no memset between
// launches, stale values of earlier launches never match.
template <bool BACKOFF = false>
__device__ __forceinline__ void flag_spin(const int* flag, int* err, int epoch) {
  long long spins = 0;
  while (ld_acquire(flag) != epoch) {   // every poll is an acquire: no second round trip once the flag has flipped
    if (BACKOFF) __nanosleep(40);       // workers: leave the issue slots of the SM to whoever shares it (the chain CTA)
    if (++spins > (1ll << 22)) { atomicExch(err, 1); break; }
  }
}

// X L^T = C for the 32 rows of C (lane = row, registers), L and its reciprocal diagonal in shared memory.
__device__ __forceinline__ void warp_trsm32(double (*Cs)[TC + 1], double (*Ls)[TC + 1], const double* rd, int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double x = a[c] * rd[c];
    a[c] = x;
#pragma unroll
    for (int k = c + 1; k < TC; ++k) a[k] = fma(-x, Ls[k][c], a[k]);
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
// same solve, the result is also stored k-major (Wt[k][row]) for the tensor-core products that consume it next
__device__ __forceinline__ void warp_trsm32_t(double (*Cs)[TC + 1], double (*Ls)[TC + 1], const double* rd, double (*Wt)[TC + 1],
                                              int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double x = a[c] * rd[c];
    a[c] = x;
#pragma unroll
    for (int k = c + 1; k < TC; ++k) a[k] = fma(-x, Ls[k][c], a[k]);
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) { Cs[lane][k] = a[k]; Wt[k][lane] = a[k]; }
}
__device__ __forceinline__ void bar_group(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 32x32 lower Cholesky by one warp, LEFT-looking and blocked by 4 columns, in place: lane r keeps row r in registers and
// mirrors its finished entries in Cs (read back by the other lanes as broadcasts).  Per block step: the four dot
// products of the row against the finished rows c0..c0+3 (8 independent chains), ten shuffles fetch the updated 4x4
// diagonal block, every lane factors it redundantly (the serial part: 4 dependent rsqrt), then solves its own four
// entries.  One shuffle round / one __syncwarp / one shared-memory round trip per FOUR pivots instead of per pivot:
// tools/bench_potrf4.cu measures it against warp_potrf32 (same pivot guard; rd = reciprocal diagonal, 0 if skipped).
__device__ __forceinline__ void warp_potrf32_b4(double (*Cs)[TC + 1], const double* dorig, double piv_tol, double* rd, int lane) {
  // fully unrolled with the row in registers; a compact-loop variant with the row in shared memory measured 8.9k cycles
  // against 6.9k for this one in isolation (tools/bench_potrf4.cu) and 60 against 55 us per 6-column launch in the kernel
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c0 = 0; c0 < TC; c0 += 4) {
    double v0 = a[c0], v1 = a[c0 + 1], v2 = a[c0 + 2], v3 = a[c0 + 3], w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
    for (int k = 0; k < c0; k += 2) {
      const double ak = a[k], ak1 = a[k + 1];
      v0 = fma(-ak, Cs[c0][k], v0); v1 = fma(-ak, Cs[c0 + 1][k], v1); v2 = fma(-ak, Cs[c0 + 2][k], v2); v3 = fma(-ak, Cs[c0 + 3][k], v3);
      w0 = fma(-ak1, Cs[c0][k + 1], w0); w1 = fma(-ak1, Cs[c0 + 1][k + 1], w1); w2 = fma(-ak1, Cs[c0 + 2][k + 1], w2);
      w3 = fma(-ak1, Cs[c0 + 3][k + 1], w3);
    }
    v0 += w0; v1 += w1; v2 += w2; v3 += w3;
    const unsigned FULL = 0xffffffffu;
    const double a00 = __shfl_sync(FULL, v0, c0), a10 = __shfl_sync(FULL, v0, c0 + 1), a11 = __shfl_sync(FULL, v1, c0 + 1),
                 a20 = __shfl_sync(FULL, v0, c0 + 2), a21 = __shfl_sync(FULL, v1, c0 + 2), a22 = __shfl_sync(FULL, v2, c0 + 2),
                 a30 = __shfl_sync(FULL, v0, c0 + 3), a31 = __shfl_sync(FULL, v1, c0 + 3), a32 = __shfl_sync(FULL, v2, c0 + 3),
                 a33 = __shfl_sync(FULL, v3, c0 + 3);
    const double t0 = piv_tol * fabs(dorig[c0]), t1 = piv_tol * fabs(dorig[c0 + 1]), t2 = piv_tol * fabs(dorig[c0 + 2]),
                 t3 = piv_tol * fabs(dorig[c0 + 3]);
    const double q0 = rsqrt(a00);
    const double r0 = (a00 > t0 && a00 > 0.0) ? q0 : 0.0;
    const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
    const double p1 = fma(-l10, l10, a11);
    const double q1 = rsqrt(p1);
    const double r1 = (p1 > t1 && p1 > 0.0) ? q1 : 0.0;
    const double l21 = fma(-l20, l10, a21) * r1, l31 = fma(-l30, l10, a31) * r1;
    const double p2 = fma(-l21, l21, fma(-l20, l20, a22));
    const double q2 = rsqrt(p2);
    const double r2 = (p2 > t2 && p2 > 0.0) ? q2 : 0.0;
    const double l32 = fma(-l31, l21, fma(-l30, l20, a32)) * r2;
    const double p3 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
    const double q3 = rsqrt(p3);
    const double r3 = (p3 > t3 && p3 > 0.0) ? q3 : 0.0;
    // own row: for the lanes of the block itself the same formulas reproduce l_ij (j < i) and l_ii = p_i r_i
    double x0 = v0 * r0;
    double x1 = fma(-x0, l10, v1) * r1;
    double x2 = fma(-x1, l21, fma(-x0, l20, v2)) * r2;
    double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, v3))) * r3;
    if (lane < c0) x0 = 0.0;
    if (lane < c0 + 1) x1 = 0.0;
    if (lane < c0 + 2) x2 = 0.0;
    if (lane < c0 + 3) x3 = 0.0;
    a[c0] = x0; a[c0 + 1] = x1; a[c0 + 2] = x2; a[c0 + 3] = x3;
    Cs[lane][c0] = x0; Cs[lane][c0 + 1] = x1; Cs[lane][c0 + 2] = x2; Cs[lane][c0 + 3] = x3;
    if (lane == 0) { rd[c0] = r0; rd[c0 + 1] = r1; rd[c0 + 2] = r2; rd[c0 + 3] = r3; }
    __syncwarp();
  }
}
// C(32x32, smem) -= A(32x32) * B(32x32)^T with A, B stored k-major (At[k][r], Bt[k][c]), on the fp64 tensor cores:
// the calling group of nthreads = 64 or 128 threads (t = index within the group) splits C into 16x16 quadrants, one or two
// per warp, 2x2 m8n8k4 accumulators each: 32 DMMAs per quadrant (~0.3 us) instead of 512 DFMAs + 256 shared loads per
// thread (~1.3 us) -- this product sits between two pivots of the critical-path CTA and inside every worker task.
__device__ __forceinline__ void tile_gemm_sub(double (*Cs)[TC + 1], double (*At)[TC + 1], double (*Bt)[TC + 1], int t,
                                              int nthreads) {
  const int lane = t & 31, w = t >> 5, nw = nthreads >> 5;
  const int g = lane >> 2, tg = lane & 3;
  for (int q = w; q < 4; q += nw) {
    const int r0 = (q >> 1) * 16, c0 = (q & 1) * 16;
    double acc[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        acc[i][j][0] = Cs[r0 + 8 * i + g][c0 + 8 * j + 2 * tg];
        acc[i][j][1] = Cs[r0 + 8 * i + g][c0 + 8 * j + 2 * tg + 1];
      }
#pragma unroll
    for (int kk = 0; kk < TC; kk += 4) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = -At[kk + tg][r0 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bt[kk + tg][c0 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884_t(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        Cs[r0 + 8 * i + g][c0 + 8 * j + 2 * tg] = acc[i][j][0];
        Cs[r0 + 8 * i + g][c0 + 8 * j + 2 * tg + 1] = acc[i][j][1];
      }
  }
}
__device__ __forceinline__ void tile_load(double (*S)[TC + 1], const double* g, int ld, int t, int nthreads, bool transpose) {
  for (int e = t; e < TC * TC; e += nthreads) {
    const int r = e >> 5, c = e & 31;
    const double v = __ldcg(&g[(size_t)r * ld + c]);
    if (transpose) S[c][r] = v; else S[r][c] = v;
  }
}
__device__ __forceinline__ void tile_store(double* g, int ld, double (*S)[TC + 1], int t, int nthreads) {
  for (int e = t; e < TC * TC; e += nthreads) {
    const int r = e >> 5, c = e & 31;
    __stcg(&g[(size_t)r * ld + c], S[r][c]);
  }
}

#define CP_THREADS 128
// Register-staged tile: 128 threads x 8 elements.  fetch = ld.cg from global (row-major tile), stash = store into a shared
// tile, transposed (k-major) -- the global-load latency of panel k+1 is hidden behind the tensor-core update with panel k.
__device__ __forceinline__ void tile_fetch(double (&r)[8], const double* g, int ld, int t) {
#pragma unroll
  for (int u = 0; u < 8; ++u) { const int e = t + CP_THREADS * u; r[u] = __ldcg(&g[(size_t)(e >> 5) * ld + (e & 31)]); }
}
__device__ __forceinline__ void tile_stash_t(double (*S)[TC + 1], const double (&r)[8], int t) {
#pragma unroll
  for (int u = 0; u < 8; ++u) { const int e = t + CP_THREADS * u; S[e & 31][e >> 5] = r[u]; }
}
// Column range / row-skip description of one launch.  A factorisation may be split in two launches so that the first
// tile columns are factored while the rows of the remaining columns are still being produced (xb_api.cu overlaps
// the SLAM-row part of the Kalman update with the MSCKF track pipeline this way):
//   launch 1: columns [0, c1), tile rows [c1, ct) skipped (their entries do not exist yet)
//   caller:   L21 = rows [c1, ct) of the factor on the columns [0, c1) (for the Kalman update a plain GEMM,
//             k_update.cu:k_wsym), then the Schur complement T[c1:, c1:] -= T[c1:, :c1] L21^T (GEMM)
//   launch 2: a plain factorisation of the sub-buffer T[c1:, c1:].
// A plain factorisation is {0, ct, 0, 0}.
struct CholRange {
  int jstart, jend;    // tile columns factored by this launch
  int skip0, skip1;    // tile rows excluded from the TRSM work of this launch (skip0 == jend or empty)
};
// TRSM tile (i, jcol): left-looking accumulation over k < jcol, then the solve against L(jcol, jcol)
__device__ __forceinline__ void worker_trsm(double* __restrict__ T, int ld, int ct, int i, int jcol, int* ready, int* err,
                                            double (*Ds)[TC + 1], double (*Ws)[TC + 1], double (*Ls)[TC + 1], double* rd,
                                            long long* trace, int epoch) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long tr0 = trace ? gtimer() : 0;
  double* gC = T + (size_t)i * TC * ld + (size_t)jcol * TC;
  tile_load(Ds, gC, ld, t, CP_THREADS, false);
  __syncthreads();
  if (jcol > 0) {
    // left-looking accumulation over the panels k < jcol, software-pipelined: while panel k is multiplied, the tiles of
    // panel k+1 are already on their way into registers if their flags are set (they almost always are: only the last
    // panel is fresh); otherwise the fetch falls back to the blocking wait after the product.
    double ra[8], rb[8];
    auto fetch_blocking = [&](int k) {
      if (t == 0) { flag_spin<true>(&ready[i * ct + k], err, epoch); flag_spin<true>(&ready[jcol * ct + k], err, epoch); }
      __syncthreads();
      tile_fetch(ra, T + (size_t)i * TC * ld + (size_t)k * TC, ld, t);
      tile_fetch(rb, T + (size_t)jcol * TC * ld + (size_t)k * TC, ld, t);
    };
    fetch_blocking(0);
    for (int k = 0; k < jcol; ++k) {
      tile_stash_t(Ws, ra, t);
      tile_stash_t(Ls, rb, t);
      int rdy = 0;
      if (k + 1 < jcol) rdy = ld_acquire(&ready[i * ct + k + 1]) == epoch && ld_acquire(&ready[jcol * ct + k + 1]) == epoch;
      const int all = __syncthreads_and(rdy);
      if (all) {
        tile_fetch(ra, T + (size_t)i * TC * ld + (size_t)(k + 1) * TC, ld, t);
        tile_fetch(rb, T + (size_t)jcol * TC * ld + (size_t)(k + 1) * TC, ld, t);
      }
      tile_gemm_sub(Ds, Ws, Ls, t, CP_THREADS);
      __syncthreads();
      if (!all && k + 1 < jcol) fetch_blocking(k + 1);
    }
  }
  const long long tr1 = trace ? gtimer() : 0;
  if (t == 0) flag_spin<true>(&ready[jcol * ct + jcol], err, epoch);
  __syncthreads();
  tile_load(Ls, T + (size_t)jcol * TC * ld + (size_t)jcol * TC, ld, t, CP_THREADS, false);
  __syncthreads();
  if (t < TC) { const double d = Ls[t][t]; rd[t] = d != 0.0 ? 1.0 / d : 0.0; }
  __syncthreads();
  if (warp == 0) warp_trsm32(Ds, Ls, rd, lane);
  __syncthreads();
  const long long tr2 = trace ? gtimer() : 0;
  tile_store(gC, ld, Ds, t, CP_THREADS);
  __syncthreads();
  if (t == 0) {
    st_release(&ready[i * ct + jcol], epoch);
    if (trace && i < ct + 2 && i >= jcol + 2) {
      long long* o = trace + 10 * (size_t)(ct + jcol * 2 + (i - jcol - 2) % 2);
      o[0] = i; o[1] = jcol; o[2] = tr0; o[3] = tr1; o[4] = tr2; o[5] = gtimer();
    }
  }
}
// flags: ready[i*ct + j] (L tile published), pre[rt*ct + j] (partial sums of D_j/E_j over the earlier panels published)
//
// Critical-path CTA, per tile column j (warp 0 = the chain, warps 1-3 = what can be taken off it):
//   (a)  all   : wait for pre(j) (partial sums over k <= j-2, done by a worker a whole column ago); D_j, E_j = (j+1, j)
//                into shared memory through registers (all loads of a thread in flight at once);
//                D -= Wt Wt^T on the tensor cores (Wt = L(j, j-1), the k-major copy the previous solve left behind)
//   (b)  warp 0: potrf(D)             | warps 1-3: wait for L(j+1, j-1) from a worker, E -= L(j+1, j-1) Wt^T (tensor cores)
//   (c)  warp 0: E <- E L_jj^-T, also stored k-major into Wt
//                                     | warps 1-3: publish L_jj
//   (d)  all   : publish E
// Measured alternatives (tools/chol_trace.py): prefetching D_{j+1}/E_{j+1} during (b) with the pre tasks moved one column
// earlier makes the chain wait for tiles the workers have only just been able to start (a TRSM tile reaches the consumer
// ~5-7 us after its diagonal tile is published: flag, tile load, solve, store, flag, tile load are six dependent L2 round
// trips): 8.9 and 14 us per column against 10 us before and the figure in DESIGN.md for this schedule.
typedef double (*Tile)[TC + 1];
#define TC_TILE_DOUBLES (TC * (TC + 1))
#define TC_SMEM_BYTES ((5 * TC_TILE_DOUBLES + 2 * TC) * sizeof(double))
__device__ __forceinline__ void tile_fetch96(double (&r)[11], const double* g, int ld, int t) {  // t = 0..95
#pragma unroll
  for (int u = 0; u < 11; ++u) { const int e = t + 96 * u; r[u] = e < TC * TC ? __ldcg(&g[(size_t)(e >> 5) * ld + (e & 31)]) : 0.0; }
}
__device__ __forceinline__ void tile_stash(double (*S)[TC + 1], const double (&r)[8], int t) {
#pragma unroll
  for (int u = 0; u < 8; ++u) { const int e = t + CP_THREADS * u; S[e >> 5][e & 31] = r[u]; }
}
__global__ void __launch_bounds__(CP_THREADS, 4) k_tallchol(double* __restrict__ T, int ld, int rt, int ct, CholRange cr,
                                                         int* __restrict__ flags, int* __restrict__ err, double piv_tol,
                                                         const double* __restrict__ diag0, long long* __restrict__ trace,
                                                         int epoch) {
  XB_PDL_LONG();
  extern __shared__ double tc_smem[];
  Tile Ds = (Tile)tc_smem, Ws = (Tile)(tc_smem + 2 * TC_TILE_DOUBLES), Ls = (Tile)(tc_smem + 3 * TC_TILE_DOUBLES);
  Tile Es[2] = {(Tile)(tc_smem + 1 * TC_TILE_DOUBLES), (Tile)(tc_smem + 4 * TC_TILE_DOUBLES)};
  double* dorig = tc_smem + 5 * TC_TILE_DOUBLES;
  double* rd = dorig + TC;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int* ready = flags;
  int* pre = flags + (size_t)rt * ct;
  const int jstart = cr.jstart, jend = cr.jend;

  if (blockIdx.x == 0) {
    // ------------------------------------------------------------------ critical-path CTA
    int eb = 0;
    for (int j = jstart; j < jend; ++j) {
      const long long tr0 = trace ? gtimer() : 0;
      const bool first = j == jstart;
      const bool has_e = !(j + 1 >= cr.skip0 && j + 1 < cr.skip1);  // E = (j+1, j) exists in this launch
      double* gD = T + (size_t)j * TC * ld + (size_t)j * TC;
      double* gE = T + (size_t)(j + 1) * TC * ld + (size_t)j * TC;
      // ---- (a)
      if (j >= 2) {
        if (t == 0) flag_spin(&pre[j], err, epoch);
        __syncthreads();
      }
      {
        double ra[8], rb[8];
        tile_fetch(ra, gD, ld, t);
        if (has_e) tile_fetch(rb, gE, ld, t);
        if (t < TC) dorig[t] = diag0 ? diag0[j * TC + t] : 0.0;
        tile_stash(Ds, ra, t);
        if (has_e) tile_stash(Es[eb], rb, t);
      }
      __syncthreads();
      if (trace && t == 0) trace[10 * (size_t)j + 9] = gtimer();
      if (!first) {
        tile_gemm_sub(Ds, Ws, Ws, t, CP_THREADS);  // D -= L(j, j-1) L(j, j-1)^T
        __syncthreads();
      }
      if (trace && t == 0) trace[10 * (size_t)j + 3] = gtimer();
      // ---- (b)
      if (warp == 0) {
        warp_potrf32_b4(Ds, dorig, piv_tol, rd, lane);
        if (trace && lane == 0) trace[10 * (size_t)j + 6] = gtimer();
      } else if (!first && has_e) {
        const int tb = t - 32;
        if (tb == 0) flag_spin(&ready[(j + 1) * ct + (j - 1)], err, epoch);
        bar_group(1, 96);
        double rg[11];
        tile_fetch96(rg, T + (size_t)(j + 1) * TC * ld + (size_t)(j - 1) * TC, ld, tb);
#pragma unroll
        for (int u = 0; u < 11; ++u) { const int e = tb + 96 * u; if (e < TC * TC) Ls[e & 31][e >> 5] = rg[u]; }
        bar_group(1, 96);
        tile_gemm_sub(Es[eb], Ls, Ws, tb, 96);   // E -= L(j+1, j-1) L(j, j-1)^T
        if (trace && tb == 0) trace[10 * (size_t)j + 7] = gtimer();
      }
      __syncthreads();
      // ---- (c)
      if (warp == 0) {
        if (has_e) warp_trsm32_t(Es[eb], Ds, rd, Ws, lane);
        if (trace && lane == 0) trace[10 * (size_t)j + 8] = gtimer();
      } else {
        const int tb = t - 32;
        tile_store(gD, ld, Ds, tb, 96);
        bar_group(1, 96);
        if (tb == 0) { st_release(&ready[j * ct + j], epoch); if (trace) trace[10 * (size_t)j + 4] = gtimer(); }
      }
      __syncthreads();
      // ---- (d)
      if (has_e) {
        tile_store(gE, ld, Es[eb], t, CP_THREADS);
        __syncthreads();
        if (t == 0) st_release(&ready[(j + 1) * ct + j], epoch);
      }
      if (trace && t == 0) {
        long long* o = trace + 10 * (size_t)j;
        o[0] = j; o[1] = j; o[2] = tr0; o[5] = gtimer();
      }
      eb ^= 1;
    }
    return;
  }

  // -------------------------------------------------------------------- workers
  // task list, dependency-ordered: for j = jstart..jend-1: [pre(j+1) if j+1 >= 2 and j+1 < jend; it needs only
  // columns <= j-1], then the TRSM tiles (i, j), i = j+2..rt-1 without the skipped rows
  const int nworkers = gridDim.x - 1;
  int task = blockIdx.x - 1;
  int jcol = jstart, base = 0;
  while (true) {
    int jp = -1, kmax = -1, trsm_i = -1, trsm_j = jcol;
    {
      // advance (jcol, base) so that task falls into column jcol's segment
      int seg, has_pre, s0, slen;
      for (;;) {
        if (jcol >= jend) return;
        has_pre = (jcol + 1 >= 2 && jcol + 1 < jend) ? 1 : 0;
        s0 = max(cr.skip0, jcol + 2);
        slen = max(0, cr.skip1 - s0);
        seg = has_pre + (rt - jcol - 2) - slen;
        if (task < base + seg) break;
        base += seg;
        ++jcol;
      }
      const int local = task - base;
      if (has_pre && local == 0) {
        jp = jcol + 1;
        kmax = jp - 2;
      } else {
        trsm_i = jcol + 2 + (local - has_pre);
        if (trsm_i >= s0) trsm_i += slen;
        trsm_j = jcol;
      }
    }
    if (jp >= 0) {
      // ---- pre(jp): D_jp -= sum_{k<=kmax} L(jp,k) L(jp,k)^T ; E_jp -= sum L(jp+1,k) L(jp,k)^T ; in place
      const bool has_e = !(jp + 1 >= cr.skip0 && jp + 1 < cr.skip1);
      double* gD = T + (size_t)jp * TC * ld + (size_t)jp * TC;
      double* gE = T + (size_t)(jp + 1) * TC * ld + (size_t)jp * TC;
      tile_load(Ds, gD, ld, t, CP_THREADS, false);
      if (has_e) tile_load(Es[0], gE, ld, t, CP_THREADS, false);
      __syncthreads();
      if (kmax >= 0) {  // same software pipeline as worker_trsm: panel k+1 in flight while panel k is multiplied
        double ra[8], rb[8];
        auto fetch_blocking = [&](int k) {
          if (t == 0) { flag_spin<true>(&ready[jp * ct + k], err, epoch); if (has_e) flag_spin<true>(&ready[(jp + 1) * ct + k], err, epoch); }
          __syncthreads();
          tile_fetch(ra, T + (size_t)jp * TC * ld + (size_t)k * TC, ld, t);
          if (has_e) tile_fetch(rb, T + (size_t)(jp + 1) * TC * ld + (size_t)k * TC, ld, t);
        };
        fetch_blocking(0);
        for (int k = 0; k <= kmax; ++k) {
          tile_stash_t(Ws, ra, t);                 // L(jp,k) k-major
          if (has_e) tile_stash_t(Ls, rb, t);      // L(jp+1,k) k-major
          int rdy = 0;
          if (k + 1 <= kmax)
            rdy = ld_acquire(&ready[jp * ct + k + 1]) == epoch && (!has_e || ld_acquire(&ready[(jp + 1) * ct + k + 1]) == epoch);
          const int all = __syncthreads_and(rdy);
          if (all) {
            tile_fetch(ra, T + (size_t)jp * TC * ld + (size_t)(k + 1) * TC, ld, t);
            if (has_e) tile_fetch(rb, T + (size_t)(jp + 1) * TC * ld + (size_t)(k + 1) * TC, ld, t);
          }
          if (t < 64) tile_gemm_sub(Ds, Ws, Ws, t, 64);
          else if (has_e) tile_gemm_sub(Es[0], Ls, Ws, t - 64, 64);
          __syncthreads();
          if (!all && k + 1 <= kmax) fetch_blocking(k + 1);
        }
      }
      tile_store(gD, ld, Ds, t, CP_THREADS);
      if (has_e) tile_store(gE, ld, Es[0], t, CP_THREADS);
      __syncthreads();
      if (t == 0) { st_release(&pre[jp], epoch); }
    } else {
      worker_trsm(T, ld, ct, trsm_i, trsm_j, ready, err, Ds, Ws, Ls, rd, trace, epoch);
    }
    __syncthreads();
    task += nworkers;
  }
}

static int tallchol_max_ctas() {
  static int max_ctas = 0;
  if (!max_ctas) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_tallchol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tallchol, CP_THREADS, TC_SMEM_BYTES);
    max_ctas = sms * (per_sm > 0 ? per_sm : 1);  // all CTAs of a launch must be co-resident (persistent dataflow)
  }
  return max_ctas;
}
// share: fraction denominator of the device this launch may occupy (2 = half of the co-resident CTA slots) so that
// dataflow launches running concurrently on different streams can always be resident together.
void tallchol_range(cudaStream_t s, double* T, int ld, int rows_pad, int cols_pad, int jstart_cols, int jend_cols,
                    int phase, int* flags, int* err, double piv_tol, const double* diag0, long long* trace, int share) {
  const int rt = rows_pad / TC, ct = cols_pad / TC;
  CholRange cr{jstart_cols / TC, jend_cols / TC, 0, 0};
  if (phase == 1) { cr.skip0 = cr.jend; cr.skip1 = ct; }
  int ntasks = 0;
  for (int j = cr.jstart; j < cr.jend; ++j) {
    const int s0 = std::max(cr.skip0, j + 2), slen = std::max(0, cr.skip1 - s0);
    ntasks += ((j + 1 >= 2 && j + 1 < cr.jend) ? 1 : 0) + (rt - j - 2) - slen;
  }
  const int max_workers = std::max(1, tallchol_max_ctas() / std::max(1, share) - 1);
  const int nworkers = ntasks < max_workers ? (ntasks > 0 ? ntasks : 1) : max_workers;
  static std::atomic<unsigned> g_epoch{0};
  unsigned ep = ++g_epoch;
  if ((int)ep == 0) ep = ++g_epoch;  // 0 is the value of a freshly allocated flag buffer
  // Cooperative launch: the CTAs of this kernel wait on one another (release/acquire flags), so they must all be
  // resident at once.  The grid is sized from the occupancy query above; the cooperative launch makes the scheduler
  // place it all-or-nothing, which also covers dataflow launches of OTHER filters on the same GPU (two partially
  // resident grids could otherwise wait for each other until the spin limit).  XB_NO_COOP=1: plain launch (A/B).
  static const bool coop = !getenv("XB_NO_COOP");
  int epi = (int)ep;
  if (coop) {
    void* args[] = {&T, &ld, (void*)&rt, (void*)&ct, &cr, &flags, &err, &piv_tol, &diag0, &trace, &epi};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_tallchol, dim3(1 + nworkers), dim3(CP_THREADS), args, TC_SMEM_BYTES, s);
    if (e != cudaSuccess) {   // too large to be co-resident: must not happen (grid sized from the occupancy query)
      fprintf(stderr, "xb200: cooperative launch of k_tallchol failed: %s\n", cudaGetErrorString(e));
      cudaMemsetAsync(err, 0xff, sizeof(int), s);
    }
  } else {
    k_tallchol<<<1 + nworkers, CP_THREADS, TC_SMEM_BYTES, s>>>(T, ld, rt, ct, cr, flags, err, piv_tol, diag0, trace, epi);
  }
  count_launch();
}
void tallchol(cudaStream_t s, double* T, int ld, int rows_pad, int cols_pad, int* flags, int* err, double piv_tol,
              const double* diag0, long long* trace) {
  tallchol_range(s, T, ld, rows_pad, cols_pad, 0, cols_pad, 0, flags, err, piv_tol, diag0, trace, 1);
}

// ------------------------------------------------------------------------------------------------
// Covariance downdate (fp64 CUDA-core variant), reference: updater.cpp:131-136 ((I-KH)P, then symmetrise):
//   P_ij <- (P_ij + P_ji)/2 - (W1_i.W2_j + W2_i.W1_j)/2 + (Z_i.Y_j + Y_i.Z_j)/2
// W1 = rows m_pad.. of the tall buffer; W2 = W1 except on the Omega rows (core + newest clone), which come
// from the Omega tile; Z/Y are the 32-wide rank-21 Woodbury factors (k_update.cu).  One CTA per 32x32
// tile pair (I <= J): it owns both P(I,J) and P(J,I), so the update is in place.
// ------------------------------------------------------------------------------------------------
// 64x64 output tile per CTA (256 threads, 4x4 register micro-tile), K in 16-wide slabs with register prefetch.
//   P_ij <- (P_ij + P_ji)/2 - W1_i.W1_j - (Q_i[k(j)] + Q_j[k(i)])/2 + (Z_i.Y_j + Y_i.Z_j)/2
// where Q_i[k] = W1_i . (W2 - W1)_{Omega_k} carries the rows on which W2 differs from W1 (k(j) = position of j in
// Omega, absent for every other j) and Z, Y are the 32-wide rank-21 Woodbury factors -- so the N^2 m main loop is
// the same symmetric product for every tile.
#define DT 64
#define DK 16
#define DKM 32  // K slab of the main loop
__global__ void __launch_bounds__(256) k_downdate(double* __restrict__ P, int n, const double* __restrict__ T, int m_pad,
                                                  int n_pad, const int* __restrict__ omega_inv, const double* __restrict__ Zb,
                                                  const double* __restrict__ Yb, const double* __restrict__ Qb) {
  XB_PDL_LONG();
  const int nt = (n + DT - 1) / DT;
  int b = blockIdx.x, I = 0;
  while (b >= nt - I) { b -= nt - I; ++I; }
  const int J = I + b;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  __shared__ __align__(16) double sm[4][DK][DT + 4];
  double (*As)[DT + 4] = sm[0];
  double (*Bs)[DT + 4] = sm[1];
  double (*As2)[DT + 4] = sm[2];
  double (*Bs2)[DT + 4] = sm[3];
  // main loop views: two 32 x 68 slabs over the same storage
  double (*Am)[DT + 4] = reinterpret_cast<double (*)[DT + 4]>(&sm[0][0][0]);
  double (*Bm)[DT + 4] = reinterpret_cast<double (*)[DT + 4]>(&sm[2][0][0]);
  const double* W1 = T + (size_t)m_pad * m_pad;
  // loader mapping (main loop): thread -> row (t >> 2) of the 64-row block, 8 consecutive k at (t & 3) * 8
  const int lr = t >> 2, lk8 = (t & 3) * 8, lk = (t & 3) * 4;
  const int gi = I * DT + lr, gj = J * DT + lr;
  const bool vi = gi < n, vj = gj < n;
  const double* a1p = W1 + (size_t)(vi ? gi : 0) * m_pad;
  const double* b1p = W1 + (size_t)(vj ? gj : 0) * m_pad;
  double acc[4][4], zy[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c] = 0.0; zy[a][c] = 0.0; }
  double ra[8], rb[8];
  const int nk = m_pad / DKM;  // m_pad is a multiple of 32
  auto gload = [&](int kb) {
    const double2* pa = reinterpret_cast<const double2*>(a1p + kb * DKM + lk8);  // 16-byte aligned: m_pad % 32 == 0
    const double2* pb = reinterpret_cast<const double2*>(b1p + kb * DKM + lk8);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 va = vi ? pa[u] : make_double2(0.0, 0.0);
      const double2 vb = vj ? pb[u] : make_double2(0.0, 0.0);
      ra[2 * u] = va.x; ra[2 * u + 1] = va.y; rb[2 * u] = vb.x; rb[2 * u + 1] = vb.y;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int u = 0; u < 8; ++u) { Am[lk8 + u][lr] = ra[u]; Bm[lk8 + u][lr] = rb[u]; }
  };
  gload(0);
  sstore();
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    if (kb + 1 < nk) gload(kb + 1);  // global loads of the next slab overlap the FMAs of this one
#pragma unroll
    for (int kk = 0; kk < DKM; ++kk) {
      const double2 a01 = *reinterpret_cast<const double2*>(&Am[kk][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&Am[kk][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bm[kk][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bm[kk][tx * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y}, bb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], bb[v], acc[u][v]);
    }
    __syncthreads();
    if (kb + 1 < nk) sstore();
    __syncthreads();
  }
  // rank-21 Woodbury tail (32-wide): zy = Z_i . Y_j + Y_i . Z_j
  for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb * DK + lk + u;
      As[lk + u][lr] = vi ? Zb[(size_t)gi * 32 + k] : 0.0;
      Bs2[lk + u][lr] = vj ? Yb[(size_t)gj * 32 + k] : 0.0;
      As2[lk + u][lr] = vi ? Yb[(size_t)gi * 32 + k] : 0.0;
      Bs[lk + u][lr] = vj ? Zb[(size_t)gj * 32 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) {
      double a[4], bb[4], a2[4], b2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; a2[u] = As2[kk][ty * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { bb[u] = Bs[kk][tx * 4 + u]; b2[u] = Bs2[kk][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) zy[u][v] = fma(a[u], b2[v], fma(a2[u], bb[v], zy[u][v]));
    }
    __syncthreads();
  }
  // epilogue: the CTA owns P(I,J) and P(J,I); stage P(J,I)^T through shared memory (row-contiguous global accesses)
  double (*Pt)[DT + 1] = reinterpret_cast<double (*)[DT + 1]>(&sm[0][0][0]);
  static_assert(sizeof(sm) >= sizeof(double) * DT * (DT + 1), "staging tile does not fit");
  for (int e = t; e < DT * DT; e += 256) {
    const int r = e >> 6, c = e & 63;  // element (r, c) of block (J, I)
    const int gr = J * DT + r, gc = I * DT + c;
    Pt[c][r] = (gr < n && gc < n) ? P[(size_t)gr * n + gc] : 0.0;
  }
  __syncthreads();
  double outv[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = ty * 4 + a, gr = I * DT + r;
    const int oi = gr < n ? omega_inv[gr] : -1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cc = tx * 4 + c, gc = J * DT + cc;
      double v = 0.0;
      if (gr < n && gc < n) {
        const int oj = omega_inv[gc];
        double q = 0.0;
        if (oj >= 0) q += Qb[(size_t)gr * 32 + oj];
        if (oi >= 0) q += Qb[(size_t)gc * 32 + oi];
        v = 0.5 * (P[(size_t)gr * n + gc] + Pt[r][cc]) - acc[a][c] - 0.5 * q + 0.5 * zy[a][c];
        P[(size_t)gr * n + gc] = v;
      }
      outv[a][c] = v;
    }
  }
  if (I != J) {
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) Pt[ty * 4 + a][tx * 4 + c] = outv[a][c];
    __syncthreads();
    for (int e = t; e < DT * DT; e += 256) {
      const int r = e >> 6, c = e & 63;
      const int gr = J * DT + r, gc = I * DT + c;
      if (gr < n && gc < n) P[(size_t)gr * n + gc] = Pt[c][r];
    }
  }
}

// Small-N variant: 32x32 tile pairs (128 threads, 2x4 micro-tile) so that a 795-state covariance gives 325 CTAs instead
// of 91, and the K range [kbeg, kend) of the W1 columns is a parameter -- the part of the downdate that belongs to the
// SLAM columns of the compressed measurement exists long before the rest (xb_api.cu starts it on a side stream):
//   do_sym : Pout_ij = (Pin_ij + Pin_ji)/2 - ...   (otherwise Pin is already symmetric and only (I,J) is read)
//   do_tail: also the Woodbury / Omega terms  - (Q_i[k(j)] + Q_j[k(i)])/2 + (Z_i.Y_j + Y_i.Z_j)/2
#define D3 32
__global__ void __launch_bounds__(128) k_downdate32(const double* __restrict__ Pin, double* __restrict__ Pout, int n,
                                                    const double* __restrict__ W1, int ldw, int kbeg, int kend, int do_sym,
                                                    int do_tail, const int* __restrict__ omega_inv,
                                                    const double* __restrict__ Zb, const double* __restrict__ Yb,
                                                    const double* __restrict__ Qb) {
  XB_PDL_LONG();
  const int nt = (n + D3 - 1) / D3;
  int b = blockIdx.x, I = 0;
  while (b >= nt - I) { b -= nt - I; ++I; }
  const int J = I + b;
  const int t = threadIdx.x, ty = t >> 3, tx = t & 7;  // rows 2*ty.., cols 4*tx.. of tile (I, J)
  __shared__ __align__(16) double As[32][D3 + 2];  // [k][row of block I]
  __shared__ __align__(16) double Bs[32][D3 + 4];  // [k][row of block J]
  __shared__ double Pt[D3][D3 + 1];
  const int lr = t >> 2, lk8 = (t & 3) * 8;
  const int gi = I * D3 + lr, gj = J * D3 + lr;
  const bool vi = gi < n, vj = gj < n;
  const double* ap = W1 + (size_t)(vi ? gi : 0) * ldw;
  const double* bp = W1 + (size_t)(vj ? gj : 0) * ldw;
  double acc[2][4], zy[2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) { acc[u][v] = 0.0; zy[u][v] = 0.0; }
  double ra[8], rb[8];
  auto gload = [&](int k0) {  // kbeg, kend are multiples of 32 and ldw is a multiple of 32: 16-byte aligned double2 loads
    const double2* pa = reinterpret_cast<const double2*>(ap + k0 + lk8);
    const double2* pb = reinterpret_cast<const double2*>(bp + k0 + lk8);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 va = vi ? pa[u] : make_double2(0.0, 0.0);
      const double2 vb = vj ? pb[u] : make_double2(0.0, 0.0);
      ra[2 * u] = va.x; ra[2 * u + 1] = va.y; rb[2 * u] = vb.x; rb[2 * u + 1] = vb.y;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int u = 0; u < 8; ++u) { As[lk8 + u][lr] = ra[u]; Bs[lk8 + u][lr] = rb[u]; }
  };
  auto mma = [&](double (&c)[2][4]) {
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[kk][ty * 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4 + 2]);
      c[0][0] = fma(a01.x, b01.x, c[0][0]); c[0][1] = fma(a01.x, b01.y, c[0][1]);
      c[0][2] = fma(a01.x, b23.x, c[0][2]); c[0][3] = fma(a01.x, b23.y, c[0][3]);
      c[1][0] = fma(a01.y, b01.x, c[1][0]); c[1][1] = fma(a01.y, b01.y, c[1][1]);
      c[1][2] = fma(a01.y, b23.x, c[1][2]); c[1][3] = fma(a01.y, b23.y, c[1][3]);
    }
  };
  if (kbeg < kend) {
    gload(kbeg);
    sstore();
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    if (k0 + 32 < kend) gload(k0 + 32);
    mma(acc);
    __syncthreads();
    if (k0 + 32 < kend) sstore();
    __syncthreads();
  }
  if (do_tail) {
    // rank-21 Woodbury tail (32-wide): zy = Z_i . Y_j + Y_i . Z_j as two more slabs
    for (int pass = 0; pass < 2; ++pass) {
      const double* Ai = pass ? Yb : Zb;
      const double* Bj = pass ? Zb : Yb;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        As[lk8 + u][lr] = vi ? Ai[(size_t)gi * 32 + lk8 + u] : 0.0;
        Bs[lk8 + u][lr] = vj ? Bj[(size_t)gj * 32 + lk8 + u] : 0.0;
      }
      __syncthreads();
      mma(zy);
      __syncthreads();
    }
  }
  // epilogue: the CTA owns (I, J) and (J, I); the transposed tile goes through shared memory
  if (do_sym) {
    for (int e = t; e < D3 * D3; e += 128) {
      const int r = e >> 5, c = e & 31;  // element (r, c) of block (J, I)
      const int gr = J * D3 + r, gc = I * D3 + c;
      Pt[c][r] = (gr < n && gc < n) ? Pin[(size_t)gr * n + gc] : 0.0;
    }
    __syncthreads();
  }
  double outv[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int r = ty * 2 + a, gr = I * D3 + r;
    const int oi = (do_tail && gr < n) ? omega_inv[gr] : -1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cc = tx * 4 + c, gc = J * D3 + cc;
      double v = 0.0;
      if (gr < n && gc < n) {
        const double pij = Pin[(size_t)gr * n + gc];
        v = (do_sym ? 0.5 * (pij + Pt[r][cc]) : pij) - acc[a][c];
        if (do_tail) {
          const int oj = omega_inv[gc];
          double q = 0.0;
          if (oj >= 0) q += Qb[(size_t)gr * 32 + oj];
          if (oi >= 0) q += Qb[(size_t)gc * 32 + oi];
          v += -0.5 * q + 0.5 * zy[a][c];
        }
        Pout[(size_t)gr * n + gc] = v;
      }
      outv[a][c] = v;
    }
  }
  // mirror: block (J, I) = block (I, J)^T; on a diagonal block the lower half takes the upper half's values so that the
  // result is symmetric to the bit
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) Pt[ty * 2 + a][tx * 4 + c] = outv[a][c];
  __syncthreads();
  for (int e = t; e < D3 * D3; e += 128) {
    const int r = e >> 5, c = e & 31;
    const int gr = J * D3 + r, gc = I * D3 + c;
    if (gr < n && gc < n && (I != J || r > c)) Pout[(size_t)gr * n + gc] = Pt[c][r];
  }
}

// fp64 tensor-core (DMMA) form of k_downdate32: same tile pairs, arguments and result; four warps (2 x 2) own 16 x 16
// quadrants of the 32 x 32 tile = 2 x 2 m8n8k4 accumulators each, operands through the 3-stage cp.async ring.  The
// Woodbury tail  +(Z_i.Y_j + Y_i.Z_j)/2  enters the same accumulators as four more slabs with the A fragments scaled by -1/2.
__global__ void __launch_bounds__(128) k_downdate_mma(const double* __restrict__ Pin, double* __restrict__ Pout, int n,
                                                      const double* __restrict__ W1, int ldw, int kbeg, int kend, int do_sym,
                                                      int do_tail, const int* __restrict__ omega_inv,
                                                      const double* __restrict__ Zb, const double* __restrict__ Yb,
                                                      const double* __restrict__ Qb) {
  XB_PDL_LONG();
  const int nt = (n + D3 - 1) / D3;
  int b = blockIdx.x, I = 0;
  while (b >= nt - I) { b -= nt - I; ++I; }
  const int J = I + b;
  __shared__ double As[MG_ST][D3 * MG_LDK];
  __shared__ double Bs[MG_ST][D3 * MG_LDK];
  __shared__ double Pt[D3][D3 + 1];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wm0 = (warp >> 1) * 16, wn0 = (warp & 1) * 16;
  const int nmain = (kend - kbeg) / MG_BK;          // kbeg, kend are multiples of 32
  const int nslab = nmain + (do_tail ? 4 : 0);      // tail: Z_I.Y_J (2 slabs of 16), then Y_I.Z_J (2 slabs)
  auto issue = [&](int slab) {
    if (slab < nslab) {
      const double *a, *bsrc;
      int lda_, k0;
      if (slab < nmain) { a = W1; bsrc = W1; lda_ = ldw; k0 = kbeg + slab * MG_BK; }
      else {
        const int ts = slab - nmain;
        a = ts < 2 ? Zb : Yb; bsrc = ts < 2 ? Yb : Zb; lda_ = 32; k0 = (ts & 1) * MG_BK;
      }
      double* as = As[slab % MG_ST];
      double* bs = Bs[slab % MG_ST];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = t + 128 * u, r = e >> 4, k = e & 15;
        const bool vi = I * D3 + r < n, vj = J * D3 + r < n;
        cp_async8(as + r * MG_LDK + k, vi ? a + (size_t)(I * D3 + r) * lda_ + k0 + k : a, vi);
        cp_async8(bs + r * MG_LDK + k, vj ? bsrc + (size_t)(J * D3 + r) * lda_ + k0 + k : bsrc, vj);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll
  for (int s_ = 0; s_ < MG_ST - 1; ++s_) issue(s_);
  for (int slab = 0; slab < nslab; ++slab) {
    asm volatile("cp.async.wait_group %0;" ::"n"(MG_ST - 2) : "memory");
    __syncthreads();
    issue(slab + MG_ST - 1);
    const double* as = As[slab % MG_ST];
    const double* bs = Bs[slab % MG_ST];
    const double sc = slab < nmain ? 1.0 : -0.5;
#pragma unroll
    for (int kk = 0; kk < MG_BK; kk += 4) {
      double a[2], bb[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = sc * as[(wm0 + 8 * i + g) * MG_LDK + kk + tg];
#pragma unroll
      for (int j = 0; j < 2; ++j) bb[j] = bs[(wn0 + 8 * j + g) * MG_LDK + kk + tg];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // epilogue: the CTA owns (I, J) and (J, I); the transposed tile goes through shared memory
  if (do_sym) {
    for (int e = t; e < D3 * D3; e += 128) {
      const int r = e >> 5, c = e & 31;  // element (r, c) of block (J, I)
      const int gr = J * D3 + r, gc = I * D3 + c;
      Pt[c][r] = (gr < n && gc < n) ? Pin[(size_t)gr * n + gc] : 0.0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = wm0 + 8 * i + g, gr = I * D3 + r;
    const int oi = (do_tail && gr < n) ? omega_inv[gr] : -1;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = wn0 + 8 * j + 2 * tg + h, gc = J * D3 + cc;
        double v = 0.0;
        if (gr < n && gc < n) {
          const double pij = Pin[(size_t)gr * n + gc];
          v = (do_sym ? 0.5 * (pij + Pt[r][cc]) : pij) - acc[i][j][h];
          if (do_tail) {
            const int oj = omega_inv[gc];
            double q = 0.0;
            if (oj >= 0) q += Qb[(size_t)gr * 32 + oj];
            if (oi >= 0) q += Qb[(size_t)gc * 32 + oi];
            v -= 0.5 * q;
          }
          Pout[(size_t)gr * n + gc] = v;
        }
        Pt[r][cc] = v;  // this lane was the only reader of Pt[r][cc]
      }
  }
  __syncthreads();
  for (int e = t; e < D3 * D3; e += 128) {
    const int r = e >> 5, c = e & 31;
    const int gr = J * D3 + r, gc = I * D3 + c;
    if (gr < n && gc < n && (I != J || r > c)) Pout[(size_t)gr * n + gc] = Pt[c][r];
  }
}

// Pout = [sym](Pin) - W1[:, kbeg:kend] W1[:, kbeg:kend]^T [+ Woodbury/Omega tail]; Pin == Pout allowed.
void downdate_f64_range(cudaStream_t s, const double* Pin, double* Pout, int n, const double* T, int m_pad, int n_pad, int kbeg,
                        int kend, int do_sym, int do_tail, const int* omega_inv, const double* Zb, const double* Yb,
                        const double* Qb) {
  const int nt = (n + D3 - 1) / D3;
  // large covariances: TMA-staged 128 x 64 tiles (k_gemm_tma.cu).  Without do_sym the input is symmetric to the bit (it is
  // the output of the do_sym pass), so symmetrising again changes nothing.
  if (g_use_mma && downdate_sym_tma(s, n, kend - kbeg, T + (size_t)m_pad * m_pad + kbeg, m_pad, Pin, Pout, n,
                                    do_tail ? omega_inv : nullptr, do_tail ? Zb : nullptr, do_tail ? Yb : nullptr,
                                    do_tail ? Qb : nullptr))
    return;
  if (g_use_mma)
    XB_LAUNCH(k_downdate_mma, nt * (nt + 1) / 2, 128, 0, s, Pin, Pout, n, T + (size_t)m_pad * m_pad, m_pad, kbeg, kend, do_sym, do_tail,
                                                     omega_inv, Zb, Yb, Qb);
  else
    XB_LAUNCH(k_downdate32, nt * (nt + 1) / 2, 128, 0, s, Pin, Pout, n, T + (size_t)m_pad * m_pad, m_pad, kbeg, kend, do_sym, do_tail,
                                                   omega_inv, Zb, Yb, Qb);
  count_launch();
}

void downdate_f64(cudaStream_t s, double* P, int n, const double* T, int m_pad, int n_pad, const int* omega_inv,
                  const double* Zb, const double* Yb, const double* Qb) {
  const int nt = (n + DT - 1) / DT;
  if (g_use_mma || nt * (nt + 1) / 2 < 296) {  // tensor-core tile pairs; or (DFMA) less than one wave of 64x64 pairs
    downdate_f64_range(s, P, P, n, T, m_pad, n_pad, 0, m_pad, 1, 1, omega_inv, Zb, Yb, Qb);
    return;
  }
  XB_LAUNCH(k_downdate, nt * (nt + 1) / 2, 256, 0, s, P, n, T, m_pad, n_pad, omega_inv, Zb, Yb, Qb);
  count_launch();
}

// P <- (P + P^T)/2 only (cov_update path with an empty measurement never reaches here; used by applyCI glue).
__global__ void k_symmetrise(double* __restrict__ P, int n) {
  XB_PDL_SHORT();
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i < n && j < n && i < j) {
    const double v = 0.5 * (P[(size_t)i * n + j] + P[(size_t)j * n + i]);
    P[(size_t)i * n + j] = v;
    P[(size_t)j * n + i] = v;
  }
}
void symmetrise(cudaStream_t s, double* P, int n) {
  dim3 g((n + 15) / 16, (n + 15) / 16), b(16, 16);
  XB_LAUNCH(k_symmetrise, g, b, 0, s, P, n);
  count_launch();
}

// y[r] = sum_k A[r, k] * x[k]   (one warp per row)
__global__ void k_gemv(int rows, int cols, const double* __restrict__ A, int lda, const double* __restrict__ x,
                       double* __restrict__ y) {
  XB_PDL_SHORT();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= rows) return;
  double s = 0.0;
  for (int k = lane; k < cols; k += 32) s = fma(A[(size_t)w * lda + k], x[k], s);
  s = xb_warp_sum(s);
  if (lane == 0) y[w] = s;
}
void gemv(cudaStream_t s, int rows, int cols, const double* A, int lda, const double* x, double* y) {
  if (rows <= 0) return;
  XB_LAUNCH(k_gemv, (rows * 32 + 127) / 128, 128, 0, s, rows, cols, A, lda, x, y);
  count_launch();
}

// out (cols x rows) = in (rows x cols)^T
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
  XB_PDL_SHORT();
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
  XB_LAUNCH(k_transpose, g, b, 0, s, in, out, rows, cols);
  count_launch();
}

}  // namespace xb
