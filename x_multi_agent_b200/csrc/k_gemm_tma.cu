// TMA-staged fp64 tensor-core contraction for the large shapes of the covariance update (cfg-5: N = 2715, 1900 rows).
//
//   C (M x N) = beta * C + alpha * A (M x K) * B (N x K)^T           all row-major, K contiguous ("NT")
//   optionally in its symmetric-downdate form (SYM): only tiles on or below the diagonal are computed,
//   C_out = (C_in + C_in^T)/2 + alpha * A B^T is written to the tile and to its mirror image (updater.cpp:131-136).
//
// reference: the dense products of Updater::applyUpdate (src/x/ekf/updater.cpp:117-141: S = H P H^T, K = P H^T S^-1,
// P <- (I - K H) P); here they appear as the Schur complement of the factored SLAM columns on the rest of the tall buffer
// and as the covariance downdate P <- sym(P) - W1 W1^T.
//
// Design (B200, sm_100a): fp64 has no tcgen05 kind, the fp64 tensor-core instruction is mma.sync.m8n8k4 (DMMA, 36.9 TFLOP/s
// measured = the DFMA peak, tools/bench_dmma.cu).  What bounds the small-tile kernel of k_linalg.cu at these sizes is operand
// staging (8-byte cp.async, 3 stages, 32 x 64 tiles: 6 fragment loads per 8 DMMA).  Here
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, one elected producer lane, 4-stage mbarrier ring):
//     a K-slab of 16 doubles is exactly one 128-byte swizzle row, so the DMMA fragment loads (8 rows x 4 k per instruction)
//     are bank-conflict free without padding and out-of-range rows / k are zero-filled by the copy engine;
//   * CTA tile 128 x 64, eight consumer warps with 32 x 32 warp tiles (16 DMMA accumulators): 8 fragment loads per 16 DMMA;
//   * lane 0 of warp 0 is the producer in between its own tiles (256 threads x 126 registers fit twice per SM, a ninth
//     producer warp would not); consumers release a stage with one mbarrier arrive per warp.
// TMA needs 16-byte aligned rows: the tall buffer (leading dimension m_pad, a multiple of 32) qualifies, the covariance P
// (leading dimension N, odd at every BASELINE size) is touched by the epilogue only.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdint>
#include <cstdlib>

#include "xb_kernels.h"

namespace xb {

#define TG_BM 128
#define TG_BN 64
#define TG_BK 16
#define TG_ST 4
#define TG_A_BYTES (TG_BM * TG_BK * 8)
#define TG_B_BYTES (TG_BN * TG_BK * 8)
#define TG_STAGE_BYTES (TG_A_BYTES + TG_B_BYTES)
#define TG_SMEM (TG_ST * TG_STAGE_BYTES + 1024 + 128)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c_inner, int c_outer, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"((uint64_t)tm), "r"(c_inner), "r"(c_outer), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma884_t(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// SYM: blockIdx.x enumerates the tile pairs (bi >= bj) of the lower triangle in 128 x 128 blocks, split into two 64-column
// halves; Cin may alias C (every tile pair is read and written by one CTA only).
// Tail (SYM only, ntail = 4): the Woodbury / Omega terms of the exact unsymmetrised update (k_update.cu), four more K-slabs
// Z.Y^T (2 slabs of 16) and Y.Z^T (2 slabs) with the A fragments scaled by -1/2, and -(Q[i][omega_inv j] + Q[j][omega_inv i])/2
// in the epilogue -- the same arithmetic as k_downdate_mma (k_linalg.cu).
struct TmaTail {
  CUtensorMap zA, yA, zB, yB;   // Zb / Yb (n x 32) with the A-side (128 rows) and B-side (64 rows) boxes
};
template <bool SYM>
__global__ void __launch_bounds__(256, 2) k_gemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  int M, int N, int K, double alpha, double beta, const double* Cin, double* __restrict__ C,
                                                  int ldc, const __grid_constant__ TmaTail tail, int ntail,
                                                  const int* __restrict__ omega_inv, const double* __restrict__ Qb) {
  XB_PDL_LONG();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(base);
  const uint32_t bars = sbase + TG_ST * TG_STAGE_BYTES;  // full[TG_ST], empty[TG_ST]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int m0, n0;
  if (SYM) {
    // tile pair index -> (bi, bj, half): bi*(bi+1)/2 + bj pairs of 128-blocks, 2 halves each
    const int pair = blockIdx.x >> 1, half = blockIdx.x & 1;
    int bi = (int)((sqrtf(8.0f * (float)pair + 1.0f) - 1.0f) * 0.5f);
    while (bi * (bi + 1) / 2 > pair) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= pair) ++bi;
    const int bj = pair - bi * (bi + 1) / 2;
    m0 = bi * TG_BM;
    n0 = bj * TG_BM + half * TG_BN;
  } else {
    m0 = blockIdx.y * TG_BM;
    n0 = blockIdx.x * TG_BN;
  }
  if (t == 0) {
    for (int s = 0; s < TG_ST; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (TG_ST + s), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nmain = (K + TG_BK - 1) / TG_BK;
  const int nslab = nmain + (SYM ? ntail : 0);
  // producer = lane 0 of warp 0, in between its own tiles: slab s + TG_ST - 1 is requested at the top of iteration s, once
  // every warp has released the stage it goes into (a dedicated ninth warp costs the second CTA per SM: 288 threads x 126
  // registers do not fit twice)
  auto produce = [&](int slab) {
    if (slab >= nslab) return;
    const int s = slab % TG_ST;
    if (slab >= TG_ST) mbar_wait(bars + 8 * (TG_ST + s), ((slab / TG_ST) - 1) & 1);
    const uint32_t full = bars + 8 * s;
    mbar_expect_tx(full, TG_STAGE_BYTES);
    if (slab < nmain) {
      tma_load_2d(sbase + s * TG_STAGE_BYTES, &tmA, slab * TG_BK, m0, full);
      tma_load_2d(sbase + s * TG_STAGE_BYTES + TG_A_BYTES, &tmB, slab * TG_BK, n0, full);
    } else {
      const int ts = slab - nmain, k0 = (ts & 1) * TG_BK;
      tma_load_2d(sbase + s * TG_STAGE_BYTES, ts < 2 ? &tail.zA : &tail.yA, k0, m0, full);
      tma_load_2d(sbase + s * TG_STAGE_BYTES + TG_A_BYTES, ts < 2 ? &tail.yB : &tail.zB, k0, n0, full);
    }
  };
  if (t == 0)
    for (int s = 0; s < TG_ST - 1; ++s) produce(s);
  const int g = lane >> 2, tg = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int slab = 0; slab < nslab; ++slab) {
    const int s = slab % TG_ST;
    if (t == 0) produce(slab + TG_ST - 1);
    __syncwarp();
    mbar_wait(bars + 8 * s, (slab / TG_ST) & 1);
    const unsigned char* as = base + s * TG_STAGE_BYTES;
    const unsigned char* bs = as + TG_A_BYTES;
#pragma unroll
    for (int kk = 0; kk < TG_BK; kk += 4) {
      const int k = kk + tg;
      // SWIZZLE_128B: 16-byte chunk index (k / 2) XOR (row % 8); row % 8 == g for every fragment row
      const int off = (((k >> 1) ^ g) << 4) + ((k & 1) << 3);
      double a[4], b[4];
      const double sc = (SYM && slab >= nmain) ? -0.5 : 1.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sc * *(const double*)(as + (wm + 8 * i + g) * 128 + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *(const double*)(bs + (wn + 8 * j + g) * 128 + off);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884_t(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (TG_ST + s));
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = m0 + wm + 8 * i + g;
    if (gr >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = n0 + wn + 8 * j + 2 * tg + h;
        if (gc >= N) continue;
        if (SYM) {
          if (gc > gr) continue;
          double v = 0.5 * (Cin[(size_t)gr * ldc + gc] + Cin[(size_t)gc * ldc + gr]) + alpha * acc[i][j][h];
          if (ntail) {
            const int oi = omega_inv[gr], oj = omega_inv[gc];
            double q = 0.0;
            if (oj >= 0) q += Qb[(size_t)gr * 32 + oj];
            if (oi >= 0) q += Qb[(size_t)gc * 32 + oi];
            v -= 0.5 * q;
          }
          C[(size_t)gr * ldc + gc] = v;
          C[(size_t)gc * ldc + gr] = v;
        } else {
          double* p = &C[(size_t)gr * ldc + gc];
          *p = (beta == 0.0) ? alpha * acc[i][j][h] : alpha * acc[i][j][h] + beta * (*p);
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (PFN_cuTensorMapEncodeTiled_v12000)p;
  }();
  return fn;
}
static bool make_map(CUtensorMap* tm, const double* base, int rows, int K, int ld, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {TG_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tma_ok(const double* A, int lda, const double* B, int ldb) {
  static const bool off = [] { const char* e = getenv("XB_NO_TMA"); return e && e[0] == '1'; }();  // A/B switch
  return !off && encode_fn() && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && lda % 2 == 0 && ldb % 2 == 0;
}
static bool g_attr_set = false;
template <bool SYM> static void set_attr() {
  cudaFuncSetAttribute(k_gemm_tma<SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM);
}

// Returns false when the shapes / alignments do not qualify (the caller then uses the cp.async kernel of k_linalg.cu).
bool gemm_nt_tma(cudaStream_t s, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                 double* C, int ldc) {
  const long tiles = (long)((M + TG_BM - 1) / TG_BM) * ((N + TG_BN - 1) / TG_BN);
  if (tiles < 96 || K < 4 * TG_BK || !tma_ok(A, lda, B, ldb)) return false;
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, M, K, lda, TG_BM) || !make_map(&tb, B, N, K, ldb, TG_BN)) return false;
  if (!g_attr_set) { set_attr<false>(); set_attr<true>(); g_attr_set = true; }
  dim3 grid((N + TG_BN - 1) / TG_BN, (M + TG_BM - 1) / TG_BM);
  static const TmaTail no_tail{};
  XB_LAUNCH((k_gemm_tma<false>), grid, 256, TG_SMEM, s, ta, tb, M, N, K, alpha, beta, nullptr, C, ldc, no_tail, 0, nullptr, nullptr);
  count_launch();
  return true;
}

// Pout = (Pin + Pin^T)/2 - W W^T [+ (Z Y^T + Y Z^T)/2 - (Q terms)/2]  for W (n x K, row-major, leading dimension ldw);
// Pin may alias Pout.  Zb / Yb / Qb (n x 32) and omega_inv as in downdate_f64_range; nullptr: no tail.
bool downdate_sym_tma(cudaStream_t s, int n, int K, const double* W, int ldw, const double* Pin, double* Pout, int ldp,
                      const int* omega_inv, const double* Zb, const double* Yb, const double* Qb) {
  const int nb = (n + TG_BM - 1) / TG_BM;
  if (nb * (nb + 1) < 96 || K < 4 * TG_BK || !tma_ok(W, ldw, W, ldw)) return false;
  CUtensorMap ta, tb;
  if (!make_map(&ta, W, n, K, ldw, TG_BM) || !make_map(&tb, W, n, K, ldw, TG_BN)) return false;
  TmaTail tail{};
  int ntail = 0;
  if (Zb && Yb && Qb && omega_inv) {
    if (!tma_ok(Zb, 32, Yb, 32) || !make_map(&tail.zA, Zb, n, 32, 32, TG_BM) || !make_map(&tail.yA, Yb, n, 32, 32, TG_BM) ||
        !make_map(&tail.zB, Zb, n, 32, 32, TG_BN) || !make_map(&tail.yB, Yb, n, 32, 32, TG_BN))
      return false;
    ntail = 4;
  }
  if (!g_attr_set) { set_attr<false>(); set_attr<true>(); g_attr_set = true; }
  XB_LAUNCH((k_gemm_tma<true>), nb * (nb + 1), 256, TG_SMEM, s, ta, tb, n, n, K, -1.0, 0.0, Pin, Pout, ldp, tail, ntail, omega_inv, Qb);
  count_launch();
  return true;
}

}  // namespace xb
