// Tensor-core covariance downdate (optional, cfg.downdate_precision = 1):  P <- sym(P) - sym(K A2^T)
// on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).
// reference arithmetic: src/x/ekf/updater.cpp:131-136.
//
// Precision scheme ("3xTF32"): every fp64 operand x is split as x ~ hi + lo with hi = tf32(x), lo = tf32(x - hi)
// (2 x 11 significant bits); the product A B^T is accumulated as  hi hi^T + hi lo^T + lo hi^T  in fp32, i.e. the
// contraction is fp32-accurate (rel. error ~1e-7 of sum |a_ik b_jk|), NOT fp64-accurate.  The covariance itself stays
// fp64: P_new = sym(P) - fp64(acc).  This is therefore opt-in; the default path (k_downdate, fp64 CUDA cores)
// is the one the parity tests hold to 1e-8.
//
// Operands.  With W1 the TRSM'd rows of the tall buffer and Z, Y the rank-21 Woodbury factors (k_update.cu)
//     K A2^T (symmetrised) = W1 W1^T - (Z Y^T + Y Z^T)/2 + Omega lookups
//                          = [W1 | Z | Y] [W1 | -Y/2 | -Z/2]^T + Omega lookups
// so one GEMM  C = A B^T  with K' = m_pad + 64 covers everything but the Omega lookups (epilogue).
// k_tc_stage writes the hi/lo fp32 operand matrices once; k_downdate_tc is one CTA per upper-triangular 128x128
// tile: operands are staged into shared memory in the canonical K-major SWIZZLE_128B layout (8 rows x 128 B atoms,
// 16-byte chunk index XOR row%8), one elected thread issues the tcgen05.mma's, tcgen05.commit + an mbarrier hand
// the shared-memory slot back, and the epilogue reads the accumulators with tcgen05.ld (lane = row).
#include "xb_kernels.h"

namespace xb {

#define TCM 128   // tile rows (UMMA M)
#define TCN 128   // tile cols (UMMA N)
#define TCK 32    // tf32 elements per k-block = one 128-byte swizzle row
#define UMMA_K 8  // tf32 elements per tcgen05.mma (32 bytes)

__device__ __forceinline__ float to_tf32(double x) {
  float f = (float)x;
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return __uint_as_float(u);
}

// fp64 -> (hi, lo) TF32 pairs for A = [W1 | Z | Y] and B = [W1 | -Y/2 | -Z/2]; rows >= n are zero.
__global__ void k_tc_stage(int n, int n128, int m_pad, int Kp, const double* __restrict__ W1, const double* __restrict__ Zb,
                           const double* __restrict__ Yb, float* __restrict__ Ahi, float* __restrict__ Alo,
                           float* __restrict__ Bhi, float* __restrict__ Blo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (k >= Kp || i >= n128) return;
  double a = 0.0, b = 0.0;
  if (i < n) {
    if (k < m_pad) { a = W1[(size_t)i * m_pad + k]; b = a; }
    else if (k < m_pad + 32) { const int c = k - m_pad; a = Zb[(size_t)i * 32 + c]; b = -0.5 * Yb[(size_t)i * 32 + c]; }
    else { const int c = k - m_pad - 32; a = Yb[(size_t)i * 32 + c]; b = -0.5 * Zb[(size_t)i * 32 + c]; }
  }
  const float ah = to_tf32(a), bh = to_tf32(b);
  const size_t o = (size_t)i * Kp + k;
  Ahi[o] = ah; Alo[o] = to_tf32(a - (double)ah);
  Bhi[o] = bh; Blo[o] = to_tf32(b - (double)bh);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major) bits [16,30), stride byte offset (8 rows x 128 B = 1024 B) >> 4 in
// bits [32,46), version 1 at bit 46, layout type SWIZZLE_128B (= 2) in bits [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(128) k_downdate_tc(double* __restrict__ P, int n, int Kp, const float* __restrict__ Ahi,
                                                     const float* __restrict__ Alo, const float* __restrict__ Bhi,
                                                     const float* __restrict__ Blo, const int* __restrict__ omega_inv,
                                                     const double* __restrict__ Qb) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand slots: A_hi, A_lo, B_hi, B_lo (128 rows x 128 B each)
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int nt = (n + TCM - 1) / TCM;
  int b = blockIdx.x, I = 0;
  while (b >= nt - I) { b -= nt - I; ++I; }
  const int J = I + b;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t bar_a = smem_u32(&bar);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_c = tmem_slot;
  const uint32_t idesc = make_idesc();
  const float* src[4] = {Ahi + (size_t)I * TCM * Kp, Alo + (size_t)I * TCM * Kp, Bhi + (size_t)J * TCN * Kp,
                         Blo + (size_t)J * TCN * Kp};
  const int nkb = Kp / TCK;
  for (int kb = 0; kb < nkb; ++kb) {
    // stage 4 operand slabs (128 rows x 32 tf32) into the swizzled K-major layout: 16-byte chunks
#pragma unroll
    for (int op = 0; op < 4; ++op) {
      unsigned char* dst = sm + op * (TCM * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int idx = t + 128 * c;
        const int row = idx >> 3, chunk = idx & 7;
        const float4 v = *reinterpret_cast<const float4*>(src[op] + (size_t)row * Kp + kb * TCK + chunk * 4);
        *reinterpret_cast<float4*>(dst + (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sm), a_lo = a_hi + TCM * 128, b_hi = a_lo + TCM * 128, b_lo = b_hi + TCN * 128;
#pragma unroll
      for (int ks = 0; ks < TCK / UMMA_K; ++ks) {
        const uint32_t off = ks * UMMA_K * 4;  // 32 bytes along K inside the 128-byte swizzle row
        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
        umma_tf32(tmem_c, make_desc(a_hi + off), make_desc(b_hi + off), idesc, first);
        umma_tf32(tmem_c, make_desc(a_hi + off), make_desc(b_lo + off), idesc, 1u);
        umma_tf32(tmem_c, make_desc(a_lo + off), make_desc(b_hi + off), idesc, 1u);
      }
      // arrives on the mbarrier once every MMA issued so far has finished reading shared memory / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    mbar_wait(bar_a, kb & 1);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= tile rows), 32 columns at a time
  const int r = warp * 32 + lane;
  const int gi = I * TCM + r;
  const int oi = gi < n ? omega_inv[gi] : -1;
  for (int cb = 0; cb < TCN / 32; ++cb) {
    uint32_t v[32];
    const uint32_t taddr = tmem_c + ((uint32_t)(warp * 32) << 16) + cb * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (gi < n) {
#pragma unroll 4
      for (int c = 0; c < 32; ++c) {
        const int gj = J * TCN + cb * 32 + c;
        if (gj >= n || (I == J && gj < gi)) continue;
        const int oj = omega_inv[gj];
        double q = 0.0;
        if (oj >= 0) q += Qb[(size_t)gi * 32 + oj];
        if (oi >= 0) q += Qb[(size_t)gj * 32 + oi];
        const double pij = P[(size_t)gi * n + gj], pji = P[(size_t)gj * n + gi];
        const double val = 0.5 * (pij + pji) - (double)__uint_as_float(v[c]) - 0.5 * q;
        P[(size_t)gi * n + gj] = val;
        P[(size_t)gj * n + gi] = val;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"(128));
}

size_t downdate_tc_workspace_bytes(int n, int m_pad) {
  const int n128 = (n + TCM - 1) / TCM * TCM, Kp = m_pad + 64;
  return (size_t)4 * n128 * Kp * sizeof(float) + 1024;
}

void downdate_tc(cudaStream_t s, double* P, int n, const double* T, int m_pad, int n_pad, const int* omega_inv,
                 const double* Zb, const double* Yb, const double* Qb, void* ws) {
  (void)n_pad;
  const int n128 = (n + TCM - 1) / TCM * TCM, Kp = m_pad + 64;
  float* Ahi = reinterpret_cast<float*>(ws);
  float* Alo = Ahi + (size_t)n128 * Kp;
  float* Bhi = Alo + (size_t)n128 * Kp;
  float* Blo = Bhi + (size_t)n128 * Kp;
  dim3 gs((Kp + 127) / 128, n128);
  k_tc_stage<<<gs, 128, 0, s>>>(n, n128, m_pad, Kp, T + (size_t)m_pad * m_pad, Zb, Yb, Ahi, Alo, Bhi, Blo);
  count_launch();
  static bool attr = false;
  const int smem = 4 * TCM * 128 + 1024;
  if (!attr) { cudaFuncSetAttribute(k_downdate_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  const int nt = n128 / TCM;
  k_downdate_tc<<<nt * (nt + 1) / 2, 128, smem, s>>>(P, n, Kp, Ahi, Alo, Bhi, Blo, omega_inv, Qb);
  count_launch();
}

}  // namespace xb
