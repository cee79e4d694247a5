// Micro-benchmark: latency of a 32x32 fp64 Cholesky / triangular solve done by ONE warp (critical path of the
// dataflow tile Cholesky).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_potrf bench_potrf.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
#define TC 32
__device__ __forceinline__ long long clk() { return clock64(); }

// A: rows in registers, fully unrolled, sqrt + div
__device__ void potrf_A(double (*Cs)[TC + 1], double* col0, double* col1, int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double piv = __shfl_sync(0xffffffffu, a[c], c);
    const double d = sqrt(piv);
    double l = 0.0;
    if (lane == c) l = d; else if (lane > c) l = a[c] / d;
    a[c] = l;
    double* col = (c & 1) ? col1 : col0;
    col[lane] = l;
    __syncwarp();
#pragma unroll
    for (int k = c + 1; k < TC; ++k) { const double lk = col[k]; if (k <= lane) a[k] = fma(-l, lk, a[k]); }
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
// C: registers, unrolled, rsqrt + mul
__device__ void potrf_C(double (*Cs)[TC + 1], double* col0, double* col1, int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double piv = __shfl_sync(0xffffffffu, a[c], c);
    const double rs = rsqrt(piv);
    double l = 0.0;
    if (lane == c) l = piv * rs; else if (lane > c) l = a[c] * rs;
    a[c] = l;
    double* col = (c & 1) ? col1 : col0;
    col[lane] = l;
    __syncwarp();
#pragma unroll
    for (int k = c + 1; k < TC; ++k) { const double lk = col[k]; if (k <= lane) a[k] = fma(-l, lk, a[k]); }
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
// B: rows in shared memory, compact loops, rsqrt, batches of 4
__device__ void potrf_B(double (*Cs)[TC + 1], int lane) {
#pragma unroll 1
  for (int c = 0; c < TC; ++c) {
    const double piv = Cs[c][c];
    const double rs = rsqrt(piv);
    double l = 0.0;
    if (lane == c) l = piv * rs; else if (lane > c) l = Cs[lane][c] * rs;
    Cs[lane][c] = l;
    __syncwarp();
    int k = c + 1;
#pragma unroll 1
    for (; k + 3 < TC; k += 4) {
      const double l0 = Cs[k][c], l1 = Cs[k + 1][c], l2 = Cs[k + 2][c], l3 = Cs[k + 3][c];
      double r0 = Cs[lane][k], r1 = Cs[lane][k + 1], r2 = Cs[lane][k + 2], r3 = Cs[lane][k + 3];
      r0 = fma(-l, l0, r0); r1 = fma(-l, l1, r1); r2 = fma(-l, l2, r2); r3 = fma(-l, l3, r3);
      if (k <= lane) Cs[lane][k] = r0;
      if (k + 1 <= lane) Cs[lane][k + 1] = r1;
      if (k + 2 <= lane) Cs[lane][k + 2] = r2;
      if (k + 3 <= lane) Cs[lane][k + 3] = r3;
    }
    for (; k < TC; ++k) { const double lk = Cs[k][c]; if (k <= lane) Cs[lane][k] = fma(-l, lk, Cs[lane][k]); }
    __syncwarp();
  }
}
// D: column-oriented: lane owns COLUMN `lane` (= row by symmetry) in registers; left-looking per column:
//    no shuffles of the pivot row: at step c every lane k>c needs l_kc; lane c broadcasts its column via smem once.
__device__ void potrf_D(double (*Cs)[TC + 1], double* col0, int lane) {
  // lane holds column `lane`: a[r] = A[r][lane], r >= lane
  double a[TC];
#pragma unroll
  for (int r = 0; r < TC; ++r) a[r] = Cs[r][lane];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    // lane c finalises its column: a[r] (r>=c) already has all updates from columns < c
    if (lane == c) {
      const double rs = rsqrt(a[c]);
#pragma unroll
      for (int r = c; r < TC; ++r) { a[r] *= rs; col0[r] = a[r]; }   // a[c] = piv*rs = sqrt
    }
    __syncwarp();
    // lanes k > c update their column: a[r] -= l_rc * l_kc  for r >= k
    const double lkc = col0[lane];
#pragma unroll
    for (int r = c + 1; r < TC; ++r) { const double lr = col0[r]; if (lane > c && r >= lane) a[r] = fma(-lr, lkc, a[r]); }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < TC; ++r) if (r >= lane) Cs[r][lane] = a[r]; else Cs[r][lane] = 0.0;
}
// E: left-looking, rows in registers, final L rows mirrored in smem; per-step chain = dot + shfl + rsqrt + mul
__device__ void potrf_E(double (*Cs)[TC + 1], double (*Ls)[TC + 1], int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < c; ++k) {
      const double lc = Ls[c][k];
      if ((k & 3) == 0) s0 = fma(a[k], lc, s0);
      else if ((k & 3) == 1) s1 = fma(a[k], lc, s1);
      else if ((k & 3) == 2) s2 = fma(a[k], lc, s2);
      else s3 = fma(a[k], lc, s3);
    }
    const double v = a[c] - ((s0 + s1) + (s2 + s3));
    const double piv = __shfl_sync(0xffffffffu, v, c);
    const double rs = rsqrt(piv);
    double l = 0.0;
    if (lane == c) l = piv * rs; else if (lane > c) l = v * rs;
    a[c] = l;
    Ls[lane][c] = l;
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
// F: right-looking, registers, column broadcast by shuffles (no smem, no syncwarp)
__device__ void potrf_F(double (*Cs)[TC + 1], int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double piv = __shfl_sync(0xffffffffu, a[c], c);
    const double rs = rsqrt(piv);
    double l = 0.0;
    if (lane == c) l = piv * rs; else if (lane > c) l = a[c] * rs;
    a[c] = l;
#pragma unroll
    for (int k = c + 1; k < TC; ++k) { const double lk = __shfl_sync(0xffffffffu, l, k); if (k <= lane) a[k] = fma(-l, lk, a[k]); }
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
// TRSM variants: X L^T = C, lane = row
__device__ void trsm_A(double (*Cs)[TC + 1], double (*Ls)[TC + 1], int lane) {
  double a[TC];
#pragma unroll
  for (int k = 0; k < TC; ++k) a[k] = Cs[lane][k];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const double x = a[c] / Ls[c][c];
    a[c] = x;
#pragma unroll
    for (int k = c + 1; k < TC; ++k) a[k] = fma(-x, Ls[k][c], a[k]);
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
__device__ void trsm_B(double (*Cs)[TC + 1], double (*Ls)[TC + 1], const double* rd, int lane) {  // reciprocal diag
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

__global__ void k(const double* A, double* out, long long* times, int variant, int reps) {
  __shared__ double Cs[TC][TC + 1], Ls[TC][TC + 1];
  __shared__ double col0[TC], col1[TC], rd[TC];
  const int lane = threadIdx.x;
  for (int rep = 0; rep < reps; ++rep) {
    for (int e = lane; e < TC * TC; e += 32) { Cs[e >> 5][e & 31] = A[e]; Ls[e >> 5][e & 31] = A[e]; }
    __syncwarp();
    if (variant >= 10) {  // prepare L for trsm
      potrf_C(Ls, col0, col1, lane);
      __syncwarp();
      rd[lane] = 1.0 / Ls[lane][lane];
      __syncwarp();
    }
    const long long t0 = clk();
    if (variant == 0) potrf_A(Cs, col0, col1, lane);
    else if (variant == 1) potrf_B(Cs, lane);
    else if (variant == 2) potrf_C(Cs, col0, col1, lane);
    else if (variant == 3) potrf_D(Cs, col0, lane);
    else if (variant == 4) potrf_E(Cs, Ls, lane);
    else if (variant == 5) potrf_F(Cs, lane);
    else if (variant == 10) trsm_A(Cs, Ls, lane);
    else if (variant == 11) trsm_B(Cs, Ls, rd, lane);
    __syncwarp();
    const long long t1 = clk();
    if (lane == 0) times[rep] = t1 - t0;
  }
  for (int e = lane; e < TC * TC; e += 32) out[e] = Cs[e >> 5][e & 31];
}

int main() {
  double hA[TC * TC];
  // SPD test matrix
  for (int i = 0; i < TC; ++i) for (int j = 0; j < TC; ++j) hA[i * TC + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1.0 + abs(i - j));
  double *dA, *dout; long long* dt;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dout, sizeof(hA)); cudaMalloc(&dt, 64 * 8);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  const int variants[] = {0, 1, 2, 3, 4, 5, 10, 11};
  const char* names[] = {"potrf_A regs/unrolled sqrt+div", "potrf_B smem loops rsqrt", "potrf_C regs/unrolled rsqrt",
                         "potrf_D column-owner regs rsqrt", "potrf_E left-looking regs rsqrt", "potrf_F right-looking shuffles", "trsm_A regs div", "trsm_B regs recip"};
  double ref[TC * TC];
  for (int v = 0; v < 8; ++v) {
    k<<<1, 32>>>(dA, dout, dt, variants[v], 6);
    cudaDeviceSynchronize();
    long long ht[6]; double ho[TC * TC];
    cudaMemcpy(ht, dt, 6 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
    if (v == 0) for (int e = 0; e < TC * TC; ++e) ref[e] = ho[e];
    double err = 0; if (v < 6) for (int i = 0; i < TC; ++i) for (int j = 0; j <= i; ++j) err = fmax(err, fabs(ho[i * TC + j] - ref[i * TC + j]));
    printf("%-36s cycles: %lld %lld %lld %lld %lld %lld   maxdiff_vs_A %.2e  err=%s\n", names[v], ht[0], ht[1], ht[2], ht[3], ht[4], ht[5], err,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
