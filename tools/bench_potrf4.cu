// Micro-benchmark: 32x32 fp64 Cholesky by one warp: blocked-by-4 left-looking (warp_potrf32_b4 of k_linalg.cu) vs the column-by-column version.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_potrf4 bench_potrf4.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
#define TC 32
__device__ __forceinline__ void warp_potrf32(double (*Cs)[TC + 1], double (*Ls)[TC + 1], const double* dorig,
                                             double piv_tol, int lane) {
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
    const bool ok = piv > piv_tol * fabs(dorig[c]) && piv > 0.0;
    const double rs = ok ? rsqrt(piv) : 0.0;
    double l = 0.0;
    if (lane == c) l = piv * rs;
    else if (lane > c) l = v * rs;
    a[c] = l;
    Ls[lane][c] = l;
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < TC; ++k) Cs[lane][k] = a[k];
}
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

__global__ void k(const double* A, double* out, long long* times, int variant, int reps) {
  __shared__ double Cs[TC][TC + 1], Ls[TC][TC + 1];
  __shared__ double dorig[TC], rd[TC];
  const int t = threadIdx.x;
  for (int rep = 0; rep < reps; ++rep) {
    for (int e = t; e < TC * TC; e += blockDim.x) Cs[e >> 5][e & 31] = A[e];
    if (t < TC) dorig[t] = A[t * TC + t];
    __syncthreads();
    const long long t0 = clock64();
    if (variant == 0) { if (t < 32) warp_potrf32(Cs, Ls, dorig, 1e-14, t); }
    else { if (t < 32) warp_potrf32_b4(Cs, dorig, 1e-14, rd, t); }
    __syncthreads();
    const long long t1 = clock64();
    if (t == 0) times[rep] = t1 - t0;
  }
  for (int e = t; e < TC * TC; e += blockDim.x) out[e] = Cs[e >> 5][e & 31];
}
__global__ void krs(double* out, long long* times) {
  double x = out[0];
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) x = rsqrt(x) + 1.5;
  const long long t1 = clock64();
  out[1] = x; times[0] = (t1 - t0) / 64;
}
int main() {
  double hA[TC * TC], L[TC * TC] = {0};
  for (int i = 0; i < TC; ++i) for (int j = 0; j < TC; ++j) hA[i * TC + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1.0 + abs(i - j));
  for (int j = 0; j < TC; ++j) {
    double d = hA[j * TC + j]; for (int k = 0; k < j; ++k) d -= L[j * TC + k] * L[j * TC + k];
    L[j * TC + j] = sqrt(d);
    for (int i = j + 1; i < TC; ++i) { double v = hA[i * TC + j]; for (int k = 0; k < j; ++k) v -= L[i * TC + k] * L[j * TC + k]; L[i * TC + j] = v / L[j * TC + j]; }
  }
  double *dA, *dout; long long* dt;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dout, sizeof(hA)); cudaMalloc(&dt, 64 * 8);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  for (int v = 0; v < 2; ++v) {
    k<<<1, 128>>>(dA, dout, dt, v, 6);
    cudaDeviceSynchronize();
    long long ht[6]; double ho[TC * TC];
    cudaMemcpy(ht, dt, 6 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < TC; ++i) for (int j = 0; j < TC; ++j) err = fmax(err, fabs(ho[i * TC + j] - L[i * TC + j]));
    printf("%-28s cycles: %lld %lld %lld %lld %lld %lld   max|L - host| %.2e  %s\n", v ? "warp_potrf32_b4 (1 warp, nb=4)" : "warp_potrf32 (1 warp)", ht[0], ht[1], ht[2], ht[3], ht[4], ht[5], err,
           cudaGetErrorString(cudaGetLastError()));
  }
  krs<<<1, 1>>>(dout, dt); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, dt, 8, cudaMemcpyDeviceToHost);
  printf("dependent rsqrt(double)+add latency: %lld cycles\n", h);
  return 0;
}
