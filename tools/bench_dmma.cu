// Micro-benchmark: fp64 tensor-core MMA (mma.sync.m8n8k4.f64) vs DFMA on sm_100a: throughput per SM and dependent latency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_dmma tools/bench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4);
  const int iters = 20000;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("sm clock (attr) %d kHz\n", clk);
  for (int warps : {1, 2, 4, 8, 16}) {
    float m1 = timeit([&] { k_dmma<1><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); });
    float m4 = timeit([&] { k_dmma<4><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); });
    float m8 = timeit([&] { k_dmma<8><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); });
    float f1 = timeit([&] { k_dfma<1><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); });
    float f8 = timeit([&] { k_dfma<8><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); });
    auto tf_mma = [&](float ms, int ilp) { return 148.0 * warps * ilp * (double)iters * 256 * 2 / (ms * 1e-3) / 1e12; };
    auto tf_fma = [&](float ms, int ilp) { return 148.0 * warps * ilp * (double)iters * 32 * 2 / (ms * 1e-3) / 1e12; };
    printf("warps/SM %2d | DMMA ilp1 %6.2f TF (%.0f ns/instr chain) ilp4 %6.2f TF ilp8 %6.2f TF | DFMA ilp1 %6.2f TF (%.1f ns chain) ilp8 %6.2f TF\n",
           warps, tf_mma(m1, 1), m1 * 1e6 / iters, tf_mma(m4, 4), tf_mma(m8, 8), tf_fma(f1, 1), f1 * 1e6 / iters, tf_fma(f8, 8));
  }
  return 0;
}
