// fp64 throughput probe on B200: DFMA (CUDA cores) vs DMMA (mma.sync m8n8k4 f64)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
  double c0[4][2];
  for (int i = 0; i < 4; ++i) { c0[i][0] = 0; c0[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i][0]), "+d"(c0[i][1]) : "d"(a), "d"(b));
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0][0] + c0[1][1] + c0[2][0] + c0[3][1];
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    int iters = 20000;
    cudaEventRecord(e0); k_dfma<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * iters * 148.0 * 8 * 256;
    printf("DFMA: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaEventRecord(e0); k_dmma<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 8 * 8 * 4 * 4 * (double)iters * 148.0 * 8 * 8;  // per warp per mma: 8x8x4 MACs
    printf("DMMA: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
