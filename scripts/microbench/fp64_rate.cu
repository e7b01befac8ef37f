// FP64 throughput of one B200 SM: vector DFMA vs the fp64 tensor-core MMA (mma.sync m8n8k4 f64), both with enough
// independent accumulators to hide latency.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* d;
  cudaMalloc(&d, 148 * 1024 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int grid : {1, 148}) {
      float ms;
      dfma_kernel<8><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e0);
      dfma_kernel<8><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)grid * threads * 8.0 * iters;
      printf("DFMA  grid %3d threads %4d: %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM at %d MHz nominal\n", grid, threads, ms, fma / ms / 1e6,
             fma / ms / 1e3 / grid / clk_khz, clk_khz / 1000);
      dmma_kernel<8><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e0);
      dmma_kernel<8><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      double mf = (double)grid * (threads / 32) * 8.0 * iters * 256.0;
      printf("DMMA  grid %3d threads %4d: %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM\n", grid, threads, ms, mf / ms / 1e6, mf / ms / 1e3 / grid / clk_khz);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
