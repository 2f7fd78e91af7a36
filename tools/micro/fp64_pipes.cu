// Microbenchmark: fp64 FMA pipe vs fp64 tensor (mma.sync.m8n8k4.f64) on sm_100a -- separate or shared?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_dfma(double *out, double a, double b)
{
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double *out, double a, double b)
{
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_both(double *out, double a, double b)
{
  double c[16], x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i] = threadIdx.x * 1e-3 + i; x[i] = c[i] + 1; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) { dmma(c[i], c[i + 1], a, b); x[i] = fma(x[i], a, b); x[i + 1] = fma(x[i + 1], a, b); }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8 * 2);
  const int blocks = 148 * 4, threads = 512;
  float t1 = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  float t2 = timeit([&] { k_dmma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  float t3 = timeit([&] { k_both<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  const double nthreads = (double)blocks * threads;
  const double dfma_flops = nthreads * ITERS * 16 * 2;
  const double dmma_flops = (nthreads / 32) * ITERS * 8 * (8 * 8 * 4 * 2);
  printf("dfma only : %.3f ms  %.2f TFLOP/s\n", t1, dfma_flops / t1 / 1e9);
  printf("dmma only : %.3f ms  %.2f TFLOP/s\n", t2, dmma_flops / t2 / 1e9);
  printf("both      : %.3f ms  dfma %.2f + dmma %.2f = %.2f TFLOP/s\n", t3, dfma_flops / t3 / 1e9, dmma_flops / t3 / 1e9,
         (dfma_flops + dmma_flops) / t3 / 1e9);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
