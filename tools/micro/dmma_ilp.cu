// Microbenchmark: DMMA.8x8x4 throughput on sm_100a as a function of warps per SM sub-partition and of the
// number of independent accumulator chains per warp (latency / issue spacing of the fp64 matrix instruction).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double *out, double a, double b)
{
  double c[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; i++) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8 / ILP; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 2 * ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(double *out, int warps_per_sm)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 32 * warps_per_sm;
  k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n_dmma = 148.0 * warps_per_sm * ITERS * 8;
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("warps/SM %2d  ILP %d : %.3f ms  %.2f TFLOP/s  cycles per DMMA per sub-partition %.1f  (per warp %.1f)\n", warps_per_sm, ILP, ms,
         n_dmma * 512 / ms / 1e9, cyc / (n_dmma / 148 / 4), cyc / (ITERS * 8.0));
}
int main()
{
  double *out; cudaMalloc(&out, 148 * 1024 * 8);
  for (int w : {4, 8, 16, 32}) { run<1>(out, w); run<2>(out, w); run<4>(out, w); run<8>(out, w); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
