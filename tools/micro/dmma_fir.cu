// Microbenchmark: the inner loop of the DMMA FIR passes (blur_mma.cu, pass B shape) from shared memory:
// how much of the DMMA pipe does the loop sustain with its operand loads, per warps / SM and block shape?
#include <cstdio>
#include <cuda_runtime.h>
#define PITCH 36
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: operands from registers; 1: samples from smem, weight in a register; 2: samples + weight from smem
template <int MB, int MODE>
__global__ void __launch_bounds__(256, 2) k(double *out, int D, int reps)
{
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  for (int e = threadIdx.x; e < 200 * PITCH + 512; e += 256) sm[e] = 1e-3 * (e % 97);
  __syncthreads();
  const double *sp = sm + (8 * MB * warp + t) * PITCH + g;
  const double *wp = sm + 200 * PITCH + 8 + t - g;
  double acc[MB][4][2];
#pragma unroll
  for (int mb = 0; mb < MB; mb++)
#pragma unroll
    for (int nb = 0; nb < 4; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  double wreg = wp[0], breg = sp[0];
  for (int r = 0; r < reps; r++) {
    for (int d = 0; d < D; d++) {
      const double wv = MODE == 2 ? wp[4 * d] : wreg;
#pragma unroll
      for (int mb = 0; mb < MB; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) dmma(acc[mb][nb][0], acc[mb][nb][1], wv, MODE >= 1 ? sp[(8 * mb + 4 * d) * PITCH + 8 * nb] : breg);
    }
  }
  double s = 0;
#pragma unroll
  for (int mb = 0; mb < MB; mb++)
#pragma unroll
    for (int nb = 0; nb < 4; nb++) s += acc[mb][nb][0] + acc[mb][nb][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MB, int MODE> void run(double *out, int ctas_per_sm, int D)
{
  const int reps = 200, smem = (200 * PITCH + 512) * 8;
  cudaFuncSetAttribute(k<MB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MB, MODE><<<148 * ctas_per_sm, 256, smem>>>(out, D, reps); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MB, MODE><<<148 * ctas_per_sm, 256, smem>>>(out, D, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n_dmma = 148.0 * ctas_per_sm * 8 * reps * D * MB * 4;
  printf("MB %d mode %d ctas/SM %d D %2d: %.3f ms  %.2f TFLOP/s (%.0f%% of 37.0)\n", MB, MODE, ctas_per_sm, D, ms, n_dmma * 512 / ms / 1e9,
         n_dmma * 512 / ms / 1e9 / 37.0 * 100);
}
int main()
{
  double *out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  for (int c : {1, 2}) for (int D : {6, 17}) {
    run<1, 0>(out, c, D); run<1, 1>(out, c, D); run<1, 2>(out, c, D);
    run<2, 0>(out, c, D); run<2, 1>(out, c, D); run<2, 2>(out, c, D);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
