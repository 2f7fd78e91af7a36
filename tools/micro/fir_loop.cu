// Microbenchmark: the register-window FIR tap loop of blur_sep.cu in isolation (no staging, no epilogue):
// what fraction of the fp64 pipe can the loop itself sustain?   nvcc -O3 -arch=sm_100a fir_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
#define FP_PITCH 33
template <int NO>
__device__ __forceinline__ void fir_window(const double *__restrict__ w, const int npad,
                                           const double *__restrict__ base, double (&a)[NO])
{
  double vw[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) { vw[k] = base[k * FP_PITCH]; a[k] = 0.0; }
  const double *nxt = base + NO * FP_PITCH;
  double2 wa = *reinterpret_cast<const double2 *>(w), wb = *reinterpret_cast<const double2 *>(w + 2);
#define FIR_GROUP(G)                                                                              \
  {                                                                                               \
    const double c[4] = { wa.x, wa.y, wb.x, wb.y };                                               \
    wa = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 4);                                 \
    wb = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 6);                                 \
    _Pragma("unroll") for (int u = 0; u < 4; u++) {                                               \
      _Pragma("unroll") for (int k = 0; k < NO; k++) a[k] = fma(c[u], vw[(k + 4 * (G) + u) & (NO - 1)], a[k]); \
      vw[(4 * (G) + u) & (NO - 1)] = nxt[(j + 4 * (G) + u) * FP_PITCH];                            \
    }                                                                                             \
  }
  int j = 0;
  for (; j + NO <= npad; j += NO) {
    FIR_GROUP(0) FIR_GROUP(1)
    if (NO == 16) { FIR_GROUP(2) FIR_GROUP(3) }
  }
}
template <int NO, int MINB>
__global__ void __launch_bounds__(256, MINB) k(double *out, int npad, int reps)
{
  extern __shared__ double smem[];
  double *wsm = smem, *tile = smem + 256;
  for (int i = threadIdx.x; i < 256; i += 256) wsm[i] = 1.0 / (1 + i);
  for (int i = threadIdx.x; i < 400 * FP_PITCH; i += 256) tile[i] = i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0;
  for (int r = 0; r < reps; r++) {
    double acc[NO];
    fir_window<NO>(wsm, npad, tile + (warp * NO + (r & 7)) * FP_PITCH + lane, acc);
#pragma unroll
    for (int k = 0; k < NO; k++) s += acc[k];
  }
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int NO, int MINB> void run(const char *name, int npad)
{
  double *out; cudaMalloc(&out, 148 * MINB * 256 * 8);
  const int reps = 400, smem = (256 + 400 * FP_PITCH) * 8;
  cudaFuncSetAttribute(k<NO, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NO, MINB><<<148 * MINB, 256, smem>>>(out, npad, reps); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NO, MINB><<<148 * MINB, 256, smem>>>(out, npad, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = 148.0 * MINB * 256 * reps * (double)npad * NO;
  printf("%-28s npad %3d: %.3f ms  %.2f TFLOP/s (%.0f%% of 36.6)  %s\n", name, npad, ms, 2 * dfma / ms / 1e9, 2 * dfma / ms / 1e9 / 36.6 * 100,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main()
{
  for (int npad : { 16, 32, 64 }) {
    run<8, 2>("NO=8  2 CTA/SM", npad);
    run<8, 3>("NO=8  3 CTA/SM", npad);
    run<8, 4>("NO=8  4 CTA/SM", npad);
    run<16, 1>("NO=16 1 CTA/SM", npad);
    run<16, 2>("NO=16 2 CTA/SM", npad);
  }
  return 0;
}
