// How many warps per SM sub-partition does the fp64 pipe need?  (a) register-only DFMA chains with 16 independent
// accumulators, (b) the two-phase sliding-window tap loop of blur_oct0.cu (1 LDS.64 + 1 LDS.128 per 16 DFMA).
// Prints DFMA per clock per SM for 1..4 warps per sub-partition.   nvcc -arch=sm_100a dfma_warps.cu -o dfma_warps
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2000
__global__ void k_reg(double *out, double a, double b, long long *cyc)
{
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 11; r++)
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
#define STEP(U, JJ) { const double2 c = w2[(JJ)]; _Pragma("unroll") for (int k = 0; k < 8; k++) { a0[k] = fma(c.x, vw[(k + (U)) & 7], a0[k]); a1[k] = fma(c.y, vw[(k + (U)) & 7], a1[k]); } vw[(U) & 7] = nxt[(JJ) * STRIDE]; }
template <int STRIDE>
__device__ __forceinline__ void window(const double *base, const double2 *w2, int np, double (&a0)[8], double (&a1)[8])
{
  double vw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) vw[k] = base[k * STRIDE];
  const double *nxt = base + 8 * STRIDE;
  { const double2 c = w2[0];
#pragma unroll
    for (int k = 0; k < 8; k++) { a0[k] = c.x * vw[k]; a1[k] = c.y * vw[k]; }
    vw[0] = nxt[0]; }
  int j = 1;
  for (; j + 8 <= np; j += 8) { STEP(1, j) STEP(2, j + 1) STEP(3, j + 2) STEP(4, j + 3) STEP(5, j + 4) STEP(6, j + 5) STEP(7, j + 6) STEP(0, j + 7) }
  const int rem = np - j;
  if (rem & 4) { STEP(1, j) STEP(2, j + 1) STEP(3, j + 2) STEP(4, j + 3) }
  if (rem & 2) { if (rem & 4) { STEP(5, j + 4) STEP(6, j + 5) } else { STEP(1, j) STEP(2, j + 1) } }
  if (rem & 1) { switch (rem & 6) { case 0: STEP(1, j) break; case 2: STEP(3, j + 2) break; case 4: STEP(5, j + 4) break; default: STEP(7, j + 6) break; } }
}
template <int STRIDE>
__global__ void k_win(double *out, int np, long long *cyc)
{
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 64 * 49 + 64; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const double2 *w2 = reinterpret_cast<const double2 *>(sm + 64 * 49);
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    double a0[8], a1[8];
    const int lane = threadIdx.x & 63;
    window<STRIDE>(STRIDE == 1 ? sm + lane * 49 + (it & 7) : sm + (it & 7) * 49 + (lane % 48), w2, np, a0, a1);
#pragma unroll
    for (int k = 0; k < 8; k++) acc += a0[k] + a1[k];     // 16 DADD per window: counted as overhead
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
  double *out; long long *cyc, h[148];
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k_win<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_win<49>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int wps = 1; wps <= 4; wps++) {
    const int threads = 128 * wps;
    auto report = [&](const char *name, double dfma_per_thread) {
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
      printf("%-28s %d warps/SMSP: %.1f DFMA/clk/SM (%.0f cycles)\n", name, wps, dfma_per_thread * threads / c, c);
    };
    k_reg<<<148, threads>>>(out, 1.0000001, 1e-9, cyc); k_reg<<<148, threads>>>(out, 1.0000001, 1e-9, cyc);
    report("register chains x16", (double)ITERS * 11 * 16);
    for (int np : { 7, 11, 17 }) {
      char nm[64];
      k_win<1><<<148, threads, 26 * 1024>>>(out, np, cyc); k_win<1><<<148, threads, 26 * 1024>>>(out, np, cyc);
      snprintf(nm, sizeof nm, "window stride 1, np=%d", np); report(nm, (double)ITERS * np * 16);
      k_win<49><<<148, threads, 26 * 1024>>>(out, np, cyc); k_win<49><<<148, threads, 26 * 1024>>>(out, np, cyc);
      snprintf(nm, sizeof nm, "window stride 49, np=%d", np); report(nm, (double)ITERS * np * 16);
    }
  }
  return 0;
}
