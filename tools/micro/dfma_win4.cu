// Sliding window of 4 positions x 2 phases (8 accumulators, ~64 registers): how close to the fp64 peak with 4..8
// warps per sub-partition?  Weights from shared memory (LDS.128 broadcast) or from the constant bank (kernel params).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2000
struct W { double w[40]; };
#define STEP4(U, JJ) { const double c0 = WX(2 * (JJ)), c1 = WX(2 * (JJ) + 1); _Pragma("unroll") for (int k = 0; k < 4; k++) { a0[k] = fma(c0, vw[(k + (U)) & 3], a0[k]); a1[k] = fma(c1, vw[(k + (U)) & 3], a1[k]); } vw[(U) & 3] = nxt[(JJ) * STRIDE]; }
template <int STRIDE, bool CONSTW>
__global__ void __launch_bounds__(1024, 1) k_win4(double *out, const __grid_constant__ W cw, int np, long long *cyc)
{
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 64 * 49 + 64; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const double *ws = sm + 64 * 49;
#define WX(i) (CONSTW ? cw.w[(i)] : ws[(i)])
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    double a0[4], a1[4], vw[4];
    const int lane = threadIdx.x & 63;
    const double *base = STRIDE == 1 ? sm + lane * 49 + (it & 7) : sm + (it & 7) * 49 + (lane % 48);
#pragma unroll
    for (int k = 0; k < 4; k++) vw[k] = base[k * STRIDE];
    const double *nxt = base + 4 * STRIDE;
    { const double c0 = WX(0), c1 = WX(1);
#pragma unroll
      for (int k = 0; k < 4; k++) { a0[k] = c0 * vw[k]; a1[k] = c1 * vw[k]; }
      vw[0] = nxt[0]; }
    int j = 1;
    for (; j + 4 <= np; j += 4) { STEP4(1, j) STEP4(2, j + 1) STEP4(3, j + 2) STEP4(0, j + 3) }
    const int rem = np - j;
    if (rem & 2) { STEP4(1, j) STEP4(2, j + 1) }
    if (rem & 1) { if (rem & 2) STEP4(3, j + 2) else STEP4(1, j) }
#pragma unroll
    for (int k = 0; k < 4; k++) acc += a0[k] + a1[k];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
  double *out; long long *cyc, h[148];
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  W w; for (int i = 0; i < 40; i++) w.w[i] = 0.01 * i;
  for (int wps : { 2, 4, 6, 8 }) {
    const int threads = 128 * wps;
    auto report = [&](const char *name, int np) {
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
      printf("%-34s np=%2d %d warps/SMSP: %.1f DFMA/clk/SM (+DADD: %.1f)\n", name, np, wps, (double)ITERS * np * 8 * threads / c,
             (double)ITERS * (np * 8 + 8) * threads / c);
    };
    for (int np : { 7, 11, 17 }) {
      k_win4<1, false><<<148, threads, 26 * 1024>>>(out, w, np, cyc); k_win4<1, false><<<148, threads, 26 * 1024>>>(out, w, np, cyc);
      report("4x2 window stride 1, smem weights", np);
      k_win4<1, true><<<148, threads, 26 * 1024>>>(out, w, np, cyc); k_win4<1, true><<<148, threads, 26 * 1024>>>(out, w, np, cyc);
      report("4x2 window stride 1, const weights", np);
    }
  }
  return 0;
}
