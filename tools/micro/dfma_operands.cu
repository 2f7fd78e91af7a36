// fp64 pipe rate vs operand pattern (registers only, 4 warps per sub-partition):
//  A: x = fma(x, a, b)            one fresh register operand (x), a and b held in the reuse cache
//  B: acc[k] = fma(c, v[k], acc[k])   two fresh operands per DFMA (v[k], acc[k]), c reused   (the tap-loop pattern)
//  C: acc0[k] = fma(c0, v[k], acc0[k]); acc1[k] = fma(c1, v[k], acc1[k])  interleaved: v[k] reused, c alternates
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 3000
template <int MODE>
__global__ void k(double *out, double a, double b, long long *cyc)
{
  double acc0[8], acc1[8], v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { acc0[i] = threadIdx.x * 1e-3 + i; acc1[i] = acc0[i] + 1; v[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
  double c0 = a, c1 = b;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) { acc0[i] = fma(acc0[i], a, b); acc1[i] = fma(acc1[i], a, b); }
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc0[i] = fma(c0, v[(i + r) & 7], acc0[i]);
#pragma unroll
        for (int i = 0; i < 8; i++) acc1[i] = fma(c1, v[(i + r) & 7], acc1[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) { acc0[i] = fma(c0, v[(i + r) & 7], acc0[i]); acc1[i] = fma(c1, v[(i + r) & 7], acc1[i]); }
      }
      c0 += 1e-12; c1 -= 1e-12;       // 2 DADD per 16 DFMA: keeps the weights from being loop invariants
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc0[i] + acc1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
  double *out; long long *cyc, h[148];
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  for (int wps : { 1, 2, 4 }) {
    const int threads = 128 * wps;
    auto report = [&](const char *name) {
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
      printf("%-44s %d warps/SMSP: %.1f DFMA/clk/SM\n", name, wps, (double)ITERS * 8 * 16 * threads / c);
    };
    k<0><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); k<0><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); report("A one fresh operand");
    k<1><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); k<1><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); report("B two fresh operands, weight reused");
    k<2><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); k<2><<<148, threads>>>(out, 1.0000001, 1e-9, cyc); report("C two fresh, sample reused, weights alternate");
  }
  return 0;
}
