// Microbenchmark: store-only bandwidth of the B200 (what a kernel that only WRITES its planes can reach), against
// the copy bandwidth MEASURED_PEAKS.json quotes (read + write).  Patterns: one contiguous stream, and 11 planes
// written tile by tile (64 x 64 floats per CTA and plane, 16-byte stores covering 8 rows x 64 bytes per warp
// instruction) like the octave-0 blur kernel does.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_stream(float4 *out, size_t n)
{
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_copy(const float4 *in, float4 *out, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
// 11 planes of w x h floats; CTA = 64 x 64 tile of every plane; warp = 32 rows x 16 columns, lane (g, t) stores float4 at row g + 8 mb
__global__ void __launch_bounds__(256) k_tiles(float *base, int w, int h, size_t plane)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int x = blockIdx.x * 64 + 16 * (warp & 3) + 4 * t, y0 = blockIdx.y * 64 + 32 * (warp >> 2) + g;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int p = 0; p < 11; p++)
#pragma unroll
    for (int mb = 0; mb < 4; mb++) *reinterpret_cast<float4 *>(base + p * plane + (size_t)(y0 + 8 * mb) * w + x) = v;
}
template <typename F> float timeit(F f, int reps = 10)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < reps; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main()
{
  const int w = 3840, h = 2176;                       // octave 0 of a 1920 x 1080 frame (rounded to tiles)
  const size_t plane = (size_t)w * h, bytes = plane * 11 * 4;
  float *buf, *src; cudaMalloc(&buf, bytes * 4); cudaMalloc(&src, bytes);
  cudaMemset(src, 0, bytes);
  float t;
  t = timeit([&] { k_stream<<<148 * 8, 256>>>((float4 *)buf, bytes / 16); });
  printf("contiguous stores, %zu MB: %.3f ms  %.0f GB/s\n", bytes >> 20, t, bytes / t / 1e6);
  t = timeit([&] { k_tiles<<<dim3(w / 64, h / 64), 256>>>(buf, w, h, plane); });
  printf("11 planes tile by tile (octave-0 pattern), %zu MB: %.3f ms  %.0f GB/s\n", bytes >> 20, t, bytes / t / 1e6);
  // rotate over 4 pyramids like 4 frames in flight (no L2 write hits on the same lines)
  int r = 0;
  t = timeit([&] { k_tiles<<<dim3(w / 64, h / 64), 256>>>(buf + (size_t)(r++ & 3) * (bytes / 4), w, h, plane); }, 12);
  printf("same, rotating over 4 pyramids: %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = timeit([&] { cudaMemsetAsync(buf, 0, bytes); });
  printf("cudaMemsetAsync: %.3f ms  %.0f GB/s\n", t, bytes / t / 1e6);
  t = timeit([&] { k_copy<<<148 * 8, 256>>>((const float4 *)src, (float4 *)buf, bytes / 16); });
  printf("copy (read + write bytes): %.3f ms  %.0f GB/s\n", t, 2.0 * bytes / t / 1e6);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
