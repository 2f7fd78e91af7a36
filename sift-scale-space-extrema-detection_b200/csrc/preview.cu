// preview.cu -- the reference's display products of a pyramid level, as an RGBA8 image (ImageData layout).
//
// The reference posts, beside every stage result, an ImageData for its canvas:
//   GRAY     ImageUtils_convertMatrix2DToImageData(grayChannelMatrix) : Math.round(v * 255) into a
//            Uint8ClampedArray, alpha 255 (image-utils.js:171-220; Gaussian levels, background.js:194-220)
//   SIGMOID  Matrix2D_sigmoidNormalize(level, coefficient) = 1 / (1 + exp(coefficient * (-1 * v)))
//            (matrix2d.js:148-156), coefficient 5 for DoG chunks (background.js:303-307), then GRAY
//   MINMAX   Matrix2D_sampledNormalize(level) = (v - min) / (max - min) over the whole level
//            (matrix2d.js:169-192; DoG images, background.js:336), then GRAY
// Levels are held as fp32 here (fp64 in the reference): a preview byte can differ by one grey level where
// v * 255 falls within ~1e-5 of a rounding boundary.  Not on the detection path; nothing here is timed.
#include <cfloat>
#include "common.cuh"

__device__ __forceinline__ unsigned order_f32(float f)           // monotone float -> unsigned map
{
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(unsigned k)
{
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// mm[0] = ordered min, mm[1] = ordered max (initialised to 0xffffffff / 0 by the launcher)
__global__ void preview_minmax_kernel(const float *__restrict__ src, int w, int h, size_t pitch, unsigned *mm)
{
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int y = blockIdx.x; y < h; y += gridDim.x)
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
      const float v = src[(size_t)y * pitch + x];
      lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], order_f32(lo)); atomicMax(&mm[1], order_f32(hi)); }
}

__device__ __forceinline__ unsigned char round_u8(double x)
{
  // Math.round(x * 255): nearest integer, ties toward +infinity (t - floor(t) is exact); a Uint8ClampedArray
  // store then clamps to 0..255 and maps NaN to 0
  const double t = x * 255.0, f = floor(t);
  const double r = (t - f >= 0.5) ? f + 1.0 : f;
  return (unsigned char)(r >= 255.0 ? 255 : (r >= 0.0 ? (int)r : 0));   // NaN fails both compares -> 0
}

__global__ void preview_kernel(const float *__restrict__ src, int w, int h, size_t pitch, int mode, double coefficient,
                               const unsigned *__restrict__ mm, uchar4 *__restrict__ out, double *mm_out)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  double lo = 0.0, hi = 1.0;
  if (mode == SIFT_PREVIEW_MINMAX) { lo = (double)unorder_f32(mm[0]); hi = (double)unorder_f32(mm[1]); }
  if (mm_out && x == 0 && y == 0) { mm_out[0] = lo; mm_out[1] = hi; }
  if (x >= w || y >= h) return;
  const double v = (double)src[(size_t)y * pitch + x];
  double t = v;
  if (mode == SIFT_PREVIEW_SIGMOID) t = 1.0 / (1.0 + exp(coefficient * (-1.0 * v)));
  else if (mode == SIFT_PREVIEW_MINMAX) t = (v - lo) / (hi - lo);
  const unsigned char g = round_u8(t);
  out[(size_t)y * w + x] = make_uchar4(g, g, g, 255);
}

// scratch: 2 unsigned + 2 doubles (32 bytes); out: w*h uchar4 on the device.  Returns the launches made.
int launch_preview(cudaStream_t st, const float *src, int w, int h, size_t pitch, int mode, double coefficient,
                   void *scratch32, void *d_out)
{
  unsigned *mm = (unsigned *)scratch32;
  double *mm_out = (double *)((char *)scratch32 + 16);
  int launches = 1;
  if (mode == SIFT_PREVIEW_MINMAX) {
    const unsigned init[2] = { 0xffffffffu, 0u };
    cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, st);
    preview_minmax_kernel<<<min(h, 1184), 256, 0, st>>>(src, w, h, pitch, mm);
    launches++;
  }
  dim3 grid((w + 255) / 256, h);
  preview_kernel<<<grid, 256, 0, st>>>(src, w, h, pitch, mode, coefficient, mm, (uchar4 *)d_out, mm_out);
  return launches;
}
