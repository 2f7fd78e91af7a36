// blur_generic.cu -- radius-generic separable Gaussian scale-space kernels.
//
// Restates (as a separable fp64 computation) the reference's dense clamped 2D
// correlation SIFT_blurMatrix2DChunk (src/sift.js:72-149) applied level by level
// to the octave base (background.js:103-224), the DoG subtraction
// SIFT_subtractMatrix2DChunk (src/sift.js:154-188, finer minus coarser) and the
// 2x decimation that seeds the next octave (matrix2d.js:112-138 with rate 2.0,
// background.js:114-130).  The 2x nearest-neighbour upsample of the input
// (background.js:84) is folded into the octave-0 loads.
//
// Two passes through an fp64 intermediate T (L2 resident for the small, high
// octaves this path is kept for; see blur_fused.cu for octaves 0 and 1):
//   hblur: T_s[y][x]  = sum_i w_s[i] * base[y][clamp(x + i - R_s)]
//   vblur: G_s[y][x]  = sum_j w_s[j] * T_s[clamp(y + j - R_s)][x]
//          D_{s-1}    = G_{s-1} - G_s  formed from the UNROUNDED fp64 accumulators,
//          both stored fp32; level `spo` is also decimated into the next seed (fp64 + fp32).
// Accumulation is fp64 in both passes (fp32 accumulation fails the
// 1e-5 / 1e-3 px parity bars).  Each thread produces NOUT=8 neighbouring outputs
// from a sliding window so every loaded sample feeds 8 DFMAs.
#include "common.cuh"

#define NOUT 8

struct BlurLevels {
  int nlev;                       // number of blurred levels handled
  int level[SIFT_MAX_LEVELS];     // pyramid level index of each
  int radius[SIFT_MAX_LEVELS];
  int woff[SIFT_MAX_LEVELS];
  double *T[SIFT_MAX_LEVELS];
};

// acc[k] += sum_t w[t-k] * v(t),  t = 0 .. 2R+NOUT-1; wpad holds w[0..2R] then >= SIFT_WPAD zeros.
template <typename LoadV>
__device__ __forceinline__ void conv_window(const double *__restrict__ wpad, int R, LoadV loadv, double acc[NOUT])
{
  double wr[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) { wr[k] = 0.0; acc[k] = 0.0; }
  const int T = 2 * R + NOUT;
  for (int t0 = 0; t0 < T; t0 += NOUT) {
#pragma unroll
    for (int u = 0; u < NOUT; u++) {
      const int t = t0 + u;
      const double v = loadv(t);
      wr[u] = __ldg(wpad + t);
#pragma unroll
      for (int k = 0; k < NOUT; k++) acc[k] = fma(wr[(u - k + NOUT) % NOUT], v, acc[k]);
    }
  }
}

// ---- source pixel -> double, exactly the reference's float64 image value ------
template <int DTYPE> struct SrcLoad;
template <> struct SrcLoad<SIFT_U8> {
  static __device__ __forceinline__ double at(const void *row, int x) {
    return (double)((const unsigned char *)row)[x] / 255.0;                     // image-utils.js:114
  }
};
template <> struct SrcLoad<SIFT_F32> {
  static __device__ __forceinline__ double at(const void *row, int x) { return (double)((const float *)row)[x]; }
};
template <> struct SrcLoad<SIFT_F64> {
  static __device__ __forceinline__ double at(const void *row, int x) { return ((const double *)row)[x]; }
};
template <> struct SrcLoad<SIFT_RGBA8> {
  static __device__ __forceinline__ double at(const void *row, int x) {
    const uchar4 p = ((const uchar4 *)row)[x];
    // (R*0.299) + (G*0.587) + (B*0.114), then / 255.0 -- image-utils.js:107-114, unfused
    const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)p.x, 0.299), __dmul_rn((double)p.y, 0.587)),
                               __dmul_rn((double)p.z, 0.114));
    return g / 255.0;
  }
};

// ------------------------------------------------------------------- hblur ----
#define HB_ROWS 4
#define HB_THREADS_X 32
#define HB_TW (HB_THREADS_X * NOUT)   // 256 outputs per row per CTA

__device__ __forceinline__ int padidx(int e) { return e + (e >> 3); }   // 9-double stride per 8: conflict-free LDS.64

template <int DTYPE, int UPS>
__global__ void __launch_bounds__(HB_THREADS_X *HB_ROWS)
hblur_kernel(const void *__restrict__ src, size_t src_pitch, int src_w, int w, int hrows, int rmax,
             const double *__restrict__ weights, BlurLevels L)
{
  extern __shared__ double smem[];
  const int span = HB_TW + 2 * rmax + NOUT;          // samples staged per row (NOUT slack for the window tail)
  const int rowstride = padidx(span) + 1;
  const int x_tile = blockIdx.x * HB_TW;
  const int row0 = blockIdx.y * HB_ROWS;
  const int tid = threadIdx.y * HB_THREADS_X + threadIdx.x;

  for (int r = 0; r < HB_ROWS; r++) {
    const int y = row0 + r;
    if (y >= hrows) break;
    const char *rowp = (const char *)src + (size_t)y * src_pitch;
    for (int e = tid; e < span; e += HB_THREADS_X * HB_ROWS) {
      int col = x_tile - rmax + e;
      col = min(max(col, 0), w - 1);                                           // sift.js:116-117 clamp
      const int sc = UPS ? (col >> 1) : col;                                   // matrix2d.js:129 floor(j*0.5)
      smem[r * rowstride + padidx(e)] = SrcLoad<DTYPE>::at(rowp, min(sc, src_w - 1));
    }
  }
  __syncthreads();

  const int y = row0 + threadIdx.y;
  if (y >= hrows) return;
  const int x0 = x_tile + threadIdx.x * NOUT;
  if (x0 >= w) return;
  const double *rowsm = smem + threadIdx.y * rowstride;
  for (int li = 0; li < L.nlev; li++) {
    const int R = L.radius[li];
    const int e0 = threadIdx.x * NOUT + rmax - R;     // smem sample feeding tap 0 of output 0
    double acc[NOUT];
    conv_window(weights + L.woff[li], R, [&](int t) { return rowsm[padidx(e0 + t)]; }, acc);
    double *out = L.T[li] + (size_t)y * w + x0;
    if (x0 + NOUT <= w && ((w & 1) == 0)) {
#pragma unroll
      for (int k = 0; k < NOUT; k += 2) *reinterpret_cast<double2 *>(out + k) = make_double2(acc[k], acc[k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < NOUT; k++) if (x0 + k < w) out[k] = acc[k];
    }
  }
}

template <int DTYPE>
static void hblur_dispatch(cudaStream_t st, const void *src, size_t src_pitch, int src_w, int upsample, int w,
                           int hrows, int rmax, const double *d_weights, const BlurLevels &L)
{
  dim3 block(HB_THREADS_X, HB_ROWS);
  dim3 grid((w + HB_TW - 1) / HB_TW, (hrows + HB_ROWS - 1) / HB_ROWS);
  const int span = HB_TW + 2 * rmax + NOUT;
  const size_t smem = (size_t)HB_ROWS * ((span + (span >> 3)) + 1) * sizeof(double);
  if (upsample) {
    cudaFuncSetAttribute(hblur_kernel<DTYPE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    hblur_kernel<DTYPE, 1><<<grid, block, smem, st>>>(src, src_pitch, src_w, w, hrows, rmax, d_weights, L);
  } else {
    cudaFuncSetAttribute(hblur_kernel<DTYPE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    hblur_kernel<DTYPE, 0><<<grid, block, smem, st>>>(src, src_pitch, src_w, w, hrows, rmax, d_weights, L);
  }
}

static BlurLevels make_levels(const LevelPlan *plans, int first_level, int nlev, double *const *T, int *rmax)
{
  BlurLevels L;
  L.nlev = nlev - first_level;
  *rmax = 0;
  for (int i = 0; i < L.nlev; i++) {
    const LevelPlan &p = plans[first_level + i];
    L.level[i] = first_level + i;
    L.radius[i] = p.radius;
    L.woff[i] = p.woff;
    L.T[i] = T[i];
    if (p.radius > *rmax) *rmax = p.radius;
  }
  return L;
}

void launch_hblur(cudaStream_t st, const void *src, int dtype, size_t src_pitch_bytes, int src_w, int src_h,
                  int upsample, int w, int hrows, const double *d_weights, const LevelPlan *plans, int first_level,
                  int nlev, double *const *T, double **)
{
  (void)src_h;
  int rmax;
  BlurLevels L = make_levels(plans, first_level, nlev, T, &rmax);
  switch (dtype) {
    case SIFT_U8: hblur_dispatch<SIFT_U8>(st, src, src_pitch_bytes, src_w, upsample, w, hrows, rmax, d_weights, L); break;
    case SIFT_F32: hblur_dispatch<SIFT_F32>(st, src, src_pitch_bytes, src_w, upsample, w, hrows, rmax, d_weights, L); break;
    case SIFT_F64: hblur_dispatch<SIFT_F64>(st, src, src_pitch_bytes, src_w, upsample, w, hrows, rmax, d_weights, L); break;
    case SIFT_RGBA8: hblur_dispatch<SIFT_RGBA8>(st, src, src_pitch_bytes, src_w, upsample, w, hrows, rmax, d_weights, L); break;
  }
}

// ------------------------------------------------------------------- vblur ----
#define VB_THREADS 128

struct VBlurArgs {
  OctaveDev oct;
  OctaveDev next;        // valid when has_next
  int has_next, spo, keep_gauss, seed_is_level0;
  BlurLevels L;
};

template <int UPS>
__global__ void __launch_bounds__(VB_THREADS)
vblur_kernel(const double *__restrict__ weights, VBlurArgs A)
{
  const int x = blockIdx.x * VB_THREADS + threadIdx.x;
  const int y0 = blockIdx.y * NOUT;
  const int w = A.oct.w, h = A.oct.h, pitch = A.oct.pitch;
  if (x >= w) return;

  double prev[NOUT];
  if (A.seed_is_level0) {                               // octaves >= 1: level 0 is the unblurred seed
#pragma unroll
    for (int k = 0; k < NOUT; k++) prev[k] = A.oct.seed64[(size_t)min(y0 + k, h - 1) * w + x];
  }
  for (int li = 0; li < A.L.nlev; li++) {
    const int R = A.L.radius[li];
    const int s = A.L.level[li];
    const double *__restrict__ T = A.L.T[li];
    double acc[NOUT];
    conv_window(weights + A.L.woff[li], R, [&](int t) {
      int yy = min(max(y0 + t - R, 0), h - 1);                                 // sift.js:118-119 clamp
      if (UPS) yy >>= 1;                                                       // rows 2b and 2b+1 are equal
      return __ldg(T + (size_t)yy * w + x);
    }, acc);
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      const int y = y0 + k;
      if (y < h) {
        if (A.keep_gauss) A.oct.gauss[s][(size_t)y * pitch + x] = (float)acc[k];
        if (s > 0) A.oct.dog[s - 1][(size_t)y * pitch + x] = (float)(prev[k] - acc[k]);   // sift.js:172
        if (A.has_next && s == A.spo && ((y | x) & 1) == 0) {                  // matrix2d.js:129 in[2a][2b]
          A.next.seed64[(size_t)(y >> 1) * A.next.w + (x >> 1)] = acc[k];
          A.next.gauss[0][(size_t)(y >> 1) * A.next.pitch + (x >> 1)] = (float)acc[k];
        }
      }
      prev[k] = acc[k];
    }
  }
}

void launch_vblur(cudaStream_t st, int upsample, const OctaveDev &oct, const double *d_weights,
                  const LevelPlan *plans, int first_level, double *const *T, const OctaveDev *next, int spo,
                  int keep_gauss)
{
  VBlurArgs A;
  A.oct = oct;
  A.has_next = next ? 1 : 0;
  if (next) A.next = *next; else A.next = oct;
  A.spo = spo;
  A.keep_gauss = keep_gauss;
  A.seed_is_level0 = first_level > 0;
  int rmax;
  A.L = make_levels(plans, first_level, oct.nlev, T, &rmax);
  dim3 grid((oct.w + VB_THREADS - 1) / VB_THREADS, (oct.h + NOUT - 1) / NOUT);
  if (upsample) vblur_kernel<1><<<grid, VB_THREADS, 0, st>>>(d_weights, A);
  else vblur_kernel<0><<<grid, VB_THREADS, 0, st>>>(d_weights, A);
}

// ------------------------------------------------ step-function helpers (fp64) --
// SIFT_blurMatrix2DChunk on Matrix2D numbers: fp64 in, fp64 out, any radius.
__global__ void blur_h_f64_kernel(const double *__restrict__ in, int cols, double *__restrict__ tmp,
                                  const double *__restrict__ wgt, int R, int x1, int x2, int ya, int yb)
{
  const int x = x1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = ya + blockIdx.y;
  if (x >= x2 || y >= yb) return;
  const double *row = in + (size_t)y * cols;
  double acc = 0.0;
  for (int i = 0; i <= 2 * R; i++) {
    int xx = min(max(x + i - R, 0), cols - 1);
    acc = fma(row[xx], wgt[i], acc);
  }
  tmp[(size_t)y * cols + x] = acc;
}

__global__ void blur_v_f64_kernel(const double *__restrict__ tmp, int rows, int cols, double *__restrict__ out,
                                  const double *__restrict__ wgt, int R, int x1, int y1, int x2, int y2)
{
  const int x = x1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y1 + blockIdx.y;
  if (x >= x2 || y >= y2) return;
  double acc = 0.0;
  for (int j = 0; j <= 2 * R; j++) {
    int yy = min(max(y + j - R, 0), rows - 1);
    acc = fma(tmp[(size_t)yy * cols + x], wgt[j], acc);
  }
  out[(size_t)y * cols + x] = acc;
}

void launch_blur_plane_f64(cudaStream_t st, const double *d_in, int rows, int cols, double *d_tmp, double *d_out,
                           const double *d_w, int radius, int x1, int y1, int x2, int y2)
{
  const int ya = max(0, y1 - radius), yb = min(rows, y2 + radius);
  dim3 block(128);
  dim3 gh((x2 - x1 + 127) / 128, yb - ya);
  blur_h_f64_kernel<<<gh, block, 0, st>>>(d_in, cols, d_tmp, d_w, radius, x1, x2, ya, yb);
  dim3 gv((x2 - x1 + 127) / 128, y2 - y1);
  blur_v_f64_kernel<<<gv, block, 0, st>>>(d_tmp, rows, cols, d_out, d_w, radius, x1, y1, x2, y2);
}

__global__ void subtract_f64_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                    double *__restrict__ out, int cols, int x1, int y1, int x2, int y2)
{
  const int x = x1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y1 + blockIdx.y;
  if (x >= x2 || y >= y2) return;
  const size_t i = (size_t)y * cols + x;
  out[i] = a[i] - b[i];                                                         // sift.js:172
}

void launch_subtract_f64(cudaStream_t st, const double *a, const double *b, double *out, int cols,
                         int x1, int y1, int x2, int y2)
{
  dim3 g((x2 - x1 + 127) / 128, y2 - y1);
  subtract_f64_kernel<<<g, 128, 0, st>>>(a, b, out, cols, x1, y1, x2, y2);
}

// Matrix2D_linearResize (matrix2d.js:112-138): out[a][b] = in[floor(a*rate)][floor(b*rate)].
// The reference advances a double counter by `rate`; for the rates it uses (0.5, 2.0) and any
// dyadic rate a*rate is exact, so the product reproduces the accumulated counter.
__global__ void resize_f64_kernel(const double *__restrict__ in, int cols, double rate, double *__restrict__ out,
                                  int orows, int ocols)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  if (b >= ocols || a >= orows) return;
  out[(size_t)a * ocols + b] = in[(size_t)((int)floor(a * rate)) * cols + (int)floor(b * rate)];
}

void launch_resize_f64(cudaStream_t st, const double *in, int rows, int cols, double rate, double *out,
                       int orows, int ocols)
{
  (void)rows;
  dim3 g((ocols + 127) / 128, orows);
  resize_f64_kernel<<<g, 128, 0, st>>>(in, cols, rate, out, orows, ocols);
}

__global__ void seed_to_f32_kernel(const double *__restrict__ seed, int w, int h, float *__restrict__ dst, int pitch)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x < w && y < h) dst[(size_t)y * pitch + x] = (float)seed[(size_t)y * w + x];
}

void launch_seed_to_f32(cudaStream_t st, const double *seed, int w, int h, float *dst, int pitch)
{
  dim3 g((w + 127) / 128, h);
  seed_to_f32_kernel<<<g, 128, 0, st>>>(seed, w, h, dst, pitch);
}
