// blur_generic.cu -- radius-generic separable Gaussian scale-space kernels.
//
// Restates (as a separable fp64 computation) the reference's dense clamped 2D
// correlation SIFT_blurMatrix2DChunk (src/sift.js:72-149) applied level by level
// to the octave base (background.js:103-224), the DoG subtraction
// SIFT_subtractMatrix2DChunk (src/sift.js:154-188, finer minus coarser) and the
// 2x decimation that seeds the next octave (matrix2d.js:112-138 with rate 2.0,
// background.js:114-130).  Used for octaves >= 1 (any radius, also wider than the
// image) and as the octave-0 path for radii the fused polyphase kernel
// (blur_fused.cu) does not cover; the 2x nearest-neighbour upsample of the input
// (background.js:84) is then folded into the loads.
//
// Two passes through an fp64 intermediate T (L2 resident for these octaves):
//   hblur: T_s[y][x]  = sum_i w_s[i] * base[y][clamp(x + i - R_s)]     one level per CTA (blockIdx.z)
//   vblur: G_s[y][x]  = sum_j w_s[j] * T_s[clamp(y + j - R_s)][x]      levels across threadIdx.z
//          D_{s-1}    = G_{s-1} - G_s  formed from the UNROUNDED fp64 accumulators (exchanged
//          through shared memory), both stored fp32; level `spo` is also decimated into
//          the next seed (fp64 + fp32).
// Accumulation is fp64 in both passes (fp32 accumulation fails the
// 1e-5 / 1e-3 px parity bars).  Every thread slides a register window of 8 samples
// over its 8 neighbouring outputs (static rotation, runtime tap count, exact: no
// padded taps), so each loaded sample feeds 8 DFMAs.
#include "common.cuh"

#define NOUT 8

struct BlurLevels {
  int nlev;                       // number of blurred levels handled
  int level[SIFT_MAX_LEVELS];     // pyramid level index of each
  int radius[SIFT_MAX_LEVELS];
  int woff[SIFT_MAX_LEVELS];
  double *T[SIFT_MAX_LEVELS];
};

// a[k] = sum_{j<n} w[j] * v(k + j), k < 8.  loadv(p) is called for p in [0, n+7] (p <= n+6 used).
template <typename LoadV>
__device__ __forceinline__ void window8(const double *__restrict__ w, const int n, LoadV loadv, double (&a)[NOUT])
{
  double vw[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) { vw[k] = loadv(k); a[k] = 0.0; }
  int j = 0;
  for (; j + NOUT <= n; j += NOUT) {
#pragma unroll
    for (int u = 0; u < NOUT; u++) {
      const double c = w[j + u];
#pragma unroll
      for (int k = 0; k < NOUT; k++) a[k] = fma(c, vw[(k + u) & 7], a[k]);
      vw[u] = loadv(j + u + NOUT);
    }
  }
#pragma unroll
  for (int u = 0; u < NOUT - 1; u++) {
    if (j + u < n) {
      const double c = w[j + u];
#pragma unroll
      for (int k = 0; k < NOUT; k++) a[k] = fma(c, vw[(k + u) & 7], a[k]);
      vw[u] = loadv(j + u + NOUT);
    }
  }
}

// ---- source pixel -> double, exactly the reference's float64 image value ------
__device__ __forceinline__ double src_at(const void *row, int x, int dtype)
{
  switch (dtype) {
    case SIFT_U8: return (double)((const unsigned char *)row)[x] / 255.0;        // image-utils.js:114
    case SIFT_F32: return (double)((const float *)row)[x];
    case SIFT_F64: return ((const double *)row)[x];
    default: {
      const uchar4 p = ((const uchar4 *)row)[x];
      // (R*0.299) + (G*0.587) + (B*0.114), then / 255.0 -- image-utils.js:107-114, unfused
      const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)p.x, 0.299), __dmul_rn((double)p.y, 0.587)),
                                 __dmul_rn((double)p.z, 0.114));
      return g / 255.0;
    }
  }
}

// ------------------------------------------------------------------- hblur ----
#define HB_WARPS 4                    // rows per CTA (one warp per row)
#define HB_TW (32 * NOUT)             // 256 outputs per row per CTA
#define HB_THREADS (32 * HB_WARPS)

__device__ __forceinline__ int padidx(int e) { return e + (e >> 3); }   // 9-double stride per 8: conflict-free LDS.64

// grid.z == 1: the CTA loops over all levels (rows staged once with the largest halo);
// grid.z == nlev: one level per CTA (small octaves: more CTAs in flight).
__global__ void __launch_bounds__(HB_THREADS)
hblur_kernel(const void *__restrict__ src, size_t src_pitch, int src_w, int dtype, int ups, int w, int hrows,
             int rmax, int wtotal, const double *__restrict__ weights, BlurLevels L)
{
  extern __shared__ double smem[];
  const bool all_levels = gridDim.z == 1;
  const int li_first = all_levels ? 0 : blockIdx.z;
  const int li_last = all_levels ? L.nlev : blockIdx.z + 1;
  const int RS = all_levels ? rmax : L.radius[li_first];   // staged halo
  const int span = HB_TW + 2 * RS + NOUT;           // samples staged per row (window prefetch reads n+7 past the first)
  const int rowstride = padidx(span) + 1;
  double *wsm = smem;                               // taps of the handled levels, back to back
  double *rows = smem + ((wtotal + 1) & ~1);
  const int x_tile = blockIdx.x * HB_TW;
  const int row0 = blockIdx.y * HB_WARPS;
  const int tid = threadIdx.y * 32 + threadIdx.x;

  {
    int wo = 0;
    for (int li = li_first; li < li_last; li++) {
      const int n = 2 * L.radius[li] + 1;
      for (int e = tid; e < n; e += HB_THREADS) wsm[wo + e] = __ldg(weights + L.woff[li] + e);
      wo += n;
    }
  }
  // flat staging loop, 4 independent loads in flight per thread
  const int nrow = min(HB_WARPS, hrows - row0);
  const int total = nrow * span;
  for (int f0 = tid; f0 < total; f0 += 4 * HB_THREADS) {
    double v[4];
    int so[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int f = f0 + i * HB_THREADS;
      const int r = f / span, e = f - r * span;
      so[i] = -1;
      v[i] = 0.0;
      if (f < total) {
        int col = x_tile - RS + e;
        col = min(max(col, 0), w - 1);                                           // sift.js:116-117 clamp
        const int sc = ups ? (col >> 1) : col;                                   // matrix2d.js:129 floor(j*0.5)
        v[i] = src_at((const char *)src + (size_t)(row0 + r) * src_pitch, min(sc, src_w - 1), dtype);
        so[i] = r * rowstride + padidx(e);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) if (so[i] >= 0) rows[so[i]] = v[i];
  }
  __syncthreads();

  const int y = row0 + threadIdx.y;
  const int x0 = x_tile + threadIdx.x * NOUT;
  if (y >= hrows || x0 >= w) return;
  const double *rowsm = rows + threadIdx.y * rowstride;
  int wo = 0;
  for (int li = li_first; li < li_last; li++) {
    const int R = L.radius[li], n = 2 * R + 1;
    const int e0 = threadIdx.x * NOUT + RS - R;      // sample feeding tap 0 of output 0
    double acc[NOUT];
    window8(wsm + wo, n, [&](int p) { return rowsm[padidx(e0 + p)]; }, acc);
    wo += n;
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

#define SIFT_BIG_OCTAVE_WARPS (148 * 16)   // enough warps to fill the chip with one thread per 8 outputs of all levels

void launch_hblur(cudaStream_t st, const void *src, int dtype, size_t src_pitch_bytes, int src_w, int src_h,
                  int upsample, int w, int hrows, const double *d_weights, const LevelPlan *plans, int first_level,
                  int nlev, double *const *T, double **)
{
  (void)src_h;
  int rmax;
  BlurLevels L = make_levels(plans, first_level, nlev, T, &rmax);
  dim3 block(32, HB_WARPS);
  dim3 grid((w + HB_TW - 1) / HB_TW, (hrows + HB_WARPS - 1) / HB_WARPS, 1);
  int wtotal = 0;
  for (int i = 0; i < L.nlev; i++) wtotal += 2 * L.radius[i] + 1;
  const int span = HB_TW + 2 * rmax + NOUT;
  const size_t smem = ((size_t)((wtotal + 2) & ~1) + (size_t)HB_WARPS * ((span + (span >> 3)) + 1)) * sizeof(double);
  // rows are staged once and every level is computed from them, unless that no longer fits shared memory
  const bool all_levels = smem <= 160 * 1024;
  if (!all_levels) { grid.z = L.nlev; }
  cudaFuncSetAttribute(hblur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  hblur_kernel<<<grid, block, smem, st>>>(src, src_pitch_bytes, src_w, dtype, upsample, w, hrows, rmax,
                                          all_levels ? wtotal : 2 * rmax + 1, d_weights, L);
}

// ------------------------------------------------------------------- vblur ----
#define VB_COLS 32
#define VB_RG 2                        // level-parallel variant: row groups of 8 per CTA -> 16 rows
#define VS_RG 8                        // level-sequential variant: 64 rows per CTA

struct VBlurArgs {
  OctaveDev oct;
  OctaveDev next;        // valid when has_next
  int has_next, spo, keep_gauss, seed_is_level0, ups;
  BlurLevels L;
};

// Write one thread's 8 outputs of level s: G_s, D_{s-1} = prev - G_s (sift.js:172), seed of the next octave.
__device__ __forceinline__ void vblur_store(const VBlurArgs &A, int s, int x, int y0, const double (&acc)[NOUT],
                                            const double (&prev)[NOUT])
{
  const int h = A.oct.h, pitch = A.oct.pitch;
#pragma unroll
  for (int k = 0; k < NOUT; k++) {
    const int y = y0 + k;
    if (y >= h) break;
    const size_t o = (size_t)y * pitch + x;
    if (A.keep_gauss) A.oct.gauss[s][o] = (float)acc[k];
    if (s > 0) A.oct.dog[s - 1][o] = (float)(prev[k] - acc[k]);
    const int nr = (y >> 1) + A.oct.seed_off;                                  // row of the next octave (strip-local)
    if (A.has_next && s == A.spo && ((y | x) & 1) == 0 && nr >= 0 && nr < A.next.h) {   // matrix2d.js:129 in[2a][2b]
      A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = acc[k];
      A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)acc[k];
    }
  }
}

__device__ __forceinline__ void vblur_window(const VBlurArgs &A, const double *__restrict__ wsm, int li, int x, int y0,
                                             double (&acc)[NOUT])
{
  const int w = A.oct.w, h = A.oct.h;
  const int R = A.L.radius[li];
  const double *__restrict__ T = A.L.T[li] + x;
  if (!A.ups && y0 - R >= 0 && y0 + R + 2 * NOUT <= h) {        // interior: no clamping, plain strides
    const double *base = T + (size_t)(y0 - R) * w;
    window8(wsm, 2 * R + 1, [&](int p) { return __ldg(base + (size_t)p * w); }, acc);
  } else {
    const int ups = A.ups;
    window8(wsm, 2 * R + 1, [&](int p) {
      int yy = min(max(y0 + p - R, 0), h - 1);                                 // sift.js:118-119 clamp
      if (ups) yy >>= 1;                                                       // rows 2b and 2b+1 are equal
      return __ldg(T + (size_t)yy * w);
    }, acc);
  }
}

// Level-sequential: one thread = one column x 8 rows x all levels, previous level kept in registers.
// The T rows a CTA needs for a level ((64 + 2R + 1) x 32 columns) are brought into shared memory with
// cp.async (LDGSTS) one level ahead of the arithmetic, so the sliding windows never wait on global memory.
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#define VS_THREADS (VB_COLS * VS_RG)

__device__ __forceinline__ void vseq_issue_tile(const VBlurArgs &A, int li, double *__restrict__ tile, int x0, int ybase,
                                                int tid)
{
  const int w = A.oct.w, h = A.oct.h, R = A.L.radius[li];
  const int rows = VS_RG * NOUT + 2 * R + 1;
  const double *__restrict__ T = A.L.T[li];
  const int c = tid & (VB_COLS - 1);
  if (x0 + c < w) {
    for (int i = tid / VB_COLS; i < rows; i += VS_THREADS / VB_COLS) {
      int yy = min(max(ybase - R + i, 0), h - 1);                              // sift.js:118-119 clamp
      if (A.ups) yy >>= 1;                                                     // rows 2b and 2b+1 are equal
      cp_async8(tile + i * VB_COLS + c, T + (size_t)yy * w + x0 + c);
    }
  }
  cp_async_commit();
}

__global__ void __launch_bounds__(VS_THREADS, 3)
vblur_seq_kernel(const double *__restrict__ weights, VBlurArgs A, int wtotal, int tile_elems)
{
  extern __shared__ double smem[];
  double *wsm = smem;
  double *tile0 = smem + ((wtotal + 1) & ~1);
  double *tile1 = tile0 + tile_elems;
  const int tid = threadIdx.y * VB_COLS + threadIdx.x;
  const int w = A.oct.w, h = A.oct.h;
  const int x0 = blockIdx.x * VB_COLS, ybase = blockIdx.y * VS_RG * NOUT;
  vseq_issue_tile(A, 0, tile0, x0, ybase, tid);
  {
    int wo = 0;
    for (int i = 0; i < A.L.nlev; i++) {
      const int n_i = 2 * A.L.radius[i] + 1;
      for (int e = tid; e < n_i; e += VS_THREADS) wsm[wo + e] = __ldg(weights + A.L.woff[i] + e);
      wo += n_i;
    }
  }
  const int x = x0 + threadIdx.x;
  const int y0 = ybase + threadIdx.y * NOUT;
  const bool active = x < w && y0 < h;
  double prev[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) prev[k] = 0.0;
  if (A.seed_is_level0 && active) {                     // octaves >= 1: level 0 is the unblurred seed
#pragma unroll
    for (int k = 0; k < NOUT; k++) prev[k] = A.oct.seed64[(size_t)min(y0 + k, h - 1) * w + x];
  }
  int wo = 0;
  for (int li = 0; li < A.L.nlev; li++) {
    double *cur = (li & 1) ? tile1 : tile0;
    if (li + 1 < A.L.nlev) {
      vseq_issue_tile(A, li + 1, (li & 1) ? tile0 : tile1, x0, ybase, tid);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                    // tile li (and, first time, the taps) visible to the CTA
    const int n = 2 * A.L.radius[li] + 1;
    if (active) {
      double acc[NOUT];
      const double *base = cur + (threadIdx.y * NOUT) * VB_COLS + threadIdx.x;
      window8(wsm + wo, n, [&](int p) { return base[p * VB_COLS]; }, acc);
      vblur_store(A, A.L.level[li], x, y0, acc, prev);
#pragma unroll
      for (int k = 0; k < NOUT; k++) prev[k] = acc[k];
    }
    wo += n;
    __syncthreads();                                    // cur is overwritten by the copy issued next iteration
  }
}

// Level-parallel (small octaves): levels across threadIdx.z, unrounded values exchanged through shared memory.
__global__ void __launch_bounds__(VB_COLS *VB_RG *SIFT_MAX_LEVELS)
vblur_par_kernel(const double *__restrict__ weights, VBlurArgs A)
{
  extern __shared__ double smem[];
  double *G64 = smem;                                     // [li][row][col]
  const int nl = A.L.nlev;
  double *wsm = smem + nl * (VB_RG * NOUT) * VB_COLS;
  const int li = threadIdx.z;
  const int tid = (threadIdx.z * VB_RG + threadIdx.y) * VB_COLS + threadIdx.x;
  const int nthreads = VB_COLS * VB_RG * nl;
  int wbase = 0;
  for (int i = 0; i < li; i++) wbase += 2 * A.L.radius[i] + 1;
  int wtotal = 0;
  for (int i = 0; i < nl; i++) {
    const int n_i = 2 * A.L.radius[i] + 1;
    for (int e = tid; e < n_i; e += nthreads) wsm[wtotal + e] = __ldg(weights + A.L.woff[i] + e);
    wtotal += n_i;
  }
  __syncthreads();

  const int w = A.oct.w, h = A.oct.h;
  const int x = blockIdx.x * VB_COLS + threadIdx.x;
  const int y0 = (blockIdx.y * VB_RG + threadIdx.y) * NOUT;
  const bool active = x < w && y0 < h;
  double acc[NOUT], prev[NOUT];
  if (active) {
    vblur_window(A, wsm + wbase, li, x, y0, acc);
    double *g = G64 + ((li * VB_RG + threadIdx.y) * NOUT) * VB_COLS + threadIdx.x;
#pragma unroll
    for (int k = 0; k < NOUT; k++) g[k * VB_COLS] = acc[k];
  }
  __syncthreads();
  if (!active) return;
#pragma unroll
  for (int k = 0; k < NOUT; k++) {
    if (li > 0) prev[k] = G64[(((li - 1) * VB_RG + threadIdx.y) * NOUT + k) * VB_COLS + threadIdx.x];
    else prev[k] = A.seed_is_level0 ? A.oct.seed64[(size_t)min(y0 + k, h - 1) * w + x] : 0.0;
  }
  vblur_store(A, A.L.level[li], x, y0, acc, prev);
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
  A.ups = upsample;
  int rmax;
  A.L = make_levels(plans, first_level, oct.nlev, T, &rmax);
  int wtotal = 0;
  for (int i = 0; i < A.L.nlev; i++) wtotal += 2 * A.L.radius[i] + 1;
  const int tile_elems = (VS_RG * NOUT + 2 * rmax + 1 + 1) * VB_COLS;      // +1 row: window prefetch slack
  const size_t smem_seq = ((size_t)((wtotal + 2) & ~1) + 2 * (size_t)tile_elems) * sizeof(double);
  if (smem_seq <= 200 * 1024) {
    dim3 block(VB_COLS, VS_RG);
    dim3 grid((oct.w + VB_COLS - 1) / VB_COLS, (oct.h + VS_RG * NOUT - 1) / (VS_RG * NOUT));
    const size_t smem = smem_seq;
    cudaFuncSetAttribute(vblur_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vblur_seq_kernel<<<grid, block, smem, st>>>(d_weights, A, wtotal, tile_elems);
  } else {
    dim3 block(VB_COLS, VB_RG, A.L.nlev);
    dim3 grid((oct.w + VB_COLS - 1) / VB_COLS, (oct.h + VB_RG * NOUT - 1) / (VB_RG * NOUT));
    const size_t smem = ((size_t)A.L.nlev * VB_RG * NOUT * VB_COLS + wtotal + 1) * sizeof(double);
    cudaFuncSetAttribute(vblur_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vblur_par_kernel<<<grid, block, smem, st>>>(d_weights, A);
  }
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
