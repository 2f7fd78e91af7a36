// blur_oct0p.cu -- octave 0 as TWO polyphase passes over row bands of the image, the intermediate held in L2.
//
// Reference path restated (as blur_fused.cu): Matrix2D_linearResize(input, 0.5) (background.js:84,
// matrix2d.js:112-138), SIFT_blurMatrix2DChunk of that pixel-doubled base for every level (background.js:145-210,
// sift.js:72-149, clamp-to-edge per axis sift.js:116-119), SIFT_subtractMatrix2DChunk between neighbouring levels
// (sift.js:154-188) and the rate-2.0 resize of level `spo` that seeds the next octave (background.js:114-130).
// Polyphase: the base is a pixel-doubled image, so the 2R+1 taps collapse onto R+1 merged taps per output phase
// over SOURCE samples, in both directions (blur_fused.cu has the algebra).
//
// Why two passes: the fused tile kernels (blur_fused.cu, blur_oct0.cu, blur_oct0s.cu) keep the fp64 pipe 30-38 %
// busy -- every level costs a CTA barrier or two, all warps of a CTA convert and store at the same time, and the
// tile halo is filtered twice (422 instead of 360 DFMA per input pixel) -- while the two-pass kernels of octaves
// >= 1 (blur_sep.cu), whose tap loops run level after level without a CTA-wide phase change, reach 59 % / 42 %.
// What kept octave 0 out of that scheme is the intermediate: 6 levels x 2W x H doubles (199 MB at 1080p) would go
// through DRAM twice.  So the image is processed in BANDS of source rows: pass A' writes the band's intermediate
// (tens of MB, an L2-persisting window), pass B' reads it back from L2.
//   pass A' (x direction, source rows -> T^T):  in = source [y][x];  T_s^T[X = 2x + phase][y], fp64.  A lane owns
//            a source row, a thread 8 source positions x 2 phases, a CTA 32 rows x 64 positions.  No halo rows are
//            recomputed in x; a band carries 8 extra rows above and below (the y window of pass B').
//   pass B' (y direction, T^T -> levels):  in = T_s^T[X][y] (contiguous along y);  G_s / D_{s-1}[Y = 2y + phase][X].
//            A lane owns a column X, a thread 8 source rows x 2 phases; the box of each level arrives by TMA
//            (cp.async.bulk.tensor, mbarrier completion, one level ahead); the previous level's unrounded values
//            stay in registers for the DoG; lanes run along X, so every store instruction writes one 128-byte row
//            segment.
// "Filter along the contiguous axis, write transposed": both sides of both passes are coalesced.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "common.cuh"

#define OP_LINES 32                    // lines per CTA = lanes
#define OP_NW 8                        // warps per CTA
#define OP_POS (8 * OP_NW)             // 64 source positions along the filtered axis per CTA
#define OP_HALO 8                      // supports R <= 16
#define OP_PITCH (OP_LINES + 1)        // pass A' tile: [position][33]
#define OP_A_SPAN (OP_POS + 2 * OP_HALO + 1)   // staged positions (+ 1 for the window prefetch)
#define OP_B_SPAN (OP_POS + 2 * OP_HALO + 2)   // 82: box width of pass B' (= 2 mod 4: lanes spread over the bank pairs)
#define OP_MAXR 16
#define OP_WSTRIDE 24                  // merged tap table of blur_fused.cu: [nlev][24]{w0, w1}

struct Oct0pArgs {
  // source image
  const void *src;
  size_t src_pitch;
  int src_w, src_h, dtype;
  // band: T^T sample 0 = source row ya; pass A' fills rows [ya, yb); pass B' emits source rows [y0, y1)
  int ya, yb, y0, y1;
  double *T;                      // level s at T + s * t_plane; line X at + X * t_pitch
  size_t t_pitch, t_plane;
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, nlev;
  int radius[SIFT_MAX_LEVELS];
  int woff;
  const CUtensorMap *maps;        // [nlev] over the T^T planes: box OP_B_SPAN samples x 32 lines
};

__device__ __forceinline__ unsigned op_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double op_u8_over_255(unsigned v)      // image-utils.js:114, exactly (see blur_fused.cu)
{
  const double r = 1.0 / 255.0;
  const double x = (double)v;
  const double q = x * r;
  return fma(fma(-q, 255.0, x), r, q);
}

// Two-phase sliding window over 8 neighbouring positions (as blur_oct0.cu): a0[k] = sum_j w0[j] v[k+j],
// a1[k] = sum_j w1[j] v[k+j], v[p] = base[p * STRIDE]; first tap initialises; positions up to np + 7 are read.
#define OP_STEP(U, JJ)                                                                      \
  {                                                                                         \
    const double2 c = w2[(JJ)];                                                             \
    _Pragma("unroll") for (int k = 0; k < 8; k++) {                                         \
      a0[k] = fma(c.x, vw[(k + (U)) & 7], a0[k]);                                           \
      a1[k] = fma(c.y, vw[(k + (U)) & 7], a1[k]);                                           \
    }                                                                                       \
    vw[(U) & 7] = nxt[(JJ) * STRIDE];                                                       \
  }
template <int STRIDE>
__device__ __forceinline__ void op_window(const double *__restrict__ base, const double2 *__restrict__ w2, const int np,
                                          double (&a0)[8], double (&a1)[8])
{
  double vw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) vw[k] = base[k * STRIDE];
  const double *nxt = base + 8 * STRIDE;
  {
    const double2 c = w2[0];
#pragma unroll
    for (int k = 0; k < 8; k++) { a0[k] = c.x * vw[k]; a1[k] = c.y * vw[k]; }
    vw[0] = nxt[0];
  }
  int j = 1;
  for (; j + 8 <= np; j += 8) {
    OP_STEP(1, j) OP_STEP(2, j + 1) OP_STEP(3, j + 2) OP_STEP(4, j + 3)
    OP_STEP(5, j + 4) OP_STEP(6, j + 5) OP_STEP(7, j + 6) OP_STEP(0, j + 7)
  }
  const int rem = np - j;
  if (rem & 4) { OP_STEP(1, j) OP_STEP(2, j + 1) OP_STEP(3, j + 2) OP_STEP(4, j + 3) }
  if (rem & 2) {
    if (rem & 4) { OP_STEP(5, j + 4) OP_STEP(6, j + 5) }
    else { OP_STEP(1, j) OP_STEP(2, j + 1) }
  }
  if (rem & 1) {
    switch (rem & 6) {
      case 0: OP_STEP(1, j) break;
      case 2: OP_STEP(3, j + 2) break;
      case 4: OP_STEP(5, j + 4) break;
      default: OP_STEP(7, j + 6) break;
    }
  }
}

// ---------------------------------------------------------------------------------------------- pass A'
__global__ void __launch_bounds__(32 * OP_NW, 2)
oct0p_pass_a_kernel(const double *__restrict__ weights, const __grid_constant__ Oct0pArgs A)
{
  extern __shared__ __align__(128) double smem[];
  double *tile = smem;                                 // [OP_A_SPAN][33]: tile[e][line] = source(row a0 + line, col b_tile - 8 + e)
  double *Wt = smem + OP_A_SPAN * OP_PITCH + 1;        // (+1: keeps the 16-byte alignment of the tap pairs: 81 * 33 is odd)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a0 = A.ya + blockIdx.y * OP_LINES;         // first source row of the CTA
  const int b_tile = blockIdx.x * OP_POS;              // first source column

  for (int e = threadIdx.x; e < A.nlev * 2 * OP_WSTRIDE; e += 32 * OP_NW) Wt[e] = __ldg(weights + A.woff + e);
  // stage: warp w brings lines w, w + 8, ...; lanes run along the source row (coalesced), stores are transposed
#pragma unroll 1
  for (int l = warp; l < OP_LINES; l += OP_NW) {
    const int gy = min(a0 + l, A.src_h - 1);           // rows past the band / image replicate the last one (never stored)
    const char *row = (const char *)A.src + (size_t)gy * A.src_pitch;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int e = lane + 32 * i;
      if (e < OP_A_SPAN) {
        const int gx = min(max(b_tile - OP_HALO + e, 0), A.src_w - 1);          // clamp-to-edge, sift.js:116-119
        double v;
        switch (A.dtype) {
          case SIFT_U8: v = op_u8_over_255(__ldg((const unsigned char *)row + gx)); break;
          case SIFT_F32: v = (double)__ldg((const float *)row + gx); break;
          case SIFT_F64: v = __ldg((const double *)row + gx); break;
          default: {
            const uchar4 c = __ldg((const uchar4 *)row + gx);
            const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)c.x, 0.299), __dmul_rn((double)c.y, 0.587)),
                                       __dmul_rn((double)c.z, 0.114));                            // image-utils.js:107
            v = g / 255.0;
          }
        }
        tile[e * OP_PITCH + l] = v;
      }
    }
  }
  __syncthreads();

  const int a = a0 + lane;                             // this lane's source row
  const int b0 = b_tile + 8 * warp;                    // first of this thread's 8 source positions
  if (b0 >= A.src_w) return;                           // warp-uniform
  const bool line_ok = a < A.yb;
  const size_t t_line = (size_t)(a - A.ya);
  for (int s = 0; s < A.nlev; s++) {                   // octave 0 blurs every level from the base (background.js:110)
    const int R = A.radius[s];
    const int clo = -((R + 1) / 2);
    const int np = R + 1 + (R & 1);
    double a0v[8], a1v[8];
    op_window<OP_PITCH>(tile + (8 * warp + OP_HALO + clo) * OP_PITCH + lane, reinterpret_cast<const double2 *>(Wt + s * 2 * OP_WSTRIDE),
                        np, a0v, a1v);
    if (line_ok) {
      double *out = A.T + (size_t)s * A.t_plane + (size_t)(2 * b0) * A.t_pitch + t_line;
      const size_t tp = A.t_pitch;
      if (b0 + 8 <= A.src_w) {
#pragma unroll
        for (int k = 0; k < 8; k++) { *out = a0v[k]; out += tp; *out = a1v[k]; out += tp; }
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (b0 + k < A.src_w) { out[(size_t)(2 * k) * tp] = a0v[k]; out[(size_t)(2 * k + 1) * tp] = a1v[k]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- pass B'
// Epilogue of one level for one thread: rows Y = 2 (r + k) + phase, column X; c*: this level, p*: the previous one.
__device__ __forceinline__ void op_emit(const Oct0pArgs &A, int s, int X, int r, int rows_ok, bool seed_lane,
                                        const double (&c0)[8], const double (&c1)[8], const double (&p0)[8], const double (&p1)[8])
{
  const size_t rowp = (size_t)A.oct.pitch;
  float *gp = A.oct.gauss[s] + (size_t)(2 * r) * rowp + X;
  float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + (size_t)(2 * r) * rowp + X;
  const bool wg = A.keep_gauss != 0, wd = s > 0;
  if (rows_ok == 8) {
    if (wg) {
#pragma unroll
      for (int k = 0; k < 8; k++) { *gp = (float)c0[k]; gp += rowp; *gp = (float)c1[k]; gp += rowp; }
    }
    if (wd) {
#pragma unroll
      for (int k = 0; k < 8; k++) { *dp = (float)(p0[k] - c0[k]); dp += rowp; *dp = (float)(p1[k] - c1[k]); dp += rowp; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < rows_ok) {
        if (wg) { gp[(size_t)(2 * k) * rowp] = (float)c0[k]; gp[(size_t)(2 * k + 1) * rowp] = (float)c1[k]; }
        if (wd) { dp[(size_t)(2 * k) * rowp] = (float)(p0[k] - c0[k]); dp[(size_t)(2 * k + 1) * rowp] = (float)(p1[k] - c1[k]); }
      }
    }
  }
  if (s == A.spo && seed_lane) {                       // in[2a][2b] (matrix2d.js:129): even rows (phase 0), even columns
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int nr = r + k + A.oct.seed_off;           // row of the next octave (strip-local)
      if (k < rows_ok && nr >= 0 && nr < A.next.h) {
        A.next.seed64[(size_t)nr * A.next.w + (X >> 1)] = c0[k];
        A.next.gauss[0][(size_t)nr * A.next.pitch + (X >> 1)] = (float)c0[k];
      }
    }
  }
}

__global__ void __launch_bounds__(32 * OP_NW, 2)
oct0p_pass_b_kernel(const double *__restrict__ weights, const __grid_constant__ Oct0pArgs A)
{
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  constexpr int BUF = (OP_LINES * OP_B_SPAN + 15) & ~15;   // doubles per buffer, 128-byte multiple
  double *buf0 = smem, *buf1 = smem + BUF;
  double *Wt = smem + 2 * BUF;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X0 = blockIdx.y * OP_LINES;                // first column (line of T^T)
  const int r_tile = A.y0 + blockIdx.x * OP_POS;       // first source row of the CTA
  const int start = r_tile - OP_HALO - A.ya;           // first sample of the box in T^T coordinates (even; < 0 only at the image top)

  auto issue = [&](int s) {                            // one thread: arm the barrier, one box of level s
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was read / patched through the generic proxy
    const unsigned b = op_smem_u32(&bar[s & 1]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(OP_LINES * OP_B_SPAN * 8)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(op_smem_u32((s & 1) ? buf1 : buf0)), "l"(A.maps + s), "r"(b), "r"(start), "r"(X0)
        : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(op_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(op_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = threadIdx.x; e < A.nlev * 2 * OP_WSTRIDE; e += 32 * OP_NW) Wt[e] = __ldg(weights + A.woff + e);
  __syncthreads();
  if (threadIdx.x == 0) issue(0);

  const int X = X0 + lane;
  const int r = r_tile + 8 * warp;                     // first of this thread's 8 source rows
  const bool active = r < A.y1;                        // warp-uniform
  const bool line_ok = X < A.oct.w;
  const int rows_ok = line_ok ? min(8, A.y1 - r) : 0;  // source rows of this thread inside the band (and the image)
  const bool seed_lane = A.has_next && (X & 1) == 0 && line_ok;
  // clamp-to-edge along y (sift.js:118-119): the samples of T^T that exist are rows [ya, yb); what the box holds
  // beyond them is zero fill (above the image) or stale (below the band) and is replaced by the edge sample.  Only the
  // first / last band can reach them: interior bands carry 8 computed halo rows on each side.
  const int first_ok = max(0, -start);                 // box index of source row max(ya, ...) = first valid sample
  const int last_ok = min(OP_B_SPAN - 1, (A.yb - 1 - A.ya) - start);
  const bool patch = first_ok > 0 || last_ok < OP_POS + 2 * OP_HALO - 1;      // CTA-uniform

  double a0[8], a1[8], b0[8], b1[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { b0[k] = 0.0; b1[k] = 0.0; }
#pragma unroll 1
  for (int s = 0; s < A.nlev; s++) {
    if (threadIdx.x == 0 && s + 1 < A.nlev) issue(s + 1);      // the other buffer was released by the barrier that ended level s-1
    {
      const unsigned parity = (s >> 1) & 1;
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(op_smem_u32(&bar[s & 1])), "r"(parity) : "memory");
      }
    }
    double *cur = (s & 1) ? buf1 : buf0;
    if (patch) {
      double *ln = cur + lane * OP_B_SPAN;
      if (first_ok > 0) {
        const double v = ln[first_ok];
        for (int e = warp; e < first_ok; e += OP_NW) ln[e] = v;
      }
      if (last_ok < OP_B_SPAN - 1) {
        const double v = ln[last_ok];
        for (int e = last_ok + 1 + warp; e < OP_B_SPAN; e += OP_NW) ln[e] = v;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    if (active) {
      const int R = A.radius[s];
      const int clo = -((R + 1) / 2);
      const int np = R + 1 + (R & 1);
      const double *base = cur + lane * OP_B_SPAN + 8 * warp + OP_HALO + clo;
      const double2 *w2 = reinterpret_cast<const double2 *>(Wt + s * 2 * OP_WSTRIDE);
      // the two accumulator sets swap roles every level (this level's values are the next level's "previous")
      if (s & 1) { op_window<1>(base, w2, np, b0, b1); op_emit(A, s, X, r, rows_ok, seed_lane, b0, b1, a0, a1); }
      else { op_window<1>(base, w2, np, a0, a1); op_emit(A, s, X, r, rows_ok, seed_lane, a0, a1, b0, b1); }
    }
    __syncthreads();                                   // this buffer may be refilled (level s + 2)
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*Oct0pEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool oct0p_supported(const LevelPlan *plans, int nlev)
{
  // Off by default: 0.259 ms per 1080p frame against 0.146 for the fused tile kernel.  The bands that keep the
  // intermediate in L2 (192 source rows at 1080p) give grids of 210 / 360 CTAs -- under and just over one wave of
  // 296 -- twelve launches per frame, and the short polyphase windows (7-17 taps) leave pass A' at 25 % and pass B' at
  // 29 % of the fp64 pipe (profiles/r02_oct0_experiments.md).  SIFT_B200_OCT0_BANDS=1 selects it.
  static const bool on = getenv("SIFT_B200_OCT0_BANDS") != nullptr;
  if (!on) return false;
  for (int s = 0; s < nlev; s++)
    if (plans[s].radius < 1 || plans[s].radius > OP_MAXR) return false;
  return true;
}

// Source rows per band: the band's intermediate (nlev planes of 2W lines x (rows + 16) doubles) should sit in L2
// beside the band's output stream; bands are multiples of 64 rows (the pass B' tile).
int oct0p_band_rows(int src_w, int src_h, int nlev)
{
  static const char *env = getenv("SIFT_B200_OCT0_BAND");
  int rows = env && atoi(env) > 0 ? atoi(env) : 0;
  if (!rows) {
    const double budget = 48e6;                                        // bytes of intermediate per band
    rows = (int)(budget / ((double)nlev * 2.0 * src_w * 8.0)) - 16;
    rows = rows < 64 ? 64 : (rows / 64) * 64;
  }
  rows = (rows + 63) / 64 * 64;
  const int need = (src_h + 63) / 64 * 64;
  return rows < need ? rows : need;
}

static size_t oct0p_pitch(int band_rows) { return (size_t)((band_rows + 16 + 15) & ~15); }

size_t oct0p_t_bytes(int src_w, int src_h, int nlev)
{
  return (size_t)nlev * 2 * src_w * oct0p_pitch(oct0p_band_rows(src_w, src_h, nlev)) * sizeof(double);
}

size_t oct0p_map_bytes(int nlev) { return (size_t)nlev * sizeof(CUtensorMap); }

bool oct0p_build_maps(double *tbase, int src_w, int src_h, int nlev, void *h_maps)
{
  static Oct0pEncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return false;
    encode = (Oct0pEncodeTiledFn)fn;
  }
  const size_t pitch = oct0p_pitch(oct0p_band_rows(src_w, src_h, nlev));
  CUtensorMap *maps = (CUtensorMap *)h_maps;
  for (int s = 0; s < nlev; s++) {
    const cuuint64_t gdim[2] = { (cuuint64_t)pitch, (cuuint64_t)(2 * src_w) };
    const cuuint64_t gstride[1] = { (cuuint64_t)pitch * sizeof(double) };
    const cuuint32_t box[2] = { OP_B_SPAN, OP_LINES };
    const cuuint32_t estride[2] = { 1, 1 };
    if (encode(&maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)(tbase + (size_t)s * 2 * src_w * pitch), gdim, gstride, box,
               estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

// Returns the number of kernel launches.
int launch_oct0p(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                 const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                 int poly_woff, int nlev, int spo, int keep_gauss, double *tbase, const void *d_maps)
{
  Oct0pArgs A;
  memset(&A, 0, sizeof A);
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < SIFT_MAX_LEVELS; s++) A.radius[s] = s < nlev ? plans[s].radius : 0;
  A.woff = poly_woff;
  const int band = oct0p_band_rows(src_w, src_h, nlev);
  A.t_pitch = oct0p_pitch(band);
  A.t_plane = (size_t)2 * src_w * A.t_pitch;
  A.T = tbase;
  A.maps = (const CUtensorMap *)d_maps;
  const size_t smem_a = (size_t)(OP_A_SPAN * OP_PITCH + 1 + nlev * 2 * OP_WSTRIDE) * sizeof(double);
  const size_t smem_b = (size_t)(2 * ((OP_LINES * OP_B_SPAN + 15) & ~15) + nlev * 2 * OP_WSTRIDE) * sizeof(double);
  cudaFuncSetAttribute(oct0p_pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
  cudaFuncSetAttribute(oct0p_pass_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
  int launches = 0;
  for (int y0 = 0; y0 < src_h; y0 += band) {
    A.y0 = y0;
    A.y1 = y0 + band < src_h ? y0 + band : src_h;
    A.ya = y0 - OP_HALO > 0 ? y0 - OP_HALO : 0;
    A.yb = A.y1 + OP_HALO < src_h ? A.y1 + OP_HALO : src_h;
    dim3 ga((src_w + OP_POS - 1) / OP_POS, (A.yb - A.ya + OP_LINES - 1) / OP_LINES);
    oct0p_pass_a_kernel<<<ga, 32 * OP_NW, smem_a, st>>>(d_weights, A);
    static const bool dbg = getenv("SIFT_B200_OCT0P_DEBUG") != nullptr;
    if (dbg) { cudaError_t e = cudaStreamSynchronize(st); fprintf(stderr, "[oct0p] pass A band %d: %s\n", y0, cudaGetErrorString(e)); }
    dim3 gb((A.y1 - A.y0 + OP_POS - 1) / OP_POS, (2 * src_w + OP_LINES - 1) / OP_LINES);
    oct0p_pass_b_kernel<<<gb, 32 * OP_NW, smem_b, st>>>(d_weights, A);
    if (dbg) { cudaError_t e = cudaStreamSynchronize(st); fprintf(stderr, "[oct0p] pass B band %d: %s\n", y0, cudaGetErrorString(e)); }
    launches += 2;
  }
  return launches;
}
