// blur_fused.cu -- octave 0 in ONE kernel: 2x nearest-neighbour upsample, all Gaussian
// levels, DoG and the seed of octave 1, each output written once.
//
// Reference path restated: Matrix2D_linearResize(input, 0.5) (background.js:84,
// matrix2d.js:112-138), then for every level SIFT_blurMatrix2DChunk of that
// upsampled base (background.js:145-210, sift.js:72-149), SIFT_subtractMatrix2DChunk
// between neighbouring levels (sift.js:154-188) and the rate-2.0 resize of level
// `spo` that seeds the next octave (background.js:114-130).
//
// Polyphase form: the base is a pixel-doubled image, so for an
// output X = 2a+p the 2R+1 taps w[i] over u[X+i] = src[(X+i)>>1] collapse onto
// source samples src[a+c] with merged weights
//     phase 0: W0[c] = w[2c] + w[2c+1]      c in [floor(-R/2), floor(R/2)]
//     phase 1: W1[c] = w[2c-1] + w[2c]      c in [floor((1-R)/2), floor((1+R)/2)]
// (R+1 taps per phase, W1[c] = W0[-c]); clamping in the doubled domain equals
// clamping source coordinates.  The same holds vertically, so a level costs
// 6(R+1) fp64 FMAs per SOURCE pixel instead of 8(2R+1).
//
// One CTA = 32x32 source pixels = 64x64 outputs of every level.  Per level:
//   H pass  smem source tile (fp64, clamped halo of 8) -> Ts[row][X] (both phases)
//   V pass  Ts -> 16 outputs per thread (8 source rows x 2 phases), fp64
//   epilogue: G_s (fp32), D_{s-1} = G_{s-1} - G_s from the unrounded accumulators
//   (kept in registers across levels), seed of the next octave (fp64 + fp32).
// The tap loop is a runtime loop over a register-resident sliding window of 8
// samples (static rotation, no moves): every loaded sample feeds 16 DFMAs, the hot
// code is a few KB (a first version unrolled per radius stalled 32 % on instruction
// fetch), and any radius <= 16 runs the same code.
#include <cstdlib>
#include "common.cuh"

#define F0_SW 32
#define F0_SH 32
#define F0_HALO 8                               // supports R <= 16
#define F0_SROWS (F0_SH + 2 * F0_HALO)          // 48
#define F0_SCOLS (F0_SW + 2 * F0_HALO)          // 48
#define F0_SPITCH (F0_SCOLS + 1)                // 49: odd pitch -> lanes along rows are conflict-free
#define F0_TROWS (F0_SROWS + 1)                 // +1 slack row for the window prefetch
#define F0_TPITCH (2 * F0_SW + 1)               // 65
#define F0_THREADS 256
#define F0_MAXR 16
#define F0_WSTRIDE 24                           // padded taps per phase per level (n <= 18)

struct Fused0Args {
  const void *src;
  size_t src_pitch;
  int src_w, src_h, dtype;
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, nlev;
  int radius[SIFT_MAX_LEVELS];
  int woff;                       // offset of the staged tap table [nlev][2][F0_WSTRIDE] in the weight buffer
  const double *u8lut;            // 256 doubles v/255.0
};

// v / 255.0 exactly as the reference computes it (image-utils.js:114) without a divide: one Newton
// correction of v * (1/255) is the correctly rounded quotient for every v in 0..255
// (tests/test_u8_conversion.py checks all 256 values with exact rational arithmetic).
__device__ __forceinline__ double u8_over_255(unsigned char v)
{
  const double r = 1.0 / 255.0;
  const double x = (double)v;
  const double q = x * r;
  return fma(fma(-q, 255.0, x), r, q);
}

__device__ __forceinline__ double src0_at(const Fused0Args &A, const char *p)
{
  switch (A.dtype) {
    case SIFT_F32: return (double)*(const float *)p;
    case SIFT_F64: return *(const double *)p;
    default: {
      const uchar4 c = *(const uchar4 *)p;
      const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)c.x, 0.299), __dmul_rn((double)c.y, 0.587)),
                                 __dmul_rn((double)c.z, 0.114));                              // image-utils.js:107
      return g / 255.0;
    }
  }
}

// Two-phase sliding window over 8 neighbouring positions:
//   a0[k] = sum_{j<n} w0[j] v[k+j],  a1[k] = sum_{j<n} w1[j] v[k+j]
// v[p] = base[p * stride]; positions up to n+7 are read (n+6 used).  `w` holds the two phases interleaved
// ({w0[j], w1[j]} = one 16-byte load per tap).  The window rotates through 8 registers with compile-time
// indices: whole groups of 8 taps, then the remainder decomposed as 4 + 2 + 1 with one code block per
// (size, rotation) pair, so no tap is padded and the accumulators never change registers.
#define POLY_STEP(U, JJ)                                                                    \
  {                                                                                         \
    const double2 c = w2[(JJ)];                                                             \
    _Pragma("unroll") for (int k = 0; k < 8; k++) {                                         \
      a0[k] = fma(c.x, vw[(k + (U)) & 7], a0[k]);                                           \
      a1[k] = fma(c.y, vw[(k + (U)) & 7], a1[k]);                                           \
    }                                                                                       \
    vw[(U) & 7] = (double)nxt[(JJ) * stride];                                                       \
  }
template <typename TS>
__device__ __forceinline__ void poly_window(const TS *__restrict__ base, const int stride,
                                            const double *__restrict__ w, const int n, double (&a0)[8],
                                            double (&a1)[8])
{
  const double2 *__restrict__ w2 = reinterpret_cast<const double2 *>(w);
  double vw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) vw[k] = (double)base[k * stride];
  const TS *nxt = base + 8 * stride;
  {                                                          // the first tap initialises the accumulators (c * v = fma(c, v, 0))
    const double2 c = w2[0];
#pragma unroll
    for (int k = 0; k < 8; k++) { a0[k] = c.x * vw[k]; a1[k] = c.y * vw[k]; }
    vw[0] = (double)nxt[0];
  }
  int j = 1;
  for (; j + 8 <= n; j += 8) {
    POLY_STEP(1, j) POLY_STEP(2, j + 1) POLY_STEP(3, j + 2) POLY_STEP(4, j + 3)
    POLY_STEP(5, j + 4) POLY_STEP(6, j + 5) POLY_STEP(7, j + 6) POLY_STEP(0, j + 7)
  }
  const int rem = n - j;                                     // 0..7, uniform over the CTA
  if (rem & 4) { POLY_STEP(1, j) POLY_STEP(2, j + 1) POLY_STEP(3, j + 2) POLY_STEP(4, j + 3) }
  if (rem & 2) {
    if (rem & 4) { POLY_STEP(5, j + 4) POLY_STEP(6, j + 5) }
    else { POLY_STEP(1, j) POLY_STEP(2, j + 1) }
  }
  if (rem & 1) {
    switch (rem & 6) {
      case 0: POLY_STEP(1, j) break;
      case 2: POLY_STEP(3, j + 2) break;
      case 4: POLY_STEP(5, j + 4) break;
      default: POLY_STEP(7, j + 6) break;
    }
  }
}

// H pass of level s: rows needed by the level's vertical window, 8 source columns x 2 phases per item
__device__ __forceinline__ void fused0_hpass(const double *__restrict__ S, double *__restrict__ Ts,
                                             const double *__restrict__ Wt, int s, int R, int tid)
{
  const int clo0 = -((R + 1) / 2);
  const int n = R + 1 + (R & 1);                     // taps per phase incl. the phase-1 shift for odd R
  const double *w = Wt + s * 2 * F0_WSTRIDE;
  const int nrows = F0_SH + n - 1;
  const int row_first = F0_HALO + clo0;
  for (int i = tid; i < nrows * (F0_SW / 8); i += F0_THREADS) {
    const int g = (i >= nrows) + (i >= 2 * nrows) + (i >= 3 * nrows), rr = row_first + (i - g * nrows);
    double a0[8], a1[8];
    poly_window(S + rr * F0_SPITCH + F0_HALO + 8 * g + clo0, 1, w, n, a0, a1);
    double *t = Ts + rr * F0_TPITCH + 16 * g;
#pragma unroll
    for (int k = 0; k < 8; k++) { t[2 * k] = a0[k]; t[2 * k + 1] = a1[k]; }
  }
}

__global__ void __launch_bounds__(F0_THREADS, 2)
fused_octave0_kernel(const double *__restrict__ weights, const Fused0Args A)
{
  extern __shared__ double smem[];
  double *S = smem;                                   // [48][49] source tile, fp64
  double *Ts0 = S + F0_SROWS * F0_SPITCH;             // 2 x [49][65] horizontally blurred rows, both phases:
  double *Ts1 = Ts0 + F0_TROWS * F0_TPITCH;           // level s+1 is produced while level s is consumed
  double *Wt = Ts1 + F0_TROWS * F0_TPITCH;            // [nlev][F0_WSTRIDE]{w0, w1} zero-padded merged taps
  const int tid = threadIdx.x;
  const int a_tile = blockIdx.x * F0_SW, b_tile = blockIdx.y * F0_SH;

  for (int e = tid; e < A.nlev * 2 * F0_WSTRIDE; e += F0_THREADS) Wt[e] = __ldg(weights + A.woff + e);
  // source tile: 48*48 = 9 samples per thread, all loads in flight before the first use
  {
    static_assert(F0_SROWS * F0_SCOLS == 9 * F0_THREADS, "tile load assumes 9 samples per thread");
    int so[9];
    size_t go[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int e = tid + i * F0_THREADS;
      const int rr = e / F0_SCOLS, cc = e - rr * F0_SCOLS;
      const int gy = min(max(b_tile - F0_HALO + rr, 0), A.src_h - 1);        // clamp-to-edge, sift.js:116-119
      const int gx = min(max(a_tile - F0_HALO + cc, 0), A.src_w - 1);
      so[i] = rr * F0_SPITCH + cc;
      go[i] = (size_t)gy * A.src_pitch;
      go[i] += (size_t)gx * (A.dtype == SIFT_U8 ? 1 : (A.dtype == SIFT_F64 ? 8 : 4));
    }
    if (A.dtype == SIFT_U8) {
      unsigned char raw[9];
#pragma unroll
      for (int i = 0; i < 9; i++) raw[i] = __ldg((const unsigned char *)A.src + go[i]);
#pragma unroll
      for (int i = 0; i < 9; i++) S[so[i]] = u8_over_255(raw[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 9; i++) S[so[i]] = src0_at(A, (const char *)A.src + go[i]);
    }
  }
  for (int e = tid; e < F0_TPITCH; e += F0_THREADS) {                                            // slack rows
    Ts0[(F0_TROWS - 1) * F0_TPITCH + e] = 0.0;
    Ts1[(F0_TROWS - 1) * F0_TPITCH + e] = 0.0;
  }
  if (tid < F0_SROWS) S[tid * F0_SPITCH + F0_SCOLS] = 0.0;                                       // pad column
  __syncthreads();
  fused0_hpass(S, Ts0, Wt, 0, A.radius[0], tid);
  __syncthreads();

  // V-pass ownership: output column X of the tile, source rows 8*rg .. 8*rg+7, both row phases
  const int X = tid & (2 * F0_SW - 1), rg = tid / (2 * F0_SW);
  const int x = 2 * a_tile + X;
  const int w = A.oct.w, h = A.oct.h;
  const int y_first = 2 * (b_tile + 8 * rg);
  const size_t o_first = (size_t)y_first * A.oct.pitch + x;
  const bool col_ok = x < w;
  const bool rows_full = y_first + 16 <= h;
  const bool seed_lane = A.has_next && (X & 1) == 0 && col_ok;

  double prev[16];
#pragma unroll
  for (int i = 0; i < 16; i++) prev[i] = 0.0;

  for (int s = 0; s < A.nlev; s++) {                 // octave 0 blurs every level from the base (background.js:110)
    const int R = A.radius[s];
    const int clo0 = -((R + 1) / 2);
    const int n = R + 1 + (R & 1);
    const double *Ts = (s & 1) ? Ts1 : Ts0;

    // ---- V pass of level s
    double a0[8], a1[8];
    poly_window(Ts + (F0_HALO + 8 * rg + clo0) * F0_TPITCH + X, F0_TPITCH, Wt + s * 2 * F0_WSTRIDE, n, a0, a1);

    // ---- epilogue: G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172), seed of the next octave
    if (col_ok) {
      // rows y_first + 2k (phase 0) and y_first + 2k + 1 (phase 1): one running pointer per plane, bumped one
      // row per store (two 32-bit adds), instead of a 64-bit multiply-add per store
      float *gp = A.oct.gauss[s] + o_first;
      float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + o_first;
      const size_t rowp = (size_t)A.oct.pitch;
      const bool wg = A.keep_gauss != 0, wd = s > 0;
      if (rows_full) {
        if (wg) {
#pragma unroll
          for (int k = 0; k < 8; k++) { *gp = (float)a0[k]; gp += rowp; *gp = (float)a1[k]; gp += rowp; }
        }
        if (wd) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            *dp = (float)(prev[2 * k] - a0[k]); dp += rowp;
            *dp = (float)(prev[2 * k + 1] - a1[k]); dp += rowp;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            if (y_first + 2 * k + q < h) {
              const double v = q ? a1[k] : a0[k];
              if (wg) gp[(size_t)(2 * k + q) * rowp] = (float)v;
              if (wd) dp[(size_t)(2 * k + q) * rowp] = (float)(prev[2 * k + q] - v);
            }
          }
        }
      }
      if (s == A.spo && seed_lane) {                 // in[2a][2b] (matrix2d.js:129): even rows (phase 0), even columns
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int nr = b_tile + 8 * rg + k + A.oct.seed_off;      // row of the next octave (strip-local)
          if (y_first + 2 * k < h && nr >= 0 && nr < A.next.h) {
            A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = a0[k];
            A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)a0[k];
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { prev[2 * k] = a0[k]; prev[2 * k + 1] = a1[k]; }

    // ---- H pass of level s+1 into the other buffer (nobody reads it: its V pass ended before the last barrier)
    if (s + 1 < A.nlev) fused0_hpass(S, (s & 1) ? Ts0 : Ts1, Wt, s + 1, A.radius[s + 1], tid);
    __syncthreads();
  }
}

// ---- the same tile computation at 3 CTAs per SM (u8 / f32 sources) ---------------------------------------
// The kernel above needs 128 registers (16 accumulators x 2 phases, the window, and the previous level's 16
// unrounded values for the DoG): 2 CTAs = 16 warps per SM.  Here the previous level lives in shared memory
// ([16][256] doubles, conflict-free), the source tile is stored as float (u8 and f32 samples are exact in
// float; for u8 the 1/255 of image-utils.js:114 is folded into the H-pass taps), and one Ts buffer is used:
// 73 KB and <= 85 registers -> 3 CTAs = 24 warps per SM.
#define F0H_S_FLOATS (F0_SROWS * F0_SPITCH)                         // 2352 floats
#define F0H_TS_OFF ((F0H_S_FLOATS / 2 + 1) & ~1)                    // in doubles, 16-byte aligned
#define F0H_WV_OFF (F0H_TS_OFF + ((F0_TROWS * F0_TPITCH + 1) & ~1))
#define F0H_WH_OFF(nlev) (F0H_WV_OFF + (nlev) * 2 * F0_WSTRIDE)
#define F0H_P_OFF(nlev) (F0H_WH_OFF(nlev) + (nlev) * 2 * F0_WSTRIDE)
#define F0H_DOUBLES(nlev) (F0H_P_OFF(nlev) + 16 * F0_THREADS)

__global__ void __launch_bounds__(F0_THREADS, 3)
fused_octave0_hi_kernel(const double *__restrict__ weights, const Fused0Args A)
{
  extern __shared__ double smem[];
  float *S = reinterpret_cast<float *>(smem);           // [48][49] source tile (exact u8 / f32 values)
  double *Ts = smem + F0H_TS_OFF;                        // [49][65] horizontally blurred rows, both phases
  double *Wv = smem + F0H_WV_OFF;                        // V-pass taps [nlev][F0_WSTRIDE]{w0, w1}
  double *Wh = smem + F0H_WH_OFF(A.nlev);                // H-pass taps (scaled by 1/255 for u8 sources)
  double *P = smem + F0H_P_OFF(A.nlev);                  // previous level, unrounded: P[k * 256 + tid]
  const int tid = threadIdx.x;
  const int a_tile = blockIdx.x * F0_SW, b_tile = blockIdx.y * F0_SH;

  {
    const double hscale = A.dtype == SIFT_U8 ? 1.0 / 255.0 : 1.0;
    for (int e = tid; e < A.nlev * 2 * F0_WSTRIDE; e += F0_THREADS) {
      const double wv = __ldg(weights + A.woff + e);
      Wv[e] = wv;
      Wh[e] = wv * hscale;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) {                          // 48*48 samples, 9 per thread
    const int e = tid + i * F0_THREADS;
    const int rr = e / F0_SCOLS, cc = e - rr * F0_SCOLS;
    const int gy = min(max(b_tile - F0_HALO + rr, 0), A.src_h - 1);          // clamp-to-edge, sift.js:116-119
    const int gx = min(max(a_tile - F0_HALO + cc, 0), A.src_w - 1);
    const char *row = (const char *)A.src + (size_t)gy * A.src_pitch;
    S[rr * F0_SPITCH + cc] = A.dtype == SIFT_U8 ? (float)__ldg((const unsigned char *)row + gx) : __ldg((const float *)row + gx);
  }
  for (int e = tid; e < F0_TPITCH; e += F0_THREADS) Ts[(F0_TROWS - 1) * F0_TPITCH + e] = 0.0;    // slack row
  if (tid < F0_SROWS) S[tid * F0_SPITCH + F0_SCOLS] = 0.f;                                       // pad column

  // V-pass ownership: output column X of the tile, source rows 8*rg .. 8*rg+7, both row phases
  const int X = tid & (2 * F0_SW - 1), rg = tid / (2 * F0_SW);
  const int x = 2 * a_tile + X;
  const int w = A.oct.w, h = A.oct.h;
  const int y_first = 2 * (b_tile + 8 * rg);
  const size_t o_first = (size_t)y_first * A.oct.pitch + x;
  const bool col_ok = x < w;
  const bool rows_full = y_first + 16 <= h;
  const bool seed_lane = A.has_next && (X & 1) == 0 && col_ok;
  double *Pt = P + tid;

  for (int s = 0; s < A.nlev; s++) {                     // octave 0 blurs every level from the base (background.js:110)
    const int R = A.radius[s];
    const int clo0 = -((R + 1) / 2);
    const int n = R + 1 + (R & 1);                       // taps per phase incl. the phase-1 shift for odd R
    __syncthreads();                                     // source tile ready (s = 0) / Ts free again (s > 0)
    {                                                    // ---- H pass: one row x 8 source columns x 2 phases per item
      const int nrows = F0_SH + n - 1;
      if (tid < nrows * (F0_SW / 8)) {
        const int g = (tid >= nrows) + (tid >= 2 * nrows) + (tid >= 3 * nrows), rr = F0_HALO + clo0 + (tid - g * nrows);
        double a0[8], a1[8];
        poly_window<float>(S + rr * F0_SPITCH + F0_HALO + 8 * g + clo0, 1, Wh + s * 2 * F0_WSTRIDE, n, a0, a1);
        double *t = Ts + rr * F0_TPITCH + 16 * g;
#pragma unroll
        for (int k = 0; k < 8; k++) { t[2 * k] = a0[k]; t[2 * k + 1] = a1[k]; }
      }
    }
    __syncthreads();
    // ---- V pass
    double a0[8], a1[8];
    poly_window<double>(Ts + (F0_HALO + 8 * rg + clo0) * F0_TPITCH + X, F0_TPITCH, Wv + s * 2 * F0_WSTRIDE, n, a0, a1);
    // ---- epilogue: G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) against the unrounded previous level in P
    if (col_ok) {
      float *gp = A.oct.gauss[s] + o_first;
      float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + o_first;
      const size_t rowp = (size_t)A.oct.pitch;
      const bool wg = A.keep_gauss != 0, wd = s > 0;
      if (rows_full) {
        if (wg) {
#pragma unroll
          for (int k = 0; k < 8; k++) { *gp = (float)a0[k]; gp += rowp; *gp = (float)a1[k]; gp += rowp; }
        }
        if (wd) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            *dp = (float)(Pt[(2 * k) * F0_THREADS] - a0[k]); dp += rowp;
            *dp = (float)(Pt[(2 * k + 1) * F0_THREADS] - a1[k]); dp += rowp;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            if (y_first + 2 * k + q < h) {
              const double v = q ? a1[k] : a0[k];
              if (wg) gp[(size_t)(2 * k + q) * rowp] = (float)v;
              if (wd) dp[(size_t)(2 * k + q) * rowp] = (float)(Pt[(2 * k + q) * F0_THREADS] - v);
            }
          }
        }
      }
      if (s == A.spo && seed_lane) {                     // in[2a][2b] (matrix2d.js:129): even rows (phase 0), even columns
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int nr = b_tile + 8 * rg + k + A.oct.seed_off;        // row of the next octave (strip-local)
          if (y_first + 2 * k < h && nr >= 0 && nr < A.next.h) {
            A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = a0[k];
            A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)a0[k];
          }
        }
      }
    }
    if (s + 1 < A.nlev) {
#pragma unroll
      for (int k = 0; k < 8; k++) { Pt[(2 * k) * F0_THREADS] = a0[k]; Pt[(2 * k + 1) * F0_THREADS] = a1[k]; }
    }
  }
}

// ---- third form of the tile kernel (default for u8 / f32 sources) -------------------------------------------
// Same tile, same two phases per level and the same arithmetic (bit-identical outputs) as the kernel above, with
// the per-output instruction count cut down:
//   * V pass: a thread owns TWO neighbouring output columns x 4 source rows x 2 row phases (16 accumulators as
//     before): samples arrive as LDS.128 column pairs, the previous level as LDS.128 / STS.128 pairs and every
//     Gaussian / DoG row leaves as ONE 8-byte store per thread (a warp writes 256 contiguous bytes per row);
//   * H pass: its 16 outputs per item leave as 8 STS.128 (Ts rows are 16-byte aligned, pitch 66 doubles:
//     conflict-free for lanes along rows);
//   * per-level parameters come from shared memory and the plane pointers advance by a constant stride (no indexed
//     constant loads behind the barriers), weights are fetched one tap ahead.
#define F3_TPITCH (2 * F0_SW + 2)                                    // 66 doubles: rows start 16-byte aligned
#ifndef F3_WARPS
#define F3_WARPS 8
#endif
#ifndef F3_CTAS
#define F3_CTAS 3
#endif
template <int NW> struct F3 {
  static constexpr int THREADS = 32 * NW;
  static constexpr int SH = 4 * NW;                                  // source rows per tile
  static constexpr int SROWS = SH + 2 * F0_HALO;
  static constexpr int TROWS = SROWS + 1;                            // +1 slack row for the window prefetch
  static constexpr int S_FLOATS = SROWS * F0_SPITCH;
  static constexpr int TS_OFF = (S_FLOATS / 2 + 2) & ~1;             // doubles
  static constexpr int P_OFF = TS_OFF + TROWS * F3_TPITCH;
  static constexpr int WV_OFF = P_OFF + 16 * THREADS;
  static constexpr int doubles(int nlev) { return WV_OFF + 2 * nlev * 2 * F0_WSTRIDE; }
};

struct Fused3Args {
  Fused0Args a;
  long long plane;          // floats between consecutive Gaussian (and DoG) planes
  long long dog_delta;      // dog[s-1] = gauss[s] + dog_delta
};

// a0[k] (row phase 0) and a1[k] (row phase 1) of 4 neighbouring source rows for a pair of columns:
//   a_p[k] = sum_{j<n} w_p[j] v[k+j],  v[r] = the double2 at base + r * F3_TPITCH (doubles)
#define POLY2_STEP(U, JJ)                                                                   \
  {                                                                                         \
    const double2 c = w2[(JJ)];                                                             \
    _Pragma("unroll") for (int k = 0; k < 4; k++) {                                         \
      const double2 v = vw[(k + (U)) & 3];                                                  \
      a0[k].x = fma(c.x, v.x, a0[k].x); a0[k].y = fma(c.x, v.y, a0[k].y);                   \
      a1[k].x = fma(c.y, v.x, a1[k].x); a1[k].y = fma(c.y, v.y, a1[k].y);                   \
    }                                                                                       \
    vw[(U) & 3] = *reinterpret_cast<const double2 *>(nxt + (JJ) * F3_TPITCH);               \
  }
__device__ __forceinline__ void poly_window2(const double *__restrict__ base, const double *__restrict__ w, const int n,
                                             double2 (&a0)[4], double2 (&a1)[4])
{
  const double2 *__restrict__ w2 = reinterpret_cast<const double2 *>(w);
  double2 vw[4];
#pragma unroll
  for (int k = 0; k < 4; k++) vw[k] = *reinterpret_cast<const double2 *>(base + k * F3_TPITCH);
  const double *nxt = base + 4 * F3_TPITCH;
  {                                                          // the first tap initialises the accumulators
    const double2 c = w2[0];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      a0[k].x = c.x * vw[k].x; a0[k].y = c.x * vw[k].y;
      a1[k].x = c.y * vw[k].x; a1[k].y = c.y * vw[k].y;
    }
    vw[0] = *reinterpret_cast<const double2 *>(nxt);
  }
  int j = 1;
  for (; j + 4 <= n; j += 4) { POLY2_STEP(1, j) POLY2_STEP(2, j + 1) POLY2_STEP(3, j + 2) POLY2_STEP(0, j + 3) }
  const int rem = n - j;                                     // 0..3, uniform over the CTA
  if (rem & 2) { POLY2_STEP(1, j) POLY2_STEP(2, j + 1) }
  if (rem & 1) {
    if (rem & 2) POLY2_STEP(3, j + 2) else POLY2_STEP(1, j)
  }
}

template <int NW>
__global__ void __launch_bounds__(32 * NW, F3_CTAS)
fused_octave0_hi3_kernel(const double *__restrict__ weights, const Fused3Args B)
{
  typedef F3<NW> G;
  const Fused0Args &A = B.a;
  extern __shared__ __align__(16) double smem[];
  __shared__ int lvR[SIFT_MAX_LEVELS];
  float *S = reinterpret_cast<float *>(smem);           // [SROWS][49] source tile (exact u8 / f32 values)
  double *Ts = smem + G::TS_OFF;                         // [TROWS][66] horizontally blurred rows, both phases
  double2 *P = reinterpret_cast<double2 *>(smem + G::P_OFF);   // previous level, unrounded: P[i * THREADS + tid]
  double *Wv = smem + G::WV_OFF;                         // V-pass taps [nlev][F0_WSTRIDE]{w0, w1}
  double *Wh = Wv + A.nlev * 2 * F0_WSTRIDE;             // H-pass taps (scaled by 1/255 for u8 sources)
  const int tid = threadIdx.x;
  const int a_tile = blockIdx.x * F0_SW, b_tile = blockIdx.y * G::SH;

  {
    const double hscale = A.dtype == SIFT_U8 ? 1.0 / 255.0 : 1.0;
    for (int e = tid; e < A.nlev * 2 * F0_WSTRIDE; e += G::THREADS) {
      const double wv = __ldg(weights + A.woff + e);
      Wv[e] = wv;
      Wh[e] = wv * hscale;
    }
    if (tid < A.nlev) lvR[tid] = A.radius[tid];
  }
#pragma unroll
  for (int i = 0; i < (G::SROWS * F0_SCOLS + G::THREADS - 1) / G::THREADS; i++) {
    const int e = tid + i * G::THREADS;
    if ((G::SROWS * F0_SCOLS) % G::THREADS == 0 || e < G::SROWS * F0_SCOLS) {
      const int rr = e / F0_SCOLS, cc = e - rr * F0_SCOLS;
      const int gy = min(max(b_tile - F0_HALO + rr, 0), A.src_h - 1);          // clamp-to-edge, sift.js:116-119
      const int gx = min(max(a_tile - F0_HALO + cc, 0), A.src_w - 1);
      const char *row = (const char *)A.src + (size_t)gy * A.src_pitch;
      S[rr * F0_SPITCH + cc] = A.dtype == SIFT_U8 ? (float)__ldg((const unsigned char *)row + gx) : __ldg((const float *)row + gx);
    }
  }
  for (int e = tid; e < F3_TPITCH; e += G::THREADS) Ts[(G::TROWS - 1) * F3_TPITCH + e] = 0.0;    // slack row
  if (tid < G::SROWS) S[tid * F0_SPITCH + F0_SCOLS] = 0.f;                                       // pad column

  // V-pass ownership: output columns X, X+1 of the tile, source rows 4*rg .. 4*rg+3, both row phases
  const int X = 2 * (tid & 31), rg = tid >> 5;
  const int x = 2 * a_tile + X;
  const int y_first = 2 * (b_tile + 4 * rg);
  // row pairs of this thread inside the octave (its width and height are even): 0 = nothing to store
  const int rows_ok = x < A.oct.w ? min(4, (A.oct.h - y_first) >> 1) : 0;
  const unsigned rowp = (unsigned)A.oct.pitch;
  float *gp_level = A.oct.gauss[0] + ((size_t)y_first * rowp + x);
  double2 *Pt = P + tid;

  for (int s = 0; s < A.nlev; s++) {                     // octave 0 blurs every level from the base (background.js:110)
    __syncthreads();                                     // source tile + level table ready (s = 0) / Ts free again (s > 0)
    const int R = lvR[s];
    const int clo0 = -((R + 1) / 2);
    const int n = R + 1 + (R & 1);                       // taps per phase incl. the phase-1 shift for odd R
    {                                                    // ---- H pass: one row x 8 source columns x 2 phases per item
      const int nrows = G::SH + n - 1;
      if (tid < nrows * (F0_SW / 8)) {
        const int g = (tid >= nrows) + (tid >= 2 * nrows) + (tid >= 3 * nrows), rr = F0_HALO + clo0 + (tid - g * nrows);
        double a0[8], a1[8];
        poly_window<float>(S + rr * F0_SPITCH + F0_HALO + 8 * g + clo0, 1, Wh + s * 2 * F0_WSTRIDE, n, a0, a1);
        double2 *t = reinterpret_cast<double2 *>(Ts + rr * F3_TPITCH + 16 * g);
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = make_double2(a0[k], a1[k]);
      }
    }
    __syncthreads();
    // ---- V pass
    double2 a0[4], a1[4];
    poly_window2(Ts + (F0_HALO + 4 * rg + clo0) * F3_TPITCH + X, Wv + s * 2 * F0_WSTRIDE, n, a0, a1);
    // ---- epilogue: G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) against the unrounded previous level in P
    if (rows_ok > 0) {
      const bool wg = A.keep_gauss != 0, wd = s > 0;
      if (rows_ok >= 4) {
        if (wg) {
          float *gp = gp_level;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            *reinterpret_cast<float2 *>(gp) = make_float2((float)a0[k].x, (float)a0[k].y); gp += rowp;
            *reinterpret_cast<float2 *>(gp) = make_float2((float)a1[k].x, (float)a1[k].y); gp += rowp;
          }
        }
        if (wd) {
          float *dp = gp_level + B.dog_delta;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double2 p0 = Pt[k * G::THREADS], p1 = Pt[(4 + k) * G::THREADS];
            *reinterpret_cast<float2 *>(dp) = make_float2((float)(p0.x - a0[k].x), (float)(p0.y - a0[k].y)); dp += rowp;
            *reinterpret_cast<float2 *>(dp) = make_float2((float)(p1.x - a1[k].x), (float)(p1.y - a1[k].y)); dp += rowp;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (k < rows_ok) {
            float *gp = gp_level + (size_t)(2 * k) * rowp, *dp = gp + B.dog_delta;
            if (wg) {
              *reinterpret_cast<float2 *>(gp) = make_float2((float)a0[k].x, (float)a0[k].y);
              *reinterpret_cast<float2 *>(gp + rowp) = make_float2((float)a1[k].x, (float)a1[k].y);
            }
            if (wd) {
              const double2 p0 = Pt[k * G::THREADS], p1 = Pt[(4 + k) * G::THREADS];
              *reinterpret_cast<float2 *>(dp) = make_float2((float)(p0.x - a0[k].x), (float)(p0.y - a0[k].y));
              *reinterpret_cast<float2 *>(dp + rowp) = make_float2((float)(p1.x - a1[k].x), (float)(p1.y - a1[k].y));
            }
          }
        }
      }
      if (s == A.spo && A.has_next) {                    // in[2a][2b] (matrix2d.js:129): even rows (phase 0), even columns (.x)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int nr = b_tile + 4 * rg + k + A.oct.seed_off;        // row of the next octave (strip-local)
          if (k < rows_ok && nr >= 0 && nr < A.next.h) {
            A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = a0[k].x;
            A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)a0[k].x;
          }
        }
      }
    }
    if (s + 1 < A.nlev) {
#pragma unroll
      for (int k = 0; k < 4; k++) { Pt[k * G::THREADS] = a0[k]; Pt[(4 + k) * G::THREADS] = a1[k]; }
    }
    gp_level += B.plane;
  }
}

// Host: zero-padded merged polyphase taps of one level, the two phases interleaved: out[F0_WSTRIDE]{w0, w1}.
//   phase 0: w0[j] = W0[j], j = c - floor(-R/2);  phase 1: w1[j + d] = W0[R - j], d = R & 1.
void fused0_merge_taps(const double *w, int R, double *out)
{
  double w0[F0_WSTRIDE], w1[F0_WSTRIDE];
  for (int i = 0; i < F0_WSTRIDE; i++) w0[i] = w1[i] = 0.0;
  const int clo0 = -((R + 1) / 2), d = R & 1;
  for (int j = 0; j <= R; j++) {
    const int c = clo0 + j;
    const int i0 = 2 * c, i1 = 2 * c + 1;
    const double t0 = (i0 >= -R && i0 <= R) ? w[i0 + R] : 0.0;
    const double t1 = (i1 >= -R && i1 <= R) ? w[i1 + R] : 0.0;
    w0[j] = t0 + t1;
  }
  for (int j = 0; j <= R; j++) w1[j + d] = w0[R - j];
  for (int i = 0; i < F0_WSTRIDE; i++) { out[2 * i] = w0[i]; out[2 * i + 1] = w1[i]; }
}

int fused0_taps_per_level(void) { return 2 * F0_WSTRIDE; }

bool fused0_supported(const LevelPlan *plans, int nlev)
{
  for (int s = 0; s < nlev; s++)
    if (plans[s].radius < 1 || plans[s].radius > F0_MAXR) return false;
  return true;
}

void launch_fused_octave0(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                          const OctaveDev &oct, const OctaveDev *next, const double *d_weights,
                          const LevelPlan *plans, int poly_woff, int nlev, int spo, int keep_gauss,
                          const double *d_u8lut)
{
  Fused0Args A;
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < nlev; s++) A.radius[s] = plans[s].radius;
  A.woff = poly_woff;
  A.u8lut = d_u8lut;
  dim3 grid((src_w + F0_SW - 1) / F0_SW, (src_h + F0_SH - 1) / F0_SH);
  static const bool no_hi = getenv("SIFT_B200_FUSED0_LO") != nullptr;
  static const bool no_hi3 = getenv("SIFT_B200_FUSED0_HI1") != nullptr;
  if (!no_hi && !no_hi3 && (dtype == SIFT_U8 || dtype == SIFT_F32) && nlev >= 2) {
    // the pair-column kernel walks the planes with a constant stride: check the layout it assumes
    Fused3Args B;
    B.a = A;
    B.plane = oct.gauss[1] - oct.gauss[0];
    B.dog_delta = oct.dog[0] - oct.gauss[1];
    bool uniform = (oct.pitch & 1) == 0 && (oct.w & 1) == 0 && (oct.h & 1) == 0 && (((uintptr_t)oct.gauss[0] | (uintptr_t)oct.dog[0]) & 7) == 0 &&
                   (B.plane & 1) == 0;
    for (int s = 0; s < nlev; s++) uniform = uniform && oct.gauss[s] == oct.gauss[0] + (long long)s * B.plane;
    for (int s = 0; s + 1 < nlev; s++) uniform = uniform && oct.dog[s] == oct.gauss[s + 1] + B.dog_delta;
    size_t smem3 = (size_t)F3<F3_WARPS>::doubles(nlev) * sizeof(double);
    if (uniform && smem3 <= 75 * 1024) {
      static const char *cap3 = getenv("SIFT_B200_OCT0_CTAS");
      if (cap3 && cap3[0] == '2') smem3 = 78 * 1024;
      dim3 grid3((src_w + F0_SW - 1) / F0_SW, (src_h + F3<F3_WARPS>::SH - 1) / F3<F3_WARPS>::SH);
      cudaFuncSetAttribute(fused_octave0_hi3_kernel<F3_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
      fused_octave0_hi3_kernel<F3_WARPS><<<grid3, 32 * F3_WARPS, smem3, st>>>(d_weights, B);
      return;
    }
  }
  if (!no_hi && (dtype == SIFT_U8 || dtype == SIFT_F32)) {
    size_t smem_hi = (size_t)F0H_DOUBLES(nlev) * sizeof(double);
    // SIFT_B200_OCT0_CTAS=2: ask for more shared memory than the tile needs so that only two CTAs fit an SM and a
    // third of its registers / shared memory stays free for the scan / refine CTAs of the other frames in flight
    static const char *cap = getenv("SIFT_B200_OCT0_CTAS");
    if (smem_hi <= 75 * 1024) {
      if (cap && cap[0] == '2') smem_hi = 78 * 1024;
      if (cap && cap[0] == '1') smem_hi = 120 * 1024;
      cudaFuncSetAttribute(fused_octave0_hi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hi);
      fused_octave0_hi_kernel<<<grid, F0_THREADS, smem_hi, st>>>(d_weights, A);
      return;
    }
  }
  const size_t smem = (size_t)(F0_SROWS * F0_SPITCH + 2 * F0_TROWS * F0_TPITCH + nlev * 2 * F0_WSTRIDE) * sizeof(double);
  cudaFuncSetAttribute(fused_octave0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fused_octave0_kernel<<<grid, F0_THREADS, smem, st>>>(d_weights, A);
}
