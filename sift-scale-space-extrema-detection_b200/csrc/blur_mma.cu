// blur_mma.cu -- the Gaussian scale-space kernels restated as banded-Toeplitz products on the fp64 matrix
// instruction (mma.sync.m8n8k4.f64 = DMMA.8x8x4).
//
// Why: every blur kernel of this engine is bound by the fp64 pipe (float64 accumulation is needed
// for parity), and the scalar-FMA forms reach 35-40 % of it: a warp DFMA occupies the pipe for two cycles, so the
// tap loops need an issue slot every other cycle and everything else (loads, conversions, stores, barriers)
// competes for the rest.  DMMA.8x8x4 runs at the same fp64 rate on this B200 (37.0 vs 36.5 TFLOP/s,
// tools/micro/fp64_pipes.cu) but carries 256 multiply-adds per warp instruction -- one issue slot per 16 pipe
// cycles, operands from registers once per 8x8 block -- which turns the same arithmetic into a pipe-bound loop.
//
// A 1-D FIR  out[p] = sum_j w[j] v[p + j]  over 8 neighbouring positions and 8 independent lines is the product
// W (8 x K) * V (K x 8) of a banded Toeplitz weight matrix with the sample block; K runs over the 8 + taps - 1
// samples in chunks of 4 (one DMMA each).  The zero part of the band is multiplied too: 60-85 % of the
// multiply-adds are useful for the polyphase octave 0, 80-95 % for the wider kernels of the later octaves.
//
// Reference path restated (same as blur_fused.cu / blur_sep.cu): Matrix2D_linearResize(input, 0.5)
// (background.js:84, matrix2d.js:112-138), SIFT_blurMatrix2DChunk per level from the octave base
// (background.js:145-210, sift.js:72-149, clamp-to-edge sift.js:116-119), SIFT_subtractMatrix2DChunk
// (sift.js:154-188) and the rate-2.0 resize seeding the next octave (background.js:114-130).
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA): lane = 4 g + t;  A[g][t] (8x4, row), B[t][g] (4x8, col),
// C/D[g][2t], C/D[g][2t+1].
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// =====================================================================================================
// Octave 0: 2x nearest-neighbour upsample + all levels + DoG + seed of octave 1 in one kernel (polyphase).
//
// One CTA (8 warps) = 32 x 32 source pixels = 64 x 64 outputs of every level, as in blur_fused.cu:
//   H pass  S (source tile, fp64, clamped halo of 8)      -> Ts[row][X]   rows = M, X = (position, phase) = N
//   V pass  Ts                                             -> out[y][X]   y = (position, phase) = M, X = N
// The merged polyphase taps of a level (fused0_merge_taps: w0 / w1, n = R + 1 + (R & 1) entries) form the
// 8 x K band  Wm[(a, p)][k] = w_p[k - a],  a = 0..3 positions, p = phase, K = 4 + n - 1 samples; the same
// per-lane fragment  Wf[d] = w_{g&1}[4 d + t - (g >> 1)]  serves as B operand of the H pass and as A operand
// of the V pass.  Each warp keeps the previous level's 32 x 16 outputs (unrounded) in registers for the DoG.
#define M0_SW 32
#define M0_SH 32
#define M0_HALO 8
#define M0_SCOLS (M0_SW + 2 * M0_HALO)   // 48
#define M0_SROWS 56                      // 48 rows hold samples; the rest is zero (read against zero weights only)
#define M0_SPITCH 52                     // = 4 (mod 16): lanes (g, t) -> g * pitch + t hit 16 distinct 8-byte banks per half warp
#define M0_TW_DOUBLES (32 * 16)          // per-warp Ts: 32 rows x 16 columns
#define M0_THREADS 256
#define M0_MAXD 5                        // chunks of 4 samples per 8 x K band: ceil((n + 3) / 4), n <= 17
#define M0_S_DOUBLES (M0_SROWS * M0_SPITCH + 8)
#define M0_T_DOUBLES (8 * M0_TW_DOUBLES)
#define M0_SMEM_DOUBLES(nlev) (M0_S_DOUBLES + M0_T_DOUBLES + (nlev) * M0_MAXD * 32)

struct Mma0Args {
  const void *src;
  size_t src_pitch;
  int src_w, src_h, dtype;
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, nlev;
  int radius[SIFT_MAX_LEVELS];
  const double *wfrag;                   // [nlev][M0_MAXD][32] per-lane band fragments (zero padded)
  long long plane;                       // floats between consecutive Gaussian (and DoG) planes
  long long dog_delta;                   // dog[s-1] = gauss[s] + dog_delta
  int row_shift;                         // tiles start at source row -row_shift: keeps the 4-row blocks aligned with the
                                         // whole image's when this octave is a mosaic strip (bit-identical sums)
};

__device__ __forceinline__ double mma0_u8(unsigned char raw)
{                                        // v / 255.0 (image-utils.js:114), correctly rounded without a divide
  const double r = 1.0 / 255.0, x = (double)raw, q = x * r;
  return fma(fma(-q, 255.0, x), r, q);
}

__device__ __forceinline__ double mma0_sample(const Mma0Args &A, const char *row, int gx)
{
  switch (A.dtype) {
    case SIFT_F32: return (double)((const float *)row)[gx];
    case SIFT_F64: return ((const double *)row)[gx];
    default: {
      const uchar4 c = ((const uchar4 *)row)[gx];
      const double v = __dadd_rn(__dadd_rn(__dmul_rn((double)c.x, 0.299), __dmul_rn((double)c.y, 0.587)),
                                 __dmul_rn((double)c.z, 0.114));                              // image-utils.js:107
      return v / 255.0;
    }
  }
}

// Both passes of one level for ONE WARP, no CTA barrier: the warp blurs the rows its own vertical window needs
// (4 D + 12 rows x its 16 columns; 12 % more multiply-adds than sharing the rows across the CTA) into its private
// 4 KB of Ts and consumes them itself, so the eight warps of a CTA drift through H pass / V pass / epilogue
// independently and cover each other's latencies.  D = chunks of the band (CTA-uniform), acc = the warp's 32 x 16 outputs.
// Ts layout: [32 rows][16 columns], column index XOR-swizzled by the row (8 (r & 1) + 4 ((r >> 1) & 1)): the
// H-pass 16-byte stores (lanes along rows) and the V-pass 8-byte loads (lanes along 4 rows x 8 columns) are both
// conflict-free without padding.
template <int D>
__device__ __forceinline__ void mma0_level(const double *__restrict__ S, double *__restrict__ Tw,
                                           const double *__restrict__ wfs, const int clo0, const int warp,
                                           const int lane, double (&acc)[4][2][2])
{
  constexpr int MBH = (4 * D + 12 + 7) / 8;             // 8-row blocks of H-pass rows: 3 (D <= 3) or 4
  const int g = lane >> 2, t = lane & 3;
  const int wy = warp >> 2, wx = warp & 3;
  double wf[D];
#pragma unroll
  for (int d = 0; d < D; d++) wf[d] = wfs[d * 32 + lane];
  {
    // ---- H pass: base columns 16 wx .. + 15 (two N blocks = source positions 8 wx .. + 7), rows from the first
    //      row of this warp's vertical window
    const double *sp = S + (M0_HALO + 16 * wy + clo0 + g) * M0_SPITCH + M0_HALO + 8 * wx + clo0 + t;
    double h[MBH][2][2];
#pragma unroll
    for (int mb = 0; mb < MBH; mb++)
#pragma unroll
      for (int nb = 0; nb < 2; nb++) h[mb][nb][0] = h[mb][nb][1] = 0.0;
#pragma unroll
    for (int c = 0; c <= D; c++) {
      double a[MBH];
#pragma unroll
      for (int mb = 0; mb < MBH; mb++) a[mb] = sp[mb * 8 * M0_SPITCH + 4 * c];
#pragma unroll
      for (int nb = 0; nb < 2; nb++) {
        const int d = c - nb;
        if (d >= 0 && d < D) {
#pragma unroll
          for (int mb = 0; mb < MBH; mb++) dmma884(h[mb][nb][0], h[mb][nb][1], a[mb], wf[d]);
        }
      }
    }
    __syncwarp();                                       // the previous level's V pass has read Tw
    const int sw = 8 * (g & 1) + 4 * ((g >> 1) & 1);
    double *tp = Tw + g * 16;
#pragma unroll
    for (int mb = 0; mb < MBH; mb++)
#pragma unroll
      for (int nb = 0; nb < 2; nb++)
        *reinterpret_cast<double2 *>(tp + mb * 128 + ((8 * nb + 2 * t) ^ sw)) = make_double2(h[mb][nb][0], h[mb][nb][1]);
  }
  __syncwarp();
  // ---- V pass: 4 M blocks (4 source rows x 2 phases = 8 output rows each) x 2 N blocks (8 columns each)
  {
    const int sw = 8 * (t & 1) + 4 * (t >> 1);
    const double *tp0 = Tw + t * 16 + (g ^ sw), *tp1 = Tw + t * 16 + ((g ^ sw) ^ 8);
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
      for (int nb = 0; nb < 2; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
    for (int c = 0; c < 3 + D; c++) {
      const double b0 = tp0[c * 64], b1 = tp1[c * 64];
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int d = c - mb;
        if (d >= 0 && d < D) {
          dmma884(acc[mb][0][0], acc[mb][0][1], wf[d], b0);
          dmma884(acc[mb][1][0], acc[mb][1][1], wf[d], b1);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(M0_THREADS, 2)
oct0_mma_kernel(const Mma0Args A)
{
  extern __shared__ __align__(16) double smem[];
  __shared__ int lvR[SIFT_MAX_LEVELS];
  double *S = smem;                                     // [56][52] source tile v / 255 (rows / columns >= 48: zero)
  double *Tw = smem + M0_S_DOUBLES + (threadIdx.x >> 5) * M0_TW_DOUBLES;   // this warp's horizontally blurred rows
  double *Wf = smem + M0_S_DOUBLES + M0_T_DOUBLES;                       // [nlev][M0_MAXD][32] band fragments
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int a_tile = blockIdx.x * M0_SW, b_tile = blockIdx.y * M0_SH - A.row_shift;

  if (tid < A.nlev) lvR[tid] = A.radius[tid];
  for (int e = tid; e < A.nlev * M0_MAXD * 32; e += M0_THREADS) Wf[e] = __ldg(A.wfrag + e);
  {
    static_assert(M0_SCOLS * M0_SCOLS == 9 * M0_THREADS, "tile load assumes 9 samples per thread");
    int so[9];
    const char *rowp[9];
    int gxs[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int e = tid + i * M0_THREADS;
      const int rr = e / M0_SCOLS, cc = e - rr * M0_SCOLS;
      const int gy = min(max(b_tile - M0_HALO + rr, 0), A.src_h - 1);          // clamp-to-edge, sift.js:116-119
      gxs[i] = min(max(a_tile - M0_HALO + cc, 0), A.src_w - 1);
      so[i] = rr * M0_SPITCH + cc;
      rowp[i] = (const char *)A.src + (size_t)gy * A.src_pitch;
    }
    if (A.dtype == SIFT_U8) {
      unsigned char raw[9];
#pragma unroll
      for (int i = 0; i < 9; i++) raw[i] = __ldg((const unsigned char *)rowp[i] + gxs[i]);
#pragma unroll
      for (int i = 0; i < 9; i++) S[so[i]] = mma0_u8(raw[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 9; i++) S[so[i]] = mma0_sample(A, rowp[i], gxs[i]);
    }
    // everything else of S is read against zero weights only: keep it finite
    for (int e = tid; e < M0_SCOLS * (M0_SPITCH - M0_SCOLS); e += M0_THREADS)
      S[(e >> 2) * M0_SPITCH + M0_SCOLS + (e & 3)] = 0.0;
    for (int e = M0_SCOLS * M0_SPITCH + tid; e < M0_S_DOUBLES; e += M0_THREADS) S[e] = 0.0;
  }
  __syncthreads();                                      // source tile, level table, fragments staged: the only CTA barrier

  // V-pass ownership of this warp: output rows 32 wy .. +31 (source rows 16 wy .. +15), columns 16 wx .. +15
  const int wy = warp >> 2, wx = warp & 3;
  const int y0 = 2 * (b_tile + 16 * wy) + g;            // + 8 mb: output row of fragment row g
  const int x0 = 2 * a_tile + 16 * wx + 2 * t;          // + 8 nb: first of the two columns of this lane
  const int ow = A.oct.w, oh = A.oct.h;
  const size_t row8 = (size_t)8 * A.oct.pitch;
  const bool interior = b_tile >= 0 && 2 * (b_tile + M0_SH) <= oh && 2 * (a_tile + M0_SW) <= ow;   // CTA-uniform
  float *gthr = A.oct.gauss[0] + ((long long)y0 * A.oct.pitch + x0);      // this lane's first output of level 0

  // one level: both passes, then the epilogue -- G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) from the unrounded
  // accumulators, seed of octave 1.  `cur` receives level s, `prv` holds level s - 1 (the caller swaps them).
  auto level = [&](const int s, double (&cur)[4][2][2], const double (&prv)[4][2][2]) {
    const int R = lvR[s];
    const int clo0 = -((R + 1) / 2);
    const int n = R + 1 + (R & 1);                      // taps per phase incl. the phase-1 shift for odd R
    const int D = (n + 6) >> 2;                         // chunks of the band: ceil((n + 3) / 4)
    const double *wfs = Wf + s * M0_MAXD * 32;
    switch (D) {                                        // CTA-uniform
      case 2: mma0_level<2>(S, Tw, wfs, clo0, warp, lane, cur); break;
      case 3: mma0_level<3>(S, Tw, wfs, clo0, warp, lane, cur); break;
      case 4: mma0_level<4>(S, Tw, wfs, clo0, warp, lane, cur); break;
      default: mma0_level<5>(S, Tw, wfs, clo0, warp, lane, cur); break;
    }
    const bool wg = A.keep_gauss != 0, wd = s > 0;
    float *gl = gthr + (long long)s * A.plane;
    if (interior) {
      if (wg) {
        float *r = gl;
#pragma unroll
        for (int mb = 0; mb < 4; mb++, r += row8) {
          *reinterpret_cast<float2 *>(r) = make_float2((float)cur[mb][0][0], (float)cur[mb][0][1]);
          *reinterpret_cast<float2 *>(r + 8) = make_float2((float)cur[mb][1][0], (float)cur[mb][1][1]);
        }
      }
      if (wd) {
        float *r = gl + A.dog_delta;
#pragma unroll
        for (int mb = 0; mb < 4; mb++, r += row8) {
          *reinterpret_cast<float2 *>(r) = make_float2((float)(prv[mb][0][0] - cur[mb][0][0]), (float)(prv[mb][0][1] - cur[mb][0][1]));
          *reinterpret_cast<float2 *>(r + 8) = make_float2((float)(prv[mb][1][0] - cur[mb][1][0]), (float)(prv[mb][1][1] - cur[mb][1][1]));
        }
      }
    } else {
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int y = y0 + 8 * mb;
        if (y >= 0 && y < oh) {
#pragma unroll
          for (int nb = 0; nb < 2; nb++) {
            if (x0 + 8 * nb < ow) {                     // the octave width is even: both columns or none
              float *r = gl + mb * row8 + 8 * nb;
              if (wg) *reinterpret_cast<float2 *>(r) = make_float2((float)cur[mb][nb][0], (float)cur[mb][nb][1]);
              if (wd)
                *reinterpret_cast<float2 *>(r + A.dog_delta) =
                    make_float2((float)(prv[mb][nb][0] - cur[mb][nb][0]), (float)(prv[mb][nb][1] - cur[mb][nb][1]));
            }
          }
        }
      }
    }
    if (A.has_next && s == A.spo && (g & 1) == 0) {     // in[2a][2b] (matrix2d.js:129): even rows, even columns
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int y = y0 + 8 * mb;
        const int nr = (y >> 1) + A.oct.seed_off;       // row of the next octave (strip-local)
        if (y >= 0 && y < oh && nr >= 0 && nr < A.next.h) {
#pragma unroll
          for (int nb = 0; nb < 2; nb++) {
            const int x = x0 + 8 * nb;
            if (x < ow) {
              A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = cur[mb][nb][0];
              A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)cur[mb][nb][0];
            }
          }
        }
      }
    }
  };

  double pa[4][2][2], pb[4][2][2];
#pragma unroll
  for (int mb = 0; mb < 4; mb++)
#pragma unroll
    for (int nb = 0; nb < 2; nb++) pb[mb][nb][0] = pb[mb][nb][1] = 0.0;
  for (int s = 0; s < A.nlev; s += 2) {                 // octave 0 blurs every level from the base (background.js:110)
    level(s, pa, pb);
    if (s + 1 < A.nlev) level(s + 1, pb, pa);
  }
}

// ---- host side of octave 0 ------------------------------------------------------------------------------
int mma0_frag_doubles(int nlev) { return nlev * M0_MAXD * 32; }

bool mma0_supported(const LevelPlan *plans, int nlev)
{
  if (nlev < 2) return false;
  for (int s = 0; s < nlev; s++)
    if (plans[s].radius < 1 || plans[s].radius > 16) return false;
  return true;
}

// Per-lane band fragments of one level from its merged polyphase taps (fused0_merge_taps: out[2 i] = w0[i], out[2 i + 1] = w1[i]).
void mma0_build_frags(const double *merged, int R, double *out /* M0_MAXD * 32 */)
{
  const int n = R + 1 + (R & 1);
  for (int d = 0; d < M0_MAXD; d++)
    for (int lane = 0; lane < 32; lane++) {
      const int g = lane >> 2, t = lane & 3;
      const int j = 4 * d + t - (g >> 1);
      out[d * 32 + lane] = (j >= 0 && j < n) ? merged[2 * j + (g & 1)] : 0.0;
    }
}

bool launch_oct0_mma(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                     const OctaveDev &oct, const OctaveDev *next, const double *d_frags, const LevelPlan *plans,
                     int nlev, int spo, int keep_gauss)
{
  Mma0Args A;
  memset(&A, 0, sizeof A);
  A.plane = oct.gauss[1] - oct.gauss[0];
  A.dog_delta = oct.dog[0] - oct.gauss[1];
  // the kernel walks the planes with a constant stride and stores column pairs: check the layout it assumes
  bool ok = (oct.pitch & 1) == 0 && (oct.w & 1) == 0 && (oct.h & 1) == 0 && (A.plane & 1) == 0 && (A.dog_delta & 1) == 0 &&
            (((uintptr_t)oct.gauss[0]) & 7) == 0;
  for (int s = 0; s < nlev; s++) ok = ok && oct.gauss[s] == oct.gauss[0] + (long long)s * A.plane;
  for (int s = 0; s + 1 < nlev; s++) ok = ok && oct.dog[s] == oct.gauss[s + 1] + A.dog_delta;
  if (!ok) return false;
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < nlev; s++) A.radius[s] = plans[s].radius;
  A.wfrag = d_frags;
  A.row_shift = (oct.y_top >> 1) & 3;
  const size_t smem = (size_t)M0_SMEM_DOUBLES(nlev) * sizeof(double);
  dim3 grid((src_w + M0_SW - 1) / M0_SW, (src_h + A.row_shift + M0_SH - 1) / M0_SH);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(oct0_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  oct0_mma_kernel<<<grid, M0_THREADS, smem, st>>>(A);
  return true;
}
