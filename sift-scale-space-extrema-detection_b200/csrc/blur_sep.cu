// blur_sep.cu -- the two-pass separable Gaussian scale-space kernels of octaves >= 1 (and of octave 0
// when the fused polyphase kernel does not apply), as ONE kernel used twice.
//
// Restates SIFT_blurMatrix2DChunk (src/sift.js:72-149, clamped correlation, here separable and in fp64)
// applied level by level to the octave base (background.js:103-224), SIFT_subtractMatrix2DChunk
// (src/sift.js:154-188, finer minus coarser) and the rate-2.0 resize that seeds the next octave
// (matrix2d.js:112-138, background.js:114-130).
//
// fir_pass: "filter along the contiguous axis b of in[a][b], write out[b][a]".
//   pass A (horizontal)  in = octave base [y][x]      out = T_s^T [x][y]   (fp64 intermediate, L2 resident)
//   pass B (vertical)    in = T_s^T [x][y]            out = G_s, D_{s-1} [y][x] fp32 (+ the next seed)
// A lane owns one line a, a warp 32 neighbouring lines, a thread 8 neighbouring outputs along b.
//   * global reads run along b (contiguous), global writes run along a across the lanes (contiguous):
//     both sides of both passes are fully coalesced, with no transposition kernel.
//   * the CTA stages its 32 lines x (outputs + 2 Rmax) samples ONCE, transposed ([b][33]), and every level
//     is computed from that tile: lanes read consecutive words (no bank conflicts, no padding arithmetic)
//     and every address in the tap loop is `pointer + compile-time constant`.
//   * the tap loop is a register sliding window: one sample load and one (broadcast) weight load per
//     8 DFMAs, 2R+1 taps exactly.
//   * pass B keeps the previous level's unrounded fp64 values in registers, so DoG is formed from the
//     accumulators and each Gaussian / DoG value is written exactly once.
#include <cstdlib>
#include <cstring>
#include "common.cuh"

#define FP_LINES 32                 // lines (a) per CTA = lanes
#define FP_PITCH (FP_LINES + 1)     // odd pitch: the transposing stores of the staging loop spread over the banks

struct FirArgs {
  // ---- input: na lines of nb samples; the filter runs along b
  const void *src;
  int src_kind;          // SIFT_F64 dense doubles, or a source image dtype (octave-0 generic pass A)
  size_t in_pitch;       // bytes between lines
  int in_nb;             // samples per stored line (== nb, or nb / 2 when `ups`)
  int na, nb;
  int ups;               // input stored at half resolution along b: sample b is in[.][b >> 1] (matrix2d.js:129)
  int a_ups;             // pass B of an upsampled octave-0: output line a (x) is real, nothing to do; kept for clarity
  // ---- levels
  int nlev, rmax, wtotal;
  int level[SIFT_MAX_LEVELS], radius[SIFT_MAX_LEVELS], woff[SIFT_MAX_LEVELS];
  // ---- output
  int mode;              // 0: T^T planes, 1: Gaussian / DoG / seed
  double *T[SIFT_MAX_LEVELS];
  size_t t_pitch;        // doubles per T^T line
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, seed_is_level0;
};

// a[k] = sum_{j<npad} w[j] * v[k + j], k < NO; v[p] = base[p * FP_PITCH].  The taps are zero-padded to a
// multiple of 4 (npad) and 16-byte aligned: the loop runs whole groups of 4 taps, the weights of the next
// group are fetched (2 x LDS.128, broadcast) while the current group is consumed, the window of NO samples
// rotates through registers with compile-time indices, and the accumulators never change registers.
// Positions up to npad + NO - 1 are read.
template <int NO>
__device__ __forceinline__ void fir_window(const double *__restrict__ w, const int npad,
                                           const double *__restrict__ base, double (&a)[NO])
{
  double vw[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) { vw[k] = base[k * FP_PITCH]; a[k] = 0.0; }
  const double *nxt = base + NO * FP_PITCH;
  double2 wa = *reinterpret_cast<const double2 *>(w), wb = *reinterpret_cast<const double2 *>(w + 2);
#define FIR_GROUP(G)                                                                              \
  {                                                                                               \
    const double c[4] = { wa.x, wa.y, wb.x, wb.y };                                               \
    wa = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 4);                                 \
    wb = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 6);                                 \
    _Pragma("unroll") for (int u = 0; u < 4; u++) {                                               \
      _Pragma("unroll") for (int k = 0; k < NO; k++) a[k] = fma(c[u], vw[(k + 4 * (G) + u) & (NO - 1)], a[k]); \
      vw[(4 * (G) + u) & (NO - 1)] = nxt[(j + 4 * (G) + u) * FP_PITCH];                            \
    }                                                                                             \
  }
  int j = 0;
  for (; j + NO <= npad; j += NO) {
    FIR_GROUP(0) FIR_GROUP(1)
    if (NO == 16) { FIR_GROUP(2) FIR_GROUP(3) }
  }
  const int rem = npad - j;                                   // 0, 4, ... NO - 4 (warp-uniform)
  if (rem >= 4) FIR_GROUP(0)
  if (NO == 16) {
    if (rem >= 8) FIR_GROUP(1)
    if (rem >= 12) FIR_GROUP(2)
  }
#undef FIR_GROUP
}

__device__ __forceinline__ double fir_src_at(const void *line, int b, int kind)
{
  switch (kind) {
    case SIFT_U8: return (double)((const unsigned char *)line)[b] / 255.0;        // image-utils.js:114
    case SIFT_F32: return (double)((const float *)line)[b];
    case SIFT_F64: return ((const double *)line)[b];
    default: {
      const uchar4 p = ((const uchar4 *)line)[b];
      // (R*0.299) + (G*0.587) + (B*0.114), then / 255.0 -- image-utils.js:107-114, unfused
      const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)p.x, 0.299), __dmul_rn((double)p.y, 0.587)),
                                 __dmul_rn((double)p.z, 0.114));
      return g / 255.0;
    }
  }
}

__device__ __forceinline__ void fir_cp_async8(double *dst_smem, const double *src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}

// Bring 32 lines x `span` samples of `src` (first sample b_first, clamp-to-edge) into tile[e][line].
// Warp w brings lines w, w + NW, ...; lanes run along b (coalesced reads), stores are transposed.
template <int NW>
__device__ __forceinline__ void fir_stage(const FirArgs &A, const void *src, double *__restrict__ tile, int a0,
                                          int b_first, int span, int lane, int warp)
{
  if (A.src_kind == SIFT_F64 && !A.ups) {
    const bool interior = b_first >= 0 && b_first + span <= A.nb;
#pragma unroll
    for (int i = 0; i < FP_LINES / NW; i++) {
      const int al = warp + i * NW;
      const int a = min(a0 + al, A.na - 1);                 // lines past the end replicate the last one (never stored)
      const double *line = (const double *)((const char *)src + (size_t)a * A.in_pitch);
      double *dst = tile + lane * FP_PITCH + al;
      if (interior) {
        const double *p = line + b_first + lane;
        int e = lane;
        for (; e + 96 < span; e += 128, p += 128, dst += 128 * FP_PITCH) {
          fir_cp_async8(dst, p);
          fir_cp_async8(dst + 32 * FP_PITCH, p + 32);
          fir_cp_async8(dst + 64 * FP_PITCH, p + 64);
          fir_cp_async8(dst + 96 * FP_PITCH, p + 96);
        }
        for (; e < span; e += 32, p += 32, dst += 32 * FP_PITCH) fir_cp_async8(dst, p);
      } else {
        for (int e = lane; e < span; e += 32, dst += 32 * FP_PITCH) {
          const int b = min(max(b_first + e, 0), A.nb - 1);  // sift.js:116-119 clamp-to-edge
          fir_cp_async8(dst, line + b);
        }
      }
    }
  } else {
    for (int al = warp; al < FP_LINES; al += NW) {
      const int a = min(a0 + al, A.na - 1);
      const char *line = (const char *)src + (size_t)a * A.in_pitch;
      for (int e = lane; e < span; e += 32) {
        int b = min(max(b_first + e, 0), A.nb - 1);
        if (A.ups) b = min(b >> 1, A.in_nb - 1);             // matrix2d.js:129 floor(j * 0.5)
        tile[e * FP_PITCH + al] = fir_src_at(line, b, A.src_kind);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// mode 0 (pass A): every level reads the same base -> one tile with the largest halo.
// mode 1 (pass B): level s reads its own T_s -> one tile per level, the next one in flight (cp.async)
//                  while the current one is consumed.
// window prefetch (NO) + tap padding (up to 3) + weight prefetch group read past the last real sample
#define FIR_SLACK(NO) ((NO) + 4)

#ifndef FIR_NO8_BLOCKS
#define FIR_NO8_BLOCKS 3
#endif
template <int NW, int NO, int MODE>
__global__ void __launch_bounds__(32 * NW, (NO == 16 || NW == 4) ? 2 : FIR_NO8_BLOCKS)
fir_pass_kernel(const double *__restrict__ weights, const FirArgs A)
{
  extern __shared__ double smem[];
  double *wsm = smem;                                       // per level: taps zero-padded to a multiple of 4
  double *tile0 = smem + A.wtotal;                          // [span][FP_PITCH]
  double *tile1 = tile0 + (size_t)(NW * NO + 2 * A.rmax + FIR_SLACK(NO)) * FP_PITCH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a0 = blockIdx.y * FP_LINES;
  const int b_tile = blockIdx.x * (NW * NO);
  constexpr bool per_level = MODE == 1;

  {
    const int halo = per_level ? A.radius[0] : A.rmax;
    fir_stage<NW>(A, per_level ? (const void *)A.T[0] : A.src, tile0, a0, b_tile - halo, NW * NO + 2 * halo + FIR_SLACK(NO), lane, warp);
  }
  for (int li = 0, wo = 0; li < A.nlev; li++) {
    const int n = 2 * A.radius[li] + 1, npad = ((n + 3) & ~3) + 4;     // + one all-zero group: the weight prefetch
    for (int e = threadIdx.x; e < npad; e += 32 * NW) wsm[wo + e] = e < n ? __ldg(weights + A.woff[li] + e) : 0.0;
    wo += npad;
  }

  const int a = a0 + lane;
  const int b0 = b_tile + warp * NO;
  const bool active = a < A.na && b0 < A.nb;
  const bool full = b0 + NO <= A.nb;                       // warp-uniform

  double prev[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) prev[k] = 0.0;
  if (per_level && A.seed_is_level0 && active) {             // octaves >= 1: level 0 is the unblurred seed
#pragma unroll
    for (int k = 0; k < NO; k++) prev[k] = A.oct.seed64[(size_t)min(b0 + k, A.nb - 1) * A.oct.w + a];
  }

  int wo = 0;
  for (int li = 0; li < A.nlev; li++) {
    const int R = A.radius[li], npad = (2 * R + 4) & ~3;
    double *cur = (per_level && (li & 1)) ? tile1 : tile0;
    if (per_level && li + 1 < A.nlev) {
      const int halo = A.radius[li + 1];
      fir_stage<NW>(A, A.T[li + 1], (li & 1) ? tile0 : tile1, a0, b_tile - halo, NW * NO + 2 * halo + FIR_SLACK(NO), lane, warp);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (per_level || li == 0) __syncthreads();               // the tile (and, first time, the taps) visible to the CTA
    if (active) {
      double acc[NO];
      fir_window<NO>(wsm + wo, npad, cur + (warp * NO + (per_level ? 0 : A.rmax - R)) * FP_PITCH + lane, acc);
      if (!per_level) {
        double *out = A.T[li] + (size_t)b0 * A.t_pitch + a;
        const size_t tp = A.t_pitch;
        if (full) {
#pragma unroll
          for (int k = 0; k < NO; k++) { *out = acc[k]; out += tp; }       // running pointer: two adds per store
        } else {
#pragma unroll
          for (int k = 0; k < NO; k++)
            if (b0 + k < A.nb) out[(size_t)k * tp] = acc[k];
        }
      } else {
        // G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) from the unrounded accumulators, seed of the next octave
        const int s = A.level[li];
        const size_t pitch = (size_t)A.oct.pitch;
        const size_t o = (size_t)b0 * pitch + a;
        float *gp = A.oct.gauss[s] + o;
        float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + o;
        if (full) {
          if (A.keep_gauss) {
#pragma unroll
            for (int k = 0; k < NO; k++) { *gp = (float)acc[k]; gp += pitch; }
          }
          if (s > 0) {
#pragma unroll
            for (int k = 0; k < NO; k++) { *dp = (float)(prev[k] - acc[k]); dp += pitch; }
          }
        } else {
#pragma unroll
          for (int k = 0; k < NO; k++) {
            if (b0 + k < A.nb) {
              if (A.keep_gauss) gp[(size_t)k * pitch] = (float)acc[k];
              if (s > 0) dp[(size_t)k * pitch] = (float)(prev[k] - acc[k]);
            }
          }
        }
        if (A.has_next && s == A.spo && (a & 1) == 0) {      // matrix2d.js:129 in[2a][2b]: even rows, even columns
#pragma unroll
          for (int k = 0; k < NO; k += 2) {
            const int nr = ((b0 + k) >> 1) + A.oct.seed_off;     // row of the next octave (strip-local)
            if (b0 + k < A.nb && nr >= 0 && nr < A.next.h) {
              A.next.seed64[(size_t)nr * A.next.w + (a >> 1)] = acc[k];
              A.next.gauss[0][(size_t)nr * A.next.pitch + (a >> 1)] = (float)acc[k];
            }
          }
        }
#pragma unroll
        for (int k = 0; k < NO; k++) prev[k] = acc[k];
      }
    }
    wo += npad + 4;
    if (per_level) __syncthreads();                          // `cur` is overwritten by the copy issued next iteration
  }
}

// ------------------------------------------------------------------ host side
// Tile shapes: big octaves use 16 outputs per thread (fewer loads and less per-level overhead per DFMA);
// small octaves use 4-warp CTAs of 8 outputs per thread so that the few pixels still spread over all SMs.
struct FirShape { int nw, no; };
static FirShape fir_shape(int na, int nb)
{
  const long long px = (long long)na * nb;
  static const char *no8 = getenv("SIFT_B200_FIR_NO8");
  if (px >= (1 << 20)) return (no8 && no8[0] == '1') ? FirShape{ 8, 8 } : FirShape{ 8, 16 };
  if (px >= (1 << 18)) return { 8, 8 };
  return { 4, 8 };
}

static size_t fir_smem_bytes(int wtotal_padded, int rmax, FirShape sh, int tiles)
{
  const int span = sh.nw * sh.no + 2 * rmax + FIR_SLACK(sh.no);
  return ((size_t)wtotal_padded + (size_t)tiles * span * FP_PITCH) * sizeof(double);
}

size_t sep_t_pitch(int nb_stored) { return (size_t)((nb_stored + 3) & ~3); }

// Doubles of intermediate storage the two passes need for one octave (all blurred levels).
size_t sep_t_elems(int w, int trows, int n_levels) { return (size_t)n_levels * w * sep_t_pitch(trows); }

static int padded_taps(int radius) { return ((2 * radius + 4) & ~3) + 4; }

bool sep_supported(const LevelPlan *plans, int first_level, int nlev, int w, int h)
{
  int rmax = 0, wtotal = 0;
  for (int s = first_level; s < nlev; s++) { rmax = max(rmax, plans[s].radius); wtotal += padded_taps(plans[s].radius); }
  return fir_smem_bytes(wtotal, rmax, fir_shape(w, h), 2) <= 220 * 1024;
}

static void fill_levels(FirArgs &A, const LevelPlan *plans, int first_level, int nlev, double *tbase, size_t plane)
{
  A.nlev = nlev - first_level;
  A.rmax = 0; A.wtotal = 0;
  for (int i = 0; i < A.nlev; i++) {
    const LevelPlan &p = plans[first_level + i];
    A.level[i] = first_level + i;
    A.radius[i] = p.radius;
    A.woff[i] = p.woff;
    A.T[i] = tbase + (size_t)i * plane;
    A.rmax = max(A.rmax, p.radius);
    A.wtotal += padded_taps(p.radius);        // zero-padded to a multiple of 4 taps + one zero group
  }
}

template <int NW, int NO, int MODE>
static void fir_launch(cudaStream_t st, const double *d_weights, const FirArgs &A, size_t smem)
{
  dim3 grid((A.nb + NW * NO - 1) / (NW * NO), (A.na + FP_LINES - 1) / FP_LINES);
  cudaFuncSetAttribute(fir_pass_kernel<NW, NO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fir_pass_kernel<NW, NO, MODE><<<grid, 32 * NW, smem, st>>>(d_weights, A);
}

template <int MODE>
static void fir_dispatch(cudaStream_t st, const double *d_weights, const FirArgs &A, FirShape sh)
{
  const size_t smem = fir_smem_bytes(A.wtotal, A.rmax, sh, MODE == 1 ? 2 : 1);
  if (sh.no == 16) fir_launch<8, 16, MODE>(st, d_weights, A, smem);
  else if (sh.nw == 8) fir_launch<8, 8, MODE>(st, d_weights, A, smem);
  else fir_launch<4, 8, MODE>(st, d_weights, A, smem);
}

// pass A: base (seed64 of octaves >= 1, or the source image doubled along x for a generic octave 0) -> T^T
void launch_sep_pass_a(cudaStream_t st, const void *src, int dtype, size_t src_pitch_bytes, int src_w, int upsample,
                       int w, int hrows, int h, const double *d_weights, const LevelPlan *plans, int first_level,
                       int nlev, double *tbase)
{
  FirArgs A;
  memset(&A, 0, sizeof A);
  A.src = src; A.src_kind = dtype; A.in_pitch = src_pitch_bytes; A.in_nb = src_w;
  A.na = hrows; A.nb = w; A.ups = upsample;
  A.mode = 0;
  A.t_pitch = sep_t_pitch(hrows);
  fill_levels(A, plans, first_level, nlev, tbase, (size_t)w * A.t_pitch);
  fir_dispatch<0>(st, d_weights, A, fir_shape(w, h));
}

// pass B: T^T -> Gaussian / DoG levels (+ next seed).  `upsample`: T holds hrows = h/2 source rows (generic octave 0).
void launch_sep_pass_b(cudaStream_t st, int upsample, const OctaveDev &oct, const double *d_weights,
                       const LevelPlan *plans, int first_level, double *tbase, int hrows, const OctaveDev *next,
                       int spo, int keep_gauss)
{
  FirArgs A;
  memset(&A, 0, sizeof A);
  A.src = tbase; A.src_kind = SIFT_F64;
  A.t_pitch = sep_t_pitch(hrows);
  A.in_pitch = A.t_pitch * sizeof(double); A.in_nb = hrows;
  A.na = oct.w; A.nb = oct.h; A.ups = upsample;
  A.mode = 1;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.seed_is_level0 = first_level > 0;
  fill_levels(A, plans, first_level, oct.nlev, tbase, (size_t)oct.w * A.t_pitch);
  fir_dispatch<1>(st, d_weights, A, fir_shape(oct.w, oct.h));
}
