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
  size_t t_pitch;        // doubles per T^T line (multiple of 16: lines start on 128-byte boundaries)
  const void *tmaps;     // pass B: one CUtensorMap per level over its T^T plane (box = level's span x 32 lines), or null
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, seed_is_level0;
};

// a[k] = sum_{j<npad} w[j] * v[k + j], k < NO; v[p] = base[p * FP_PITCH].  The taps are zero-padded to a
// multiple of 4 (npad) and 16-byte aligned: the loop runs whole groups of 4 taps, the weights of the next
// group are fetched (2 x LDS.128, broadcast) while the current group is consumed, the window of NO samples
// rotates through registers with compile-time indices, and the accumulators never change registers.
// Positions up to npad + NO - 1 are read.
template <int NO, int STRIDE>
__device__ __forceinline__ void fir_window(const double *__restrict__ w, const int npad,
                                           const double *__restrict__ base, double (&a)[NO])
{
  double vw[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) { vw[k] = base[k * STRIDE]; a[k] = 0.0; }
  const double *nxt = base + NO * STRIDE;
  double2 wa = *reinterpret_cast<const double2 *>(w), wb = *reinterpret_cast<const double2 *>(w + 2);
#define FIR_GROUP(G)                                                                              \
  {                                                                                               \
    const double c[4] = { wa.x, wa.y, wb.x, wb.y };                                               \
    wa = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 4);                                 \
    wb = *reinterpret_cast<const double2 *>(w + j + 4 * (G) + 6);                                 \
    _Pragma("unroll") for (int u = 0; u < 4; u++) {                                               \
      _Pragma("unroll") for (int k = 0; k < NO; k++) a[k] = fma(c[u], vw[(k + 4 * (G) + u) & (NO - 1)], a[k]); \
      vw[(4 * (G) + u) & (NO - 1)] = nxt[(j + 4 * (G) + u) * STRIDE];                              \
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
  const bool active = b0 < A.nb;                           // warp-uniform
  const bool line_ok = a < A.na;                           // lanes past the last line compute on its replica, store nothing
  const bool full = b0 + NO <= A.nb;                       // warp-uniform

  double prev[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) prev[k] = 0.0;
  if (per_level && A.seed_is_level0 && active && line_ok) {  // octaves >= 1: level 0 is the unblurred seed
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
      fir_window<NO, FP_PITCH>(wsm + wo, npad, cur + (warp * NO + (per_level ? 0 : A.rmax - R)) * FP_PITCH + lane, acc);
      if (!per_level) {
        double *out = A.T[li] + (size_t)b0 * A.t_pitch + a;
        const size_t tp = A.t_pitch;
        if (full && line_ok) {
#pragma unroll
          for (int k = 0; k < NO; k++) { *out = acc[k]; out += tp; }       // running pointer: two adds per store
        } else if (line_ok) {
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
        if (full && line_ok) {
          if (A.keep_gauss) {
#pragma unroll
            for (int k = 0; k < NO; k++) { *gp = (float)acc[k]; gp += pitch; }
          }
          if (s > 0) {
#pragma unroll
            for (int k = 0; k < NO; k++) { *dp = (float)(prev[k] - acc[k]); dp += pitch; }
          }
        } else if (line_ok) {
#pragma unroll
          for (int k = 0; k < NO; k++) {
            if (b0 + k < A.nb) {
              if (A.keep_gauss) gp[(size_t)k * pitch] = (float)acc[k];
              if (s > 0) dp[(size_t)k * pitch] = (float)(prev[k] - acc[k]);
            }
          }
        }
        if (A.has_next && s == A.spo && (a & 1) == 0 && line_ok) {      // matrix2d.js:129 in[2a][2b]: even rows, even columns
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

// ---- pass B with TMA tile loads ---------------------------------------------------------------------------
// The tile a CTA needs for level s -- 32 lines x (outputs + 2 R_s) samples of T_s^T -- is a plain box (the TMA unit
// zero-fills what lies outside a line, border tiles then replicate the edge sample): ONE cp.async.bulk.tensor.2d per level, issued by one thread a level ahead and completed on
// an mbarrier, replaces ~25 cp.async + address arithmetic per thread per level.  The box lands line-major
// ([line][sample], dense): a lane walks its own line with unit stride, the row length is chosen = 2 (mod 4)
// doubles so that the 32 lanes of a load fall on 8 bank pairs (2-way conflicts on one load per 16 DFMAs).
#include <cuda.h>

__device__ __forceinline__ unsigned fir_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// The box starts (radius rounded up to even) samples before the tile, so that its first sample is 16-byte aligned.
__host__ __device__ __forceinline__ int fir_box_halo(int radius) { return (radius + 1) & ~1; }

__host__ __device__ __forceinline__ int fir_box_span(int outputs, int radius, int no)
{
  const int span = outputs + 2 * fir_box_halo(radius) + FIR_SLACK(no);
  return span + ((2 - (span & 3)) & 3);                     // smallest s >= span with s % 4 == 2
}

template <int NW, int NO>
__global__ void __launch_bounds__(32 * NW, 2)
fir_pass_b_tma_kernel(const double *__restrict__ weights, const FirArgs A)
{
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const int span_max = fir_box_span(NW * NO, A.rmax, NO);
  const int buf_doubles = (FP_LINES * span_max + 15) & ~15;  // 128-byte multiples: TMA destinations
  double *buf0 = smem, *buf1 = smem + buf_doubles;
  double *wsm = smem + 2 * buf_doubles;                     // per level: taps zero-padded to a multiple of 4 (+ one zero group)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a0 = blockIdx.y * FP_LINES;
  const int b_tile = blockIdx.x * (NW * NO);
  const CUtensorMap *maps = (const CUtensorMap *)A.tmaps;

  auto issue = [&](int li) {                                // one thread: arm the barrier, one box
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last read / patched through the generic proxy
    const int R = A.radius[li];
    const unsigned bytes = (unsigned)(FP_LINES * fir_box_span(NW * NO, R, NO) * sizeof(double));
    const unsigned b = fir_smem_u32(&bar[li & 1]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(fir_smem_u32((li & 1) ? buf1 : buf0)), "l"(maps + li), "r"(b), "r"(b_tile - fir_box_halo(R)), "r"(a0)
        : "memory");
  };

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fir_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fir_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int li = 0, wo = 0; li < A.nlev; li++) {
    const int n = 2 * A.radius[li] + 1, npad = ((n + 3) & ~3) + 4;
    for (int e = threadIdx.x; e < npad; e += 32 * NW) wsm[wo + e] = e < n ? __ldg(weights + A.woff[li] + e) : 0.0;
    wo += npad;
  }
  __syncthreads();
  if (threadIdx.x == 0) issue(0);

  const int a = a0 + lane;
  const int b0 = b_tile + warp * NO;
  const bool active = b0 < A.nb;                           // warp-uniform
  const bool line_ok = a < A.na;                           // lines past the end are zero-filled by the TMA unit
  const bool full = b0 + NO <= A.nb;

  double prev[NO];
#pragma unroll
  for (int k = 0; k < NO; k++) prev[k] = 0.0;
  if (A.seed_is_level0 && active && line_ok) {               // octaves >= 1: level 0 is the unblurred seed
#pragma unroll
    for (int k = 0; k < NO; k++) prev[k] = A.oct.seed64[(size_t)min(b0 + k, A.nb - 1) * A.oct.w + a];
  }

  int wo = 0;
  for (int li = 0; li < A.nlev; li++) {
    const int R = A.radius[li], npad = (2 * R + 4) & ~3;
    // the other buffer was released by the barrier that ended level li-1: fill it with level li+1
    if (threadIdx.x == 0 && li + 1 < A.nlev) issue(li + 1);
    {
      const unsigned parity = (li >> 1) & 1;
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(fir_smem_u32(&bar[li & 1])), "r"(parity) : "memory");
      }
    }
    const int pitch_l = fir_box_span(NW * NO, R, NO);
    {
      // clamp-to-edge (sift.js:118-119): the TMA unit zero-filled what lies before sample 0 / after sample nb-1;
      // border tiles replicate the edge sample over those entries (each thread: its lane's line, entries warp, warp+NW ..)
      double *cur = (li & 1) ? buf1 : buf0;
      const int b_first = b_tile - fir_box_halo(R);
      const bool left = b_first < 0, right = b_first + pitch_l > A.nb;        // CTA-uniform
      if (left || right) {
        double *ln = cur + lane * pitch_l;
        if (left) {
          const double v = ln[-b_first];
          for (int e = warp; e < -b_first; e += NW) ln[e] = v;
        }
        if (right) {
          const int last = A.nb - 1 - b_first;                                  // >= 0: the tile starts inside the line
          const double v = ln[last];
          for (int e = last + 1 + warp; e < pitch_l; e += NW) ln[e] = v;
        }
        // these generic-proxy stores land in a buffer the TMA unit (async proxy) refills two levels later: order them
        // before that write (the CTA barrier at the end of the level carries the fence to the issuing thread)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
    }
    if (active) {
      const double *cur = (li & 1) ? buf1 : buf0;
      double acc[NO];
      fir_window<NO, 1>(wsm + wo, npad, cur + lane * pitch_l + warp * NO + (fir_box_halo(R) - R), acc);
      // G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) from the unrounded accumulators, seed of the next octave
      const int s = A.level[li];
      const size_t pitch = (size_t)A.oct.pitch;
      const size_t o = (size_t)b0 * pitch + a;
      float *gp = A.oct.gauss[s] + o;
      float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + o;
      if (full && line_ok) {
        if (A.keep_gauss) {
#pragma unroll
          for (int k = 0; k < NO; k++) { *gp = (float)acc[k]; gp += pitch; }
        }
        if (s > 0) {
#pragma unroll
          for (int k = 0; k < NO; k++) { *dp = (float)(prev[k] - acc[k]); dp += pitch; }
        }
      } else if (line_ok) {
#pragma unroll
        for (int k = 0; k < NO; k++) {
          if (b0 + k < A.nb) {
            if (A.keep_gauss) gp[(size_t)k * pitch] = (float)acc[k];
            if (s > 0) dp[(size_t)k * pitch] = (float)(prev[k] - acc[k]);
          }
        }
      }
      if (A.has_next && s == A.spo && (a & 1) == 0 && line_ok) {   // matrix2d.js:129 in[2a][2b]: even rows, even columns
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
    wo += npad + 4;
    __syncthreads();                                         // this buffer may be refilled (level li+2)
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

// T^T line layout: nb samples per line, lines on 128-byte boundaries (pass A's 32 x 8 B stores stay sector-aligned;
// TMA needs 16-byte multiples).  No clamp pads in global memory: the TMA unit zero-fills what lies outside a
// line and the pass-B kernel replicates the edge sample inside the staged tile (border tiles only).
static size_t sep_t_pitch(int nb_stored) { return (size_t)((nb_stored + 15) & ~15); }

static int level_rmax(const LevelPlan *plans, int first_level, int nlev)
{
  int rmax = 0;
  for (int s = first_level; s < nlev; s++) rmax = max(rmax, plans[s].radius);
  return rmax;
}

// Doubles of intermediate storage the two passes need for one octave (all blurred levels).
size_t sep_t_elems(int w, int h, int trows, const LevelPlan *plans, int first_level, int nlev)
{
  (void)h; (void)plans;
  return (size_t)(nlev - first_level) * w * sep_t_pitch(trows);
}

static int padded_taps(int radius) { return ((2 * radius + 4) & ~3) + 4; }

bool sep_supported(const LevelPlan *plans, int first_level, int nlev, int w, int h)
{
  int rmax = 0, wtotal = 0;
  for (int s = first_level; s < nlev; s++) { rmax = max(rmax, plans[s].radius); wtotal += padded_taps(plans[s].radius); }
  return fir_smem_bytes(wtotal, rmax, fir_shape(w, h), 2) <= 220 * 1024;
}

static void fill_levels(FirArgs &A, const LevelPlan *plans, int first_level, int nlev, double *tbase, int w, int h, int trows)
{
  (void)h;
  A.t_pitch = sep_t_pitch(trows);
  const size_t plane = (size_t)w * A.t_pitch;
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
  fill_levels(A, plans, first_level, nlev, tbase, w, h, hrows);
  fir_dispatch<0>(st, d_weights, A, fir_shape(w, h));
}

// One CUtensorMap per blurred level over its T^T plane (fp64, [w lines][pitch]), box = level's span x 32 lines.
// Returns the number of maps written (0: TMA path not usable for this octave).
typedef CUresult (*FirEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

size_t sep_tma_map_bytes(int n_levels) { return (size_t)n_levels * sizeof(CUtensorMap); }

int sep_tma_build_maps(const LevelPlan *plans, int first_level, int nlev, double *tbase, int w, int h, int trows, void *h_maps)
{
  static FirEncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return 0;
    encode = (FirEncodeTiledFn)fn;
  }
  const FirShape sh = fir_shape(w, h);
  const int rmax = level_rmax(plans, first_level, nlev);
  const size_t pitch = sep_t_pitch(trows);
  if (fir_box_span(sh.nw * sh.no, rmax, sh.no) > 256) return 0;           // TMA boxes are at most 256 elements wide
  CUtensorMap *maps = (CUtensorMap *)h_maps;
  for (int i = 0; i < nlev - first_level; i++) {
    const cuuint64_t gdim[2] = { (cuuint64_t)pitch, (cuuint64_t)w };
    const cuuint64_t gstride[1] = { (cuuint64_t)pitch * sizeof(double) };
    const cuuint32_t box[2] = { (cuuint32_t)fir_box_span(sh.nw * sh.no, plans[first_level + i].radius, sh.no), FP_LINES };
    const cuuint32_t estride[2] = { 1, 1 };
    // 64-bit elements: the driver has no FLOAT64 tile type restriction issue -- use FLOAT64
    if (encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)(tbase + (size_t)i * w * pitch), gdim, gstride, box,
               estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 0;
  }
  return nlev - first_level;
}

// Whether pass B of an octave will run the TMA kernel.
bool sep_pass_b_uses_tma(const void *d_tmaps, int upsample)
{
  static const bool no_tma = getenv("SIFT_B200_NO_TMA") != nullptr || getenv("SIFT_B200_NO_TMA_BLUR") != nullptr;
  return d_tmaps && !upsample && !no_tma;
}

template <int NW, int NO>
static void fir_launch_tma(cudaStream_t st, const double *d_weights, const FirArgs &A)
{
  const int span_max = fir_box_span(NW * NO, A.rmax, NO);
  const size_t smem = ((size_t)2 * ((FP_LINES * span_max + 15) & ~15) + A.wtotal) * sizeof(double);
  dim3 grid((A.nb + NW * NO - 1) / (NW * NO), (A.na + FP_LINES - 1) / FP_LINES);
  cudaFuncSetAttribute(fir_pass_b_tma_kernel<NW, NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fir_pass_b_tma_kernel<NW, NO><<<grid, 32 * NW, smem, st>>>(d_weights, A);
}

// pass B: T^T -> Gaussian / DoG levels (+ next seed).  `upsample`: T holds hrows = h/2 source rows (generic octave 0).
// d_tmaps: device array of this octave's tensor maps (sep_tma_build_maps), or null for the cp.async path.
void launch_sep_pass_b(cudaStream_t st, int upsample, const OctaveDev &oct, const double *d_weights,
                       const LevelPlan *plans, int first_level, double *tbase, int hrows, const OctaveDev *next,
                       int spo, int keep_gauss, const void *d_tmaps)
{
  FirArgs A;
  memset(&A, 0, sizeof A);
  A.src = tbase; A.src_kind = SIFT_F64;
  A.in_nb = hrows;
  A.na = oct.w; A.nb = oct.h; A.ups = upsample;
  A.mode = 1;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.seed_is_level0 = first_level > 0;
  fill_levels(A, plans, first_level, oct.nlev, tbase, oct.w, oct.h, hrows);
  A.in_pitch = A.t_pitch * sizeof(double);
  const FirShape sh = fir_shape(oct.w, oct.h);
  if (sep_pass_b_uses_tma(d_tmaps, upsample) && fir_box_span(sh.nw * sh.no, A.rmax, sh.no) <= 256) {
    A.tmaps = d_tmaps;
    if (sh.no == 16) fir_launch_tma<8, 16>(st, d_weights, A);
    else if (sh.nw == 8) fir_launch_tma<8, 8>(st, d_weights, A);
    else fir_launch_tma<4, 8>(st, d_weights, A);
    return;
  }
  fir_dispatch<1>(st, d_weights, A, sh);
}
