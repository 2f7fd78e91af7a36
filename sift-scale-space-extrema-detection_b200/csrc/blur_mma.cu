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

// ---- TMA (cp.async.bulk.tensor) helpers: one elected thread issues a 2-D box, everybody waits on the mbarrier ----
#include <cuda.h>
__device__ __forceinline__ unsigned mm_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mm_bar_init(unsigned long long *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mm_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mm_tma_load_2d(void *dst_smem, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mm_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(mm_smem_u32(dst_smem)), "l"(map), "r"(mm_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mm_bar_wait(unsigned long long *bar, unsigned parity)
{
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mm_smem_u32(bar)), "r"(parity) : "memory");
}
typedef CUresult (*MmEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// 2-D tiled map over a pitched plane; false when the driver entry point is missing or the layout is not TMA-able
static bool mm_encode_2d(CUtensorMap *m, CUtensorMapDataType ty, size_t esz, const void *base, size_t pitch_bytes, int w, int h,
                         int box_w, int box_h)
{
  static MmEncodeTiledFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
        q == cudaDriverEntryPointSuccess)
      encode = (MmEncodeTiledFn)fn;
  }
  static const bool no_tma = getenv("SIFT_B200_NO_TMA") != nullptr || getenv("SIFT_B200_NO_TMA_BLUR") != nullptr ||
                             getenv("SIFT_B200_MMA_NO_TMA") != nullptr;
  if (!encode || no_tma) return false;
  if ((pitch_bytes % 16) != 0 || ((uintptr_t)base % 16) != 0 || box_w > 256 || box_h > 256 || ((size_t)box_w * esz) % 16 != 0) return false;
  const cuuint64_t gdim[2] = { (cuuint64_t)w, (cuuint64_t)h };
  const cuuint64_t gstride[1] = { (cuuint64_t)pitch_bytes };
  const cuuint32_t box[2] = { (cuuint32_t)box_w, (cuuint32_t)box_h };
  const cuuint32_t estride[2] = { 1, 1 };
  return encode(m, ty, 2, const_cast<void *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Grid of a DMMA kernel over n_tiles tiles: one CTA per tile, or -- SIFT_B200_MMA_HALF_SM=1 -- one CTA per SM walking
// over the tiles, so that a kernel holds half of every SM (its CTAs take half the registers) and the blur kernel of
// another frame in flight runs beside it instead of after it.
static int mm_grid(int n_tiles)
{
  static int n_sm = 0;
  static const bool half_sm = getenv("SIFT_B200_MMA_HALF_SM") != nullptr;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  return (half_sm && n_tiles > n_sm) ? n_sm : n_tiles;
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
// 1: the two groups of four warps of a CTA take turns on the DMMA pipe (measured 0.124 against 0.131 ms per 1080p frame);
// 0: free-running warps; 2: hand-over after the H pass (0.139).  M0_HALVES = 2: 16-warp CTAs, ring of four groups (0.141).
#ifndef M0_PINGPONG
#define M0_PINGPONG 1
#endif
#define M0_MAXD 5                        // chunks of 4 samples per 8 x K band: ceil((n + 3) / 4), n <= 17
#define M0_S_DOUBLES (M0_SROWS * M0_SPITCH + 16)       // 2928: keeps everything behind it on 128-byte boundaries
#define M0_T_DOUBLES (8 * M0_TW_DOUBLES)
#define M0_RAW_DOUBLES (M0_SCOLS * M0_SCOLS / 2)       // TMA destination: 48 x 48 raw samples of at most 4 bytes
#define M0_SMEM_DOUBLES(nlev) (M0_S_DOUBLES + M0_T_DOUBLES + (nlev) * M0_MAXD * 32 + M0_RAW_DOUBLES)

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
  int use_tma;                           // interior source tiles arrive as one TMA box (u8 / f32 sources with a TMA-able pitch)
  int tile_y0;                           // first tile row of this launch (mosaic strips run octave 0 band by band behind their upload)
  int tiles_x, n_tiles;                  // tiles per row / tiles of this launch
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

__device__ __forceinline__ float m0_cvt(double x) { return (float)x; }

// Both passes of one level for ONE WARP, no CTA barrier: the warp blurs the rows its own vertical window needs
// (4 D + 12 rows x its 16 columns; 12 % more multiply-adds than sharing the rows across the CTA) into its private
// 4 KB of Ts and consumes them itself, so the eight warps of a CTA drift through H pass / V pass / epilogue
// independently and cover each other's latencies.  D = chunks of the band (CTA-uniform), acc = the warp's 32 x 16 outputs.
// Ts layout: [32 rows][16 columns], column index XOR-swizzled by the row (8 (r & 1) + 2 ((r >> 1) & 1)): the
// H-pass 16-byte stores (lanes along rows) and the V-pass 8-byte loads (lanes along 4 rows x 8 columns) are both
// conflict-free without padding.
template <int D>
__device__ __forceinline__ void mma0_level(const double *__restrict__ S, double *__restrict__ Tw,
                                           const double *__restrict__ wfs, const int clo0, const int warp,
                                           const int lane, double (&acc)[4][2][2], const int stage, const int n_stages)
{
  constexpr int MBH = (4 * D + 12 + 7) / 8;             // 8-row blocks of H-pass rows: 3 (D <= 3) or 4
  const int g = lane >> 2, t = lane & 3;
  const int wy = warp >> 2, wx = warp & 3;
  double wf[D];
#pragma unroll
  for (int d = 0; d < D; d++) wf[d] = wfs[d * 32 + lane];
#if M0_PINGPONG
  // The two halves of the CTA (warps 0-3 / 4-7: one warp of each per SM sub-partition) take turns on the DMMA pipe:
  // a half blurs a level while the other converts and stores the level it has just finished.  Warps that share a
  // pipe fairly finish their DMMA phases together and then idle it together; the hand-over keeps one of them on it.
  asm volatile("bar.sync %0, 256;" ::"r"(1 + stage) : "memory");
#endif
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
#if M0_PINGPONG == 2
    asm volatile("bar.arrive %0, 256;" ::"r"(1 + (stage + 1 == n_stages ? 0 : stage + 1)) : "memory");     // early hand-over
#endif
    __syncwarp();                                       // the previous level's V pass has read Tw
    const int sw = 8 * (g & 1) + 2 * ((g >> 1) & 1);
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
    // N block nb takes the columns {0,1, 4,5, 8,9, 12,13} + 2 nb of the warp's 16: a lane's two blocks then hold FOUR
    // neighbouring columns of one row (one 16-byte store per row and plane instead of two 8-byte ones)
    const int sw = 8 * (t & 1) + 2 * (t >> 1);
    const int col = 4 * (g >> 1) + (g & 1);
    const double *tp0 = Tw + t * 16 + (col ^ sw), *tp1 = Tw + t * 16 + ((col + 2) ^ sw);
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
#if M0_PINGPONG == 1
  asm volatile("bar.arrive %0, 256;" ::"r"(1 + (stage + 1 == n_stages ? 0 : stage + 1)) : "memory");     // the next group may blur its level
#endif
}

// HV = tiles per CTA: 1 (8 warps, two CTAs per SM) or 2 (16 warps, one CTA per SM; each group of 8 warps owns a tile
// with its own S / Ts / raw regions, the fragments are shared).  With M0_PINGPONG the 2 HV groups of four warps (one
// warp per SM sub-partition each) take turns on the DMMA pipe in a ring.
template <int HV>
__global__ void __launch_bounds__(M0_THREADS * HV, 3 - HV)
oct0_mma_kernel(const __grid_constant__ Mma0Args A, const __grid_constant__ CUtensorMap src_map)
{
  extern __shared__ __align__(128) double smem[];
  __shared__ int lvR[SIFT_MAX_LEVELS];
  __shared__ __align__(8) unsigned long long src_bars[HV];
  const int half = threadIdx.x / M0_THREADS;            // which tile of the CTA
  const int tid = threadIdx.x % M0_THREADS, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int stage = (int)threadIdx.x >> 7, n_stages = 2 * HV;          // groups of four warps
  const int half_doubles = M0_S_DOUBLES + M0_T_DOUBLES + M0_RAW_DOUBLES;
  double *S = smem + half * half_doubles;               // [56][52] source tile v / 255 (rows / columns >= 48: zero)
  double *Tw = S + M0_S_DOUBLES + warp * M0_TW_DOUBLES; // this warp's horizontally blurred rows
  unsigned char *raw = reinterpret_cast<unsigned char *>(S + M0_S_DOUBLES + M0_T_DOUBLES);
  double *Wf = smem + HV * half_doubles;                // [nlev][M0_MAXD][32] band fragments
  unsigned long long &src_bar = src_bars[half];
  if (tid == 0) mm_bar_init(&src_bar, 1);
  if ((int)threadIdx.x < A.nlev) lvR[threadIdx.x] = A.radius[threadIdx.x];
  for (int e = threadIdx.x; e < A.nlev * M0_MAXD * 32; e += M0_THREADS * HV) Wf[e] = __ldg(A.wfrag + e);
  // everything of S beyond the 48 x 48 samples is read against zero weights only: keep it finite (written once)
  for (int e = tid; e < M0_SCOLS * (M0_SPITCH - M0_SCOLS); e += M0_THREADS)
    S[(e >> 2) * M0_SPITCH + M0_SCOLS + (e & 3)] = 0.0;
  for (int e = M0_SCOLS * M0_SPITCH + tid; e < M0_S_DOUBLES; e += M0_THREADS) S[e] = 0.0;
  __syncthreads();                                      // barrier initialised, level table and fragments staged
#if M0_PINGPONG
  if (stage == n_stages - 1) asm volatile("bar.arrive 1, 256;" ::: "memory");  // the first group starts
#endif

  // The CTA walks over the tiles blockIdx.x, blockIdx.x + gridDim.x, ... of this launch (one tile per CTA unless the
  // host caps the grid: SIFT_B200_MMA_HALF_SM launches one CTA per SM so that another frame's kernel shares the SM).
  unsigned tma_uses = 0;
  for (int tile0 = blockIdx.x * HV; tile0 < A.n_tiles; tile0 += gridDim.x * HV) {
  const bool live = tile0 + half < A.n_tiles;           // an odd tile count leaves the last CTA's second group without a tile:
  const int tile = live ? tile0 + half : A.n_tiles - 1; // it blurs a copy of the last one and stores nothing
  const int a_tile = (tile % A.tiles_x) * M0_SW, b_tile = (tile / A.tiles_x + A.tile_y0) * M0_SH - A.row_shift;

  // interior tiles: the 48 x 48 window of u8 / f32 source samples arrives as ONE TMA box in `raw` (the unit would
  // zero-fill outside the image where the reference clamps: border tiles are gathered with clamped loads below)
  const bool by_tma = A.use_tma && a_tile >= M0_HALO && a_tile + M0_SW + M0_HALO <= A.src_w && b_tile >= M0_HALO &&
                      b_tile + M0_SH + M0_HALO <= A.src_h;                     // CTA-uniform
  if (by_tma && tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // `raw` was read through the generic proxy (previous tile)
    // a box starts on a 16-byte boundary of its row (TMA requirement): u8 tiles take 16 columns of left halo
    // (box 64 x 48, the window starts at byte 8 of each row), f32 tiles the 8 they need (box 48 x 48)
    if (A.dtype == SIFT_U8) mm_tma_load_2d(raw, &src_map, &src_bar, a_tile - 16, b_tile - M0_HALO, 64u * M0_SCOLS);
    else mm_tma_load_2d(raw, &src_map, &src_bar, a_tile - M0_HALO, b_tile - M0_HALO, (unsigned)(M0_SCOLS * M0_SCOLS * 4));
  }
  if (by_tma) {
    mm_bar_wait(&src_bar, tma_uses & 1);
    tma_uses++;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int e = tid + i * M0_THREADS;
      const int rr = e / M0_SCOLS, cc = e - rr * M0_SCOLS;
      S[rr * M0_SPITCH + cc] = A.dtype == SIFT_U8 ? mma0_u8(raw[rr * 64 + 8 + cc]) : (double)reinterpret_cast<const float *>(raw)[e];
    }
  } else {
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
  }
  __syncthreads();                                      // source tile staged: the last CTA barrier of the tile

  // V-pass ownership of this warp: output rows 32 wy .. +31 (source rows 16 wy .. +15), columns 16 wx .. +15
  const int wy = warp >> 2, wx = warp & 3;
  const int y0 = 2 * (b_tile + 16 * wy) + g;            // + 8 mb: output row of fragment row g
  const int x0 = 2 * a_tile + 16 * wx + 4 * t;          // first of this lane's four columns: block 0 holds x0, x0 + 1, block 1 x0 + 2, x0 + 3
  const int ow = A.oct.w, oh = A.oct.h;
  const size_t row8 = (size_t)8 * A.oct.pitch;
  const bool interior = b_tile >= 0 && 2 * (b_tile + M0_SH) <= oh && 2 * (a_tile + M0_SW) <= ow;   // uniform over the group
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
      case 2: mma0_level<2>(S, Tw, wfs, clo0, warp, lane, cur, stage, n_stages); break;
      case 3: mma0_level<3>(S, Tw, wfs, clo0, warp, lane, cur, stage, n_stages); break;
      case 4: mma0_level<4>(S, Tw, wfs, clo0, warp, lane, cur, stage, n_stages); break;
      default: mma0_level<5>(S, Tw, wfs, clo0, warp, lane, cur, stage, n_stages); break;
    }
    const bool wg = A.keep_gauss != 0, wd = s > 0;
    float *gl = gthr + (long long)s * A.plane;
    if (!live) {
    } else if (interior) {
      if (wg) {
        float *r = gl;
#pragma unroll
        for (int mb = 0; mb < 4; mb++, r += row8)
          *reinterpret_cast<float4 *>(r) = make_float4(m0_cvt(cur[mb][0][0]), m0_cvt(cur[mb][0][1]), m0_cvt(cur[mb][1][0]), m0_cvt(cur[mb][1][1]));
      }
      if (wd) {
        float *r = gl + A.dog_delta;
#pragma unroll
        for (int mb = 0; mb < 4; mb++, r += row8)
          *reinterpret_cast<float4 *>(r) = make_float4(m0_cvt(prv[mb][0][0] - cur[mb][0][0]), m0_cvt(prv[mb][0][1] - cur[mb][0][1]),
                                                       m0_cvt(prv[mb][1][0] - cur[mb][1][0]), m0_cvt(prv[mb][1][1] - cur[mb][1][1]));
      }
    } else {
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int y = y0 + 8 * mb;
        if (y >= 0 && y < oh) {
#pragma unroll
          for (int nb = 0; nb < 2; nb++) {
            if (x0 + 2 * nb < ow) {                     // the octave width is even: both columns or none
              float *r = gl + mb * row8 + 2 * nb;
              if (wg) *reinterpret_cast<float2 *>(r) = make_float2(m0_cvt(cur[mb][nb][0]), m0_cvt(cur[mb][nb][1]));
              if (wd)
                *reinterpret_cast<float2 *>(r + A.dog_delta) =
                    make_float2(m0_cvt(prv[mb][nb][0] - cur[mb][nb][0]), m0_cvt(prv[mb][nb][1] - cur[mb][nb][1]));
            }
          }
        }
      }
    }
    if (live && A.has_next && s == A.spo && (g & 1) == 0) {     // in[2a][2b] (matrix2d.js:129): even rows, even columns
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int y = y0 + 8 * mb;
        const int nr = (y >> 1) + A.oct.seed_off;       // row of the next octave (strip-local)
        if (y >= 0 && y < oh && nr >= 0 && nr < A.next.h) {
#pragma unroll
          for (int nb = 0; nb < 2; nb++) {
            const int x = x0 + 2 * nb;
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
  if (tile0 + (int)gridDim.x * HV < A.n_tiles) __syncthreads();   // every warp has left its tile: S and raw are free
  }                                                     // tiles
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

int mma0_tile_rows(const OctaveDev &oct, int src_h) { return (src_h + ((oct.y_top >> 1) & 3) + M0_SH - 1) / M0_SH; }

// tile_row0 / tile_rows: the band of tile rows to run (tile_rows < 0: all of them)
bool launch_oct0_mma(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                     const OctaveDev &oct, const OctaveDev *next, const double *d_frags, const LevelPlan *plans,
                     int nlev, int spo, int keep_gauss, int tile_row0, int tile_rows)
{
  Mma0Args A;
  memset(&A, 0, sizeof A);
  A.plane = oct.gauss[1] - oct.gauss[0];
  A.dog_delta = oct.dog[0] - oct.gauss[1];
  // the kernel walks the planes with a constant stride and stores column pairs: check the layout it assumes
  bool ok = (oct.pitch & 3) == 0 && (oct.w & 1) == 0 && (oct.h & 1) == 0 && (A.plane & 3) == 0 && (A.dog_delta & 3) == 0 &&
            (((uintptr_t)oct.gauss[0]) & 15) == 0;
  for (int s = 0; s < nlev; s++) ok = ok && oct.gauss[s] == oct.gauss[0] + (long long)s * A.plane;
  for (int s = 0; s + 1 < nlev; s++) ok = ok && oct.dog[s] == oct.gauss[s + 1] + A.dog_delta;
  if (!ok) return false;
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < nlev; s++) A.radius[s] = plans[s].radius;
  A.wfrag = d_frags;
  A.row_shift = (oct.y_top >> 1) & 3;
#ifndef M0_HALVES
#define M0_HALVES 1
#endif
  const size_t smem = ((size_t)M0_HALVES * (M0_S_DOUBLES + M0_T_DOUBLES + M0_RAW_DOUBLES) + (size_t)nlev * M0_MAXD * 32) * sizeof(double);
  const int all_rows = (src_h + A.row_shift + M0_SH - 1) / M0_SH;
  A.tile_y0 = tile_rows < 0 ? 0 : tile_row0;
  const int n_rows = tile_rows < 0 ? all_rows : (tile_row0 + tile_rows <= all_rows ? tile_rows : all_rows - tile_row0);
  if (n_rows <= 0) return true;
  A.tiles_x = (src_w + M0_SW - 1) / M0_SW;
  A.n_tiles = A.tiles_x * n_rows;
  const int grid = mm_grid((A.n_tiles + M0_HALVES - 1) / M0_HALVES);
  cudaFuncSetAttribute(oct0_mma_kernel<M0_HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     // grows with the level count
  CUtensorMap src_map;
  memset(&src_map, 0, sizeof src_map);
  // the raw tile must start on a 128-byte boundary of shared memory: S, Ts and the fragments before it are whole multiples
  A.use_tma = 0;
  if ((dtype == SIFT_U8 || dtype == SIFT_F32) && ((size_t)(M0_S_DOUBLES + M0_T_DOUBLES) * sizeof(double)) % 128 == 0 &&
      ((size_t)M0_RAW_DOUBLES * sizeof(double)) % 128 == 0)
    A.use_tma = (mm_encode_2d(&src_map, dtype == SIFT_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                              dtype == SIFT_U8 ? 1 : 4, src, src_pitch, src_w, src_h, dtype == SIFT_U8 ? 64 : M0_SCOLS, M0_SCOLS)) ? 1 : 0;
  oct0_mma_kernel<M0_HALVES><<<grid, M0_THREADS * M0_HALVES, smem, st>>>(A, src_map);
  return true;
}

// =====================================================================================================
// Octaves >= 1: two passes over the octave base (the unrounded fp64 seed), all blurred levels per launch.
//   pass A (along x)  base[y][x]  ->  T_s[y][x]   fp64, row-major (no transposition: the fragment loads of a DMMA
//                                                  read a 4 x 8 patch of the staged tile in either orientation)
//   pass B (along y)  T_s[y][x]   ->  G_s, D_{s-1} = G_{s-1} - G_s (fp32), seed of the next octave
// The band of an 8-position block is  Wm[p][k] = w[k - p]  over K = 8 + taps - 1 samples; a lane's fragment of
// chunk d is  w[4 d + t - g]  for both operand orders, read straight from the zero-padded taps in shared memory
// (11 neighbouring doubles per load: no bank conflicts), so any radius runs the same code.
#define MS_WFRONT 8                       // zeros in front of a level's taps (4 d + t - g >= -7)
#define MS_WBACK 24                       // zeros behind (chunks are whole: up to 4 D + 3 - 0 taps are read)
#define MS_THREADS 256

struct MmaSepArgs {
  const double *src;                      // octave base: dense fp64 [h][w]
  int w, h;
  int nlev, rmax;
  int level[SIFT_MAX_LEVELS], radius[SIFT_MAX_LEVELS], woff[SIFT_MAX_LEVELS];   // woff: raw taps in the weight buffer
  int wsm[SIFT_MAX_LEVELS + 1];           // offset of the level's padded taps in shared memory (doubles); [nlev] = total
  double *T[SIFT_MAX_LEVELS];
  size_t t_pitch;                         // doubles per T row (even)
  int tile_pitch;                         // pass A: doubles per staged row (= 4 mod 16)
  int use_tma;                            // pass A: interior tiles arrive as one TMA box (src_map)
  int buf_rows[2];                        // pass B: rows of the two staging buffers (levels 0, 2, .. / 1, 3, ..)
  OctaveDev oct, next;
  int has_next, spo, keep_gauss;
  long long plane, dog_delta;
  int row_shift;                          // pass B tiles start at row -row_shift (mosaic strips: blocks aligned with the whole image's)
};

template <int NT = MS_THREADS>
__device__ __forceinline__ void ms_stage_taps(const double *__restrict__ weights, const MmaSepArgs &A, double *wsm)
{
  for (int li = 0; li < A.nlev; li++) {
    const int n = 2 * A.radius[li] + 1, len = A.wsm[li + 1] - A.wsm[li];
    for (int e = threadIdx.x; e < len; e += NT)
      wsm[A.wsm[li] + e] = (e >= MS_WFRONT && e < MS_WFRONT + n) ? __ldg(weights + A.woff[li] + e - MS_WFRONT) : 0.0;
  }
}

__device__ __forceinline__ void ms_cp8(double *dst_smem, const double *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void ms_cp16(double *dst_smem, const double *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// ---- pass A: CTA = 32 rows x 128 columns of every T_s; warp (wr, wc) = 16 rows (2 M blocks) x 32 columns (4 N blocks)
#define MA_ROWS 32
#define MA_COLS 128
#ifndef MSA_CTAS
#define MSA_CTAS 2
#endif
#ifndef MSA_PINGPONG
#define MSA_PINGPONG 0
#endif
#ifndef MSB_CTAS
#define MSB_CTAS 2
#endif
__global__ void __launch_bounds__(MS_THREADS, MSA_CTAS)
sep_a_mma_kernel(const double *__restrict__ weights, const __grid_constant__ MmaSepArgs A, const __grid_constant__ CUtensorMap src_map)
{
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long tile_bar;
  double *tile = smem;                                   // [32][tile_pitch]: columns x_tile - rmax .. (clamped)
  double *wsm = smem + MA_ROWS * A.tile_pitch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int x_tile = blockIdx.x * MA_COLS, y_tile = blockIdx.y * MA_ROWS;
  const int tp = A.tile_pitch;

  const int xl = x_tile - A.rmax;                        // first staged column (rmax is rounded up to even by the host)
  const bool by_tma = A.use_tma && xl >= 0 && xl + tp <= A.w && y_tile + MA_ROWS <= A.h;      // CTA-uniform
  if (by_tma) {
    // interior tile: ONE TMA box (32 rows x tile_pitch fp64 of the seed plane) lands dense in the tile, completion
    // on an mbarrier; border tiles need clamp-to-edge samples (the TMA unit would zero-fill) and are copied below
    if (tid == 0) {
      mm_bar_init(&tile_bar, 1);
      mm_tma_load_2d(tile, &src_map, &tile_bar, xl, y_tile, (unsigned)(MA_ROWS * tp * sizeof(double)));
    }
  } else if (xl >= 0 && xl + tp <= A.w && y_tile + MA_ROWS <= A.h && (A.w & 1) == 0) {
    // interior tile: whole 16-byte pairs, pointers advanced by constants (a warp copies 4 rows)
    const double *src = A.src + (size_t)(y_tile + warp) * A.w + xl + 2 * lane;
    double *dst = tile + warp * tp + 2 * lane;
#pragma unroll
    for (int i = 0; i < MA_ROWS / 8; i++, src += (size_t)8 * A.w, dst += 8 * tp)
      for (int cc = 2 * lane; cc < tp; cc += 64) ms_cp16(dst + cc - 2 * lane, src + cc - 2 * lane);
  } else {
    for (int rr = warp; rr < MA_ROWS; rr += MS_THREADS / 32) {       // every staged column holds a (clamped) sample: finite
      const double *row = A.src + (size_t)min(y_tile + rr, A.h - 1) * A.w;
      double *dst = tile + rr * tp;
      for (int cc = lane; cc < tp; cc += 32) ms_cp8(dst + cc, row + min(max(xl + cc, 0), A.w - 1));      // sift.js:116-119
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  ms_stage_taps(weights, A, wsm);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();                                       // taps, copied tile, barrier initialisation visible
  if (by_tma) mm_bar_wait(&tile_bar, 0);

  const int wr = warp >> 2, wc = warp & 3;
  const int y0 = y_tile + 16 * wr + g;                   // + 8 mb
  const int x0 = x_tile + 32 * wc + 2 * t;               // + 8 nb
#if MSA_PINGPONG
  // the two groups of four warps take turns on the DMMA pipe (see oct0_mma_kernel): a group blurs a level while the
  // other stores the level it has just finished
  if (wr == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
#endif
  for (int li = 0; li < A.nlev; li++) {
    const int R = A.radius[li];
    const int D = (2 * R + 8 + 3) >> 2;                  // chunks of the band: ceil((taps + 7) / 4)
#if MSA_PINGPONG
    asm volatile("bar.sync %0, 256;" ::"r"(1 + wr) : "memory");
#endif
    const double *wp = wsm + A.wsm[li] + MS_WFRONT + t - g;        // fragment of chunk d: wp[4 d] (zero outside the taps)
    const double *sp = tile + (16 * wr + g) * tp + (A.rmax - R) + 32 * wc + t;
    double acc[2][4][2];
#pragma unroll
    for (int mb = 0; mb < 2; mb++)
#pragma unroll
      for (int nb = 0; nb < 4; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    // fragment d of the band meets chunk d + 2 nb of the samples of block nb: no ramp, every DMMA is useful work
    for (int d = 0; d < D; d++) {
      const double wv = wp[4 * d];
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        const double a0 = sp[4 * d + 8 * nb], a1 = sp[8 * tp + 4 * d + 8 * nb];
        dmma884(acc[0][nb][0], acc[0][nb][1], a0, wv);
        dmma884(acc[1][nb][0], acc[1][nb][1], a1, wv);
      }
    }
#if MSA_PINGPONG
    asm volatile("bar.arrive %0, 256;" ::"r"(1 + (wr ^ 1)) : "memory");
#endif
    double *T = A.T[li];
#pragma unroll
    for (int mb = 0; mb < 2; mb++) {
      const int y = y0 + 8 * mb;
      if (y < A.h) {
        double *trow = T + (size_t)y * A.t_pitch;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
          const int x = x0 + 8 * nb;
          if (x + 1 < A.w) *reinterpret_cast<double2 *>(trow + x) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
          else if (x < A.w) trow[x] = acc[mb][nb][0];
        }
      }
    }
  }
}

// ---- pass B: CTA = 64 MB rows x 32 columns; warp w = rows 8 MB w .. (MB M blocks) x 32 columns (4 N blocks).
// Level s stages rows y_tile - R_s .. y_tile + 64 MB + R_s (+ slack, clamped) of T_s with cp.async, one level ahead
// of the one being blurred (two buffers); the previous level stays in registers for the DoG.
#define MB_COLS 32
#define MB_PITCH 32                                      // dense rows; the column index is XOR-swizzled by the row (see MB_SWZ)
// rows r, r+1, r+2, r+3 of a chunk are read by the lanes t = 0..3 at the columns {0,1,4,5,..} of an N block: the
// offsets 0, 8, 2, 10 keep the four column sets on distinct 8-byte banks, and 16-byte pairs stay together
#define MB_SWZ(r) (8 * ((r) & 1) + 2 * (((r) >> 1) & 1))
#define MB_SLACK 4
#ifndef MSB_WARPS
#define MSB_WARPS 8
#endif
#define MSB_THREADS (32 * MSB_WARPS)
template <int MB>
__global__ void __launch_bounds__(MSB_THREADS, MSB_CTAS)
sep_b_mma_kernel(const double *__restrict__ weights, const MmaSepArgs A)
{
  constexpr int Y = 8 * MB * MSB_WARPS;
  extern __shared__ __align__(128) double smem[];
  double *buf0 = smem, *buf1 = smem + A.buf_rows[0] * MB_PITCH;
  double *wsm = buf1 + A.buf_rows[1] * MB_PITCH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int x_tile = blockIdx.x * MB_COLS, y_tile = blockIdx.y * Y - A.row_shift;
  const int w = A.oct.w, h = A.oct.h;

  auto stage = [&](const int li) {                       // all threads: rows y_tile - R .. of T_li into its buffer
    const int R = A.radius[li];
    const double *T = A.T[li];
    double *buf = (li & 1) ? buf1 : buf0;
    const int rows = Y + 2 * R + MB_SLACK;
    const bool wide = x_tile + MB_COLS <= w;             // CTA-uniform: whole 16-byte pairs inside the rows
    if (wide && y_tile - R >= 0 && y_tile - R + rows <= h) {
      // interior tile: no clamping, pointers advanced by constants (16 threads per row, 16 rows per sweep)
      const double *src = T + (size_t)(y_tile - R + (tid >> 4)) * A.t_pitch + x_tile + (tid & 15) * 2;
      double *dst = buf + (tid >> 4) * MB_PITCH + (((tid & 15) * 2) ^ MB_SWZ(tid >> 4));
      const size_t sstep = (size_t)(MSB_THREADS / 16) * A.t_pitch;
      constexpr int RS = MSB_THREADS / 16;             // rows per sweep
      int rr = tid >> 4;
      for (; rr + 3 * RS < rows; rr += 4 * RS, src += 4 * sstep, dst += 4 * RS * MB_PITCH) {
        ms_cp16(dst, src); ms_cp16(dst + RS * MB_PITCH, src + sstep);
        ms_cp16(dst + 2 * RS * MB_PITCH, src + 2 * sstep); ms_cp16(dst + 3 * RS * MB_PITCH, src + 3 * sstep);
      }
      for (; rr < rows; rr += RS, src += sstep, dst += RS * MB_PITCH) ms_cp16(dst, src);
    } else {
      for (int e = tid; e < rows * (MB_COLS / 2); e += MSB_THREADS) {
        const int rr = e >> 4, c2 = (e & 15) * 2;
        const double *row = T + (size_t)min(max(y_tile - R + rr, 0), h - 1) * A.t_pitch;          // sift.js:116-119 clamp-to-edge
        double *dst = buf + rr * MB_PITCH + (c2 ^ MB_SWZ(rr));
        if (wide) ms_cp16(dst, row + x_tile + c2);
        else { ms_cp8(dst, row + min(x_tile + c2, w - 1)); ms_cp8(dst + 1, row + min(x_tile + c2 + 1, w - 1)); }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0);
  ms_stage_taps<MSB_THREADS>(weights, A, wsm);

  const int y0 = y_tile + 8 * MB * warp + g;             // + 8 mb
  const int x0 = x_tile + 4 * t;                         // + 16 (nb >> 1) + 2 (nb & 1): first column of block nb
  const bool interior = y_tile >= 0 && y_tile + Y <= h && x_tile + MB_COLS <= w;     // CTA-uniform
  const size_t row8 = (size_t)8 * A.oct.pitch;
  float *gthr = A.oct.gauss[A.level[0]] + ((long long)y0 * A.oct.pitch + x0);

  // G_s and D_{s-1} = G_{s-1} - G_s (sift.js:172) of the block pair (mb, 2 q), (mb, 2 q + 1), from the unrounded accumulators
  auto emit = [&](const int li, const int mb, const int q, const double (&older)[MB][4][2], const double (&newer)[MB][4][2]) {
    float *r = gthr + (long long)li * A.plane + mb * row8 + 16 * q;
    const int n0 = 2 * q, n1 = 2 * q + 1;
    if (interior) {
      if (A.keep_gauss)
        *reinterpret_cast<float4 *>(r) = make_float4((float)newer[mb][n0][0], (float)newer[mb][n0][1], (float)newer[mb][n1][0], (float)newer[mb][n1][1]);
      *reinterpret_cast<float4 *>(r + A.dog_delta) =
          make_float4((float)(older[mb][n0][0] - newer[mb][n0][0]), (float)(older[mb][n0][1] - newer[mb][n0][1]),
                      (float)(older[mb][n1][0] - newer[mb][n1][0]), (float)(older[mb][n1][1] - newer[mb][n1][1]));
    } else {
      const int y = y0 + 8 * mb, x = x0 + 16 * q;
      if (y >= 0 && y < h) {
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (x + i < w) {
            const double nv = newer[mb][n0 + (i >> 1)][i & 1], ov = older[mb][n0 + (i >> 1)][i & 1];
            if (A.keep_gauss) r[i] = (float)nv;
            r[A.dog_delta + i] = (float)(ov - nv);
          }
      }
    }
  };

  // One level: stage the next tile, blur this one, and meanwhile write out the PREVIOUS level (its conversions and
  // stores ride in the shadow of this level's DMMAs instead of idling the pipe between two levels).
  //   older = level li - 2, newer = level li - 1 (both unrounded), acc receives level li.
  auto level = [&](const int li, const double (&older)[MB][4][2], const double (&newer)[MB][4][2], double (&acc)[MB][4][2]) {
    const int R = A.radius[li];
    const int D = (2 * R + 8 + 3) >> 2;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                     // tile li (and, first time, the taps) visible; the other buffer is free
    if (li + 1 < A.nlev) stage(li + 1);
    const double *wp = wsm + A.wsm[li] + MS_WFRONT + t - g;
    // N block nb takes the columns {0,1, 4,5, 8,9, 12,13} + 2 (nb & 1) + 16 (nb >> 1): a lane's blocks 0, 1 (2, 3)
    // hold four neighbouring columns of one row -> one 16-byte store per row, plane and block pair
    const double *sp = ((li & 1) ? buf1 : buf0) + (8 * MB * warp + t) * MB_PITCH;
    const int c0 = (4 * (g >> 1) + (g & 1)) ^ MB_SWZ(t), c1 = (4 * (g >> 1) + (g & 1) + 2) ^ MB_SWZ(t);
#pragma unroll
    for (int mb = 0; mb < MB; mb++)
#pragma unroll
      for (int nb = 0; nb < 4; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    auto chunk = [&](const int d) {                      // fragment d meets the sample rows 8 mb + 4 d + t of block mb
      const double wv = wp[4 * d];
#pragma unroll
      for (int mb = 0; mb < MB; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++)
          dmma884(acc[mb][nb][0], acc[mb][nb][1], wv, sp[(8 * mb + 4 * d) * MB_PITCH + 16 * (nb >> 1) + ((nb & 1) ? c1 : c0)]);
    };
    int d = 0;
#pragma unroll
    for (int k = 0; k < 2 * MB; k++) {                   // the first chunks carry one block pair of the previous level each
      if (d < D) { chunk(d); d++; }
      if (d < D) { chunk(d); d++; }
      if (li > 0) emit(li - 1, k >> 1, k & 1, older, newer);
    }
    for (; d < D; d++) chunk(d);
    if (A.has_next && A.level[li] == A.spo && (g & 1) == 0) {      // matrix2d.js:129 in[2a][2b]: even rows, even columns
#pragma unroll
      for (int mb = 0; mb < MB; mb++) {
        const int y = y0 + 8 * mb;
        const int nr = (y >> 1) + A.oct.seed_off;        // row of the next octave (strip-local)
        if (y >= 0 && y < h && nr >= 0 && nr < A.next.h) {
#pragma unroll
          for (int nb = 0; nb < 4; nb++) {
            const int x = x0 + 16 * (nb >> 1) + 2 * (nb & 1);
            if (x < w) {
              A.next.seed64[(size_t)nr * A.next.w + (x >> 1)] = acc[mb][nb][0];
              A.next.gauss[0][(size_t)nr * A.next.pitch + (x >> 1)] = (float)acc[mb][nb][0];
            }
          }
        }
      }
    }
  };

  double ga[MB][4][2], gb[MB][4][2], gc[MB][4][2];
#pragma unroll
  for (int mb = 0; mb < MB; mb++)
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {                     // level 0 of octaves >= 1 is the unblurred seed (background.js:114-130)
      const int y = min(max(y0 + 8 * mb, 0), h - 1), x = x0 + 16 * (nb >> 1) + 2 * (nb & 1);
      gb[mb][nb][0] = A.src[(size_t)y * w + min(x, w - 1)];
      gb[mb][nb][1] = A.src[(size_t)y * w + min(x + 1, w - 1)];
      ga[mb][nb][0] = ga[mb][nb][1] = 0.0;
    }
  // roles rotate: (older, newer, acc) = (a, b, c) -> (b, c, a) -> (c, a, b)
  for (int li = 0; li < A.nlev; li += 3) {
    level(li, ga, gb, gc);
    if (li + 1 < A.nlev) level(li + 1, gb, gc, ga); else break;
    if (li + 2 < A.nlev) level(li + 2, gc, ga, gb); else break;
  }
  // the last level is written out on its own (its roles follow from its index mod 3)
  {
    const int last = A.nlev - 1;
#pragma unroll
    for (int k = 0; k < 2 * MB; k++) {
      if (last % 3 == 0) emit(last, k >> 1, k & 1, gb, gc);
      else if (last % 3 == 1) emit(last, k >> 1, k & 1, gc, ga);
      else emit(last, k >> 1, k & 1, ga, gb);
    }
  }
}

// ---- host side of octaves >= 1 ----------------------------------------------------------------------------
static int ms_tile_pitch(int rmax)
{
  int p = MA_COLS + 2 * rmax + 24;                       // + the chunks that run past the last tap (zero weights)
  return p + ((4 - (p & 15)) & 15);                      // smallest p' >= p with p' % 16 == 4
}

static size_t ms_fill(MmaSepArgs &A, const LevelPlan *plans, int nlev_total, int w, int h, double *tbase)
{
  memset(&A, 0, sizeof A);
  A.w = w; A.h = h;
  A.nlev = nlev_total - 1;
  A.t_pitch = (size_t)((w + 1) & ~1);
  int off = 0;
  for (int i = 0; i < A.nlev; i++) {
    const LevelPlan &p = plans[1 + i];
    A.level[i] = 1 + i; A.radius[i] = p.radius; A.woff[i] = p.woff;
    A.rmax = A.rmax > p.radius ? A.rmax : p.radius;
    A.wsm[i] = off;
    off += (MS_WFRONT + 2 * p.radius + 1 + MS_WBACK + 1) & ~1;
    A.T[i] = tbase + (size_t)i * h * A.t_pitch;
  }
  A.wsm[A.nlev] = off;
  A.rmax = (A.rmax + 1) & ~1;                            // even: the staged rows of pass A start on 16-byte boundaries
  A.tile_pitch = ms_tile_pitch(A.rmax);
  return (size_t)A.nlev * h * A.t_pitch;
}

static void ms_buf_rows(MmaSepArgs &A, int mb)
{
  A.buf_rows[0] = A.buf_rows[1] = 0;
  for (int i = 0; i < A.nlev; i++) {
    const int rows = 8 * MSB_WARPS * mb + 2 * A.radius[i] + MB_SLACK;
    if (rows > A.buf_rows[i & 1]) A.buf_rows[i & 1] = rows;
  }
}
static size_t ms_smem_a(const MmaSepArgs &A) { return ((size_t)MA_ROWS * A.tile_pitch + A.wsm[A.nlev]) * sizeof(double); }
static size_t ms_smem_b(MmaSepArgs &A, int mb)
{
  ms_buf_rows(A, mb);
  return ((size_t)(A.buf_rows[0] + A.buf_rows[1]) * MB_PITCH + A.wsm[A.nlev]) * sizeof(double);
}
// 64-row tiles (MB = 1): 124 registers, two CTAs per SM; 128-row tiles (MB = 2) spill at that budget and measured slower
static int ms_pick_mb(MmaSepArgs &A) { (void)A; return 1; }

size_t mma_sep_t_elems(const LevelPlan *plans, int nlev, int w, int h)
{
  (void)plans;
  return (size_t)(nlev - 1) * h * ((w + 1) & ~1);
}

bool mma_sep_supported(const LevelPlan *plans, int nlev, int w, int h)
{
  if (nlev < 2) return false;
  MmaSepArgs A;
  ms_fill(A, plans, nlev, w, h, nullptr);
  for (int i = 0; i < A.nlev; i++)
    if (A.radius[i] < 1) return false;
  return ms_smem_a(A) <= 200 * 1024 && ms_smem_b(A, 1) <= 200 * 1024;
}

void launch_mma_sep(cudaStream_t st, const OctaveDev &oct, const OctaveDev *next, const double *d_weights,
                    const LevelPlan *plans, double *tbase, int spo, int keep_gauss)
{
  MmaSepArgs A;
  ms_fill(A, plans, oct.nlev, oct.w, oct.h, tbase);
  A.src = oct.seed64;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss;
  A.plane = oct.gauss[1] - oct.gauss[0];
  A.dog_delta = oct.dog[0] - oct.gauss[1];
  A.row_shift = oct.y_top & 7;
  {
    const size_t smem = ms_smem_a(A);
    dim3 grid((oct.w + MA_COLS - 1) / MA_COLS, (oct.h + MA_ROWS - 1) / MA_ROWS);
    CUtensorMap src_map;
    memset(&src_map, 0, sizeof src_map);
    A.use_tma = (mm_encode_2d(&src_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, oct.seed64, (size_t)oct.w * sizeof(double), oct.w, oct.h,
                              A.tile_pitch, MA_ROWS)) ? 1 : 0;
    cudaFuncSetAttribute(sep_a_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sep_a_mma_kernel<<<grid, MS_THREADS, smem, st>>>(d_weights, A, src_map);
  }
  const int mb = ms_pick_mb(A);
  const size_t smem = ms_smem_b(A, mb);
  const int ytile = 8 * MSB_WARPS * mb;
  dim3 grid((oct.w + MB_COLS - 1) / MB_COLS, (oct.h + A.row_shift + ytile - 1) / ytile);
  if (mb == 2) {
    cudaFuncSetAttribute(sep_b_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sep_b_mma_kernel<2><<<grid, MSB_THREADS, smem, st>>>(d_weights, A);
  } else {
    cudaFuncSetAttribute(sep_b_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sep_b_mma_kernel<1><<<grid, MSB_THREADS, smem, st>>>(d_weights, A);
  }
}
