// blur_oct0.cu -- octave 0 in ONE persistent, warp-specialised kernel ("column pass first, row pass last").
//
// Reference path restated (as blur_fused.cu): Matrix2D_linearResize(input, 0.5) (background.js:84,
// matrix2d.js:112-138), then for every level SIFT_blurMatrix2DChunk of that pixel-doubled base
// (background.js:145-210, sift.js:72-149, clamp-to-edge per axis sift.js:116-119), SIFT_subtractMatrix2DChunk
// between neighbouring levels (sift.js:154-188, finer minus coarser) and the rate-2.0 resize of level `spo`
// that seeds the next octave (background.js:114-130).  Polyphase form as in blur_fused.cu: the 2R+1 taps over
// the doubled image collapse onto R+1 merged taps per output phase over SOURCE samples, in both directions.
//
// Why this shape (profiles/r01_ncu_fused_octave0_1.txt, r02_ncu_oct0_v2.txt): the fp64 pipe is the binding
// resource, and the tile kernels of round 1 kept it 34-38 % busy because every warp of a CTA walked through the
// same phases together -- tap loop (fp64 pipe), then conversion + stores (XU / LSU, pipe idle), then a CTA
// barrier with a third of the warps idle.  Here the phases belong to DIFFERENT warps of one resident CTA per SM:
//   * 6 "column" warps run the y pass of level L+1 on the un-doubled source columns, S -> Tv[(L+1) & 1]
//     ([64 rows][<= 48 columns], both row phases), while
//   * 8 "row" warps run the x pass of level L from Tv[L & 1] and its epilogue: a thread's 16 outputs (8 source
//     positions x 2 phases) are 16 neighbouring pixels of one row, kept unrounded in registers for the next
//     level's DoG (the two accumulator sets swap roles every level: no copies), rounded once to fp32 into
//     swizzled staging boxes, and handed to the TMA unit (cp.async.bulk.tensor stores: whole 128-byte lines
//     leave the SM whatever the thread layout, zero store instructions in the warps);
//   * the two groups meet only through mbarriers (full / empty per Tv buffer); the CTA is persistent (one per
//     SM, tiles dealt round-robin) so the column warps start the next tile -- source box by TMA, issued a tile
//     ahead -- while the row warps finish the current one.
//   * tap loops: register sliding window (one sample load + one broadcast weight-pair load per 16 DFMA), the
//     first tap initialises the accumulators, the source tile is converted to fp64 once per tile.
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "common.cuh"

#define V2_SW 32                       // source columns per tile
#define V2_SH 32                       // source rows per tile
#define V2_HALO 8                      // supports R <= 16
#define V2_SR (V2_SH + 2 * V2_HALO)    // 48 staged source rows (+ 1 slack row for the window prefetch)
#define V2_SC (V2_SW + 2 * V2_HALO)    // 48 staged source columns
#define V2_SP 49                       // S pitch in doubles (odd: conflict-free along rows and columns)
#define V2_TP 49                       // Tv pitch in doubles: 48 columns + 1 slack column for the window prefetch
#define V2_TR (2 * V2_SH)              // 64 Tv rows (both row phases)
#define V2_MAXR 16
#define V2_WSTRIDE 24                  // padded unified taps per level (np <= 18), as blur_fused.cu's table

#define WS_ROW_WARPS 8
#define WS_COL_WARPS 6
#define WS_ROW_THREADS (32 * WS_ROW_WARPS)     // 256: one output row x 16 columns each
#define WS_COL_THREADS (32 * WS_COL_WARPS)     // 192 >= 4 * 48 column-pass items
#define WS_THREADS (WS_ROW_THREADS + WS_COL_THREADS)

// shared memory, in doubles
#define V2_S_DOUBLES ((V2_SR + 1) * V2_SP)         // 2401
#define V2_T_DOUBLES (V2_TR * V2_TP)               // 3136
#define V2_T0_OFF ((V2_S_DOUBLES + 15) & ~15)
#define V2_T1_OFF (V2_T0_OFF + ((V2_T_DOUBLES + 15) & ~15))
#define V2_W_OFF (V2_T1_OFF + ((V2_T_DOUBLES + 15) & ~15))
#define V2_W_DOUBLES (SIFT_MAX_LEVELS * 2 * V2_WSTRIDE)
#define V2_RAW_OFF ((V2_W_OFF + V2_W_DOUBLES + 15) & ~15)      // TMA landing zone of the raw source box (128-byte aligned)
#define V2_RAW_DOUBLES (V2_SR * V2_SC * 4 / 8)                 // 48 x 48 f32 (u8: 64 x 48 bytes)
// output staging for the TMA stores, double-buffered by level: per buffer Gaussian then DoG, per plane two boxes of
// [64 rows][32 floats] (128-byte rows, SWIZZLE_128B); 1024-byte aligned so that the swizzle pattern
// (chunk ^ (row & 7)) starts at row 0
#define V2_STAGE_OFF ((V2_RAW_OFF + V2_RAW_DOUBLES + 127) & ~127)
#define V2_BOX_BYTES (V2_TR * 128)                 // 8192
#define V2_STAGE_BYTES (4 * V2_BOX_BYTES)          // 32 KB per buffer
#define V2_SMEM_BYTES (V2_STAGE_OFF * 8 + 2 * V2_STAGE_BYTES)

struct Oct0Args {
  const void *src;
  size_t src_pitch;
  int src_w, src_h, dtype;
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, nlev;
  int radius[SIFT_MAX_LEVELS];
  int woff;                       // offset of the merged tap table [nlev][V2_WSTRIDE]{w0, w1} in the weight buffer
  int use_tma;                    // source tile by TMA (u8 / f32 sources with 16-byte aligned rows)
  int tiles_x, n_tiles;
  const CUtensorMap *out_maps;    // [nlev] Gaussian planes then [nlev-1] DoG planes of this octave, box 32 x 64, SWIZZLE_128B
};

__device__ __forceinline__ unsigned v2_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// v / 255.0 exactly as the reference computes it (image-utils.js:114), see blur_fused.cu
__device__ __forceinline__ double v2_u8_over_255(unsigned v)
{
  const double r = 1.0 / 255.0;
  const double x = (double)v;
  const double q = x * r;
  return fma(fma(-q, 255.0, x), r, q);
}

__device__ __forceinline__ void v2_mbar_wait(unsigned bar, unsigned parity)
{
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void v2_mbar_arrive(unsigned bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Two-phase sliding window over 8 neighbouring positions:
//   a0[k] = sum_{j<np} w0[j] v[k+j],  a1[k] = sum_{j<np} w1[j] v[k+j],   v[p] = base[p * STRIDE]
// `w2` holds the two phases interleaved ({w0[j], w1[j]}: one broadcast 16-byte load per tap).  The window rotates
// through 8 registers with compile-time indices; the first tap initialises the accumulators (DMUL), whole groups
// of 8 taps follow, the remainder is decomposed as 4 + 2 + 1 with one code block per (size, rotation) pair.
// Positions up to np + 7 are read (np + 6 used).
#define V2_STEP(U, JJ)                                                                      \
  {                                                                                         \
    const double2 c = w2[(JJ)];                                                             \
    _Pragma("unroll") for (int k = 0; k < 8; k++) {                                         \
      a0[k] = fma(c.x, vw[(k + (U)) & 7], a0[k]);                                           \
      a1[k] = fma(c.y, vw[(k + (U)) & 7], a1[k]);                                           \
    }                                                                                       \
    vw[(U) & 7] = nxt[(JJ) * STRIDE];                                                       \
  }
template <int STRIDE>
__device__ __forceinline__ void v2_window(const double *__restrict__ base, const double2 *__restrict__ w2, const int np,
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
    V2_STEP(1, j) V2_STEP(2, j + 1) V2_STEP(3, j + 2) V2_STEP(4, j + 3)
    V2_STEP(5, j + 4) V2_STEP(6, j + 5) V2_STEP(7, j + 6) V2_STEP(0, j + 7)
  }
  const int rem = np - j;                                    // 0..7, uniform over the CTA
  if (rem & 4) { V2_STEP(1, j) V2_STEP(2, j + 1) V2_STEP(3, j + 2) V2_STEP(4, j + 3) }
  if (rem & 2) {
    if (rem & 4) { V2_STEP(5, j + 4) V2_STEP(6, j + 5) }
    else { V2_STEP(1, j) V2_STEP(2, j + 1) }
  }
  if (rem & 1) {
    switch (rem & 6) {
      case 0: V2_STEP(1, j) break;
      case 2: V2_STEP(3, j + 2) break;
      case 4: V2_STEP(5, j + 4) break;
      default: V2_STEP(7, j + 6) break;
    }
  }
}

// Column pass of level s: for every staged source column the row pass of this level will read, both row phases of
// the tile's 32 source rows -> Tv[2 r + phase][column].  One item = one column x 8 source rows x 2 phases.
__device__ __forceinline__ void v2_column_pass(const double *__restrict__ S, double *__restrict__ Tv,
                                               const double *__restrict__ Wt, int s, int R, int ctid)
{
  const int clo = -((R + 1) / 2);
  const int np = R + 1 + (R & 1);                    // unified taps per phase (one is zero in each phase when R is odd)
  const int ncols = V2_SW + np - 1;                  // <= 48
  const int c_first = V2_HALO + clo;
  const double2 *w2 = reinterpret_cast<const double2 *>(Wt + s * 2 * V2_WSTRIDE);
  if (ctid < 4 * ncols) {
    const int g = (ctid >= ncols) + (ctid >= 2 * ncols) + (ctid >= 3 * ncols);
    const int cc = c_first + (ctid - g * ncols);
    double a0[8], a1[8];
    v2_window<V2_SP>(S + (V2_HALO + 8 * g + clo) * V2_SP + cc, w2, np, a0, a1);
    double *t = Tv + (16 * g) * V2_TP + cc;
#pragma unroll
    for (int k = 0; k < 8; k++) { t[(2 * k) * V2_TP] = a0[k]; t[(2 * k + 1) * V2_TP] = a1[k]; }
  }
}

struct Oct0Thread {
  int y, x0, yg;          // tile row, first output column (global), global output row
  bool row_ok, seed_lane;
  unsigned stage;         // shared-memory byte address of this thread's row in its Gaussian staging box (buffer 0)
  unsigned swz;           // chunk index of its first 16-byte piece, already XORed with (row & 7)
};

__device__ __forceinline__ void v2_sts128(unsigned addr, float a, float b, float c, float d)
{
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Row pass of level s + epilogue.  c*: this level's unrounded outputs (kept for the next level's DoG);
// p*: the previous level's.  empty_bar: the Tv buffer is handed back to the column warps as soon as the window has
// been read.  The fp32 results go to the swizzled staging boxes of `stage_buf`.
__device__ __forceinline__ void v2_row_pass(const Oct0Args &A, const double *__restrict__ Tv, const double *__restrict__ Wt,
                                            int s, int R, int tid, const Oct0Thread &T, unsigned empty_bar, unsigned stage_buf,
                                            double (&c0)[8], double (&c1)[8], const double (&p0)[8], const double (&p1)[8])
{
  const int clo = -((R + 1) / 2);
  const int np = R + 1 + (R & 1);
  const int gq = tid >> 6;
  v2_window<1>(Tv + T.y * V2_TP + V2_HALO + 8 * gq + clo, reinterpret_cast<const double2 *>(Wt + s * 2 * V2_WSTRIDE), np, c0, c1);
  __syncwarp();
  if ((tid & 31) == 0) v2_mbar_arrive(empty_bar);
  // G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) from the unrounded accumulators; pixel x0 + 2k + phase
  const unsigned st = T.stage + stage_buf * V2_STAGE_BYTES;
  if (A.keep_gauss) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      v2_sts128(st + (((T.swz ^ q) & 7) << 4), (float)c0[2 * q], (float)c1[2 * q], (float)c0[2 * q + 1], (float)c1[2 * q + 1]);
  }
  if (s > 0) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      v2_sts128(st + 2 * V2_BOX_BYTES + (((T.swz ^ q) & 7) << 4), (float)(p0[2 * q] - c0[2 * q]), (float)(p1[2 * q] - c1[2 * q]),
                (float)(p0[2 * q + 1] - c0[2 * q + 1]), (float)(p1[2 * q + 1] - c1[2 * q + 1]));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA unit
  if (s == A.spo && T.seed_lane && T.row_ok) {        // in[2a][2b] (matrix2d.js:129): even rows, even columns (phase 0)
    const int nr = (T.yg >> 1) + A.oct.seed_off;      // row of the next octave (strip-local)
    if (nr >= 0 && nr < A.next.h) {
      const int nc = T.x0 >> 1;
      double *sp = A.next.seed64 + (size_t)nr * A.next.w + nc;
      float *fp = A.next.gauss[0] + (size_t)nr * A.next.pitch + nc;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (nc + k < A.next.w) { sp[k] = c0[k]; fp[k] = (float)c0[k]; }
    }
  }
}

// One thread, after the row warps' barrier: the staged Gaussian level s and DoG level s-1 -> global memory (the TMA
// unit clips boxes that hang over the right / bottom edge of the plane).
__device__ __forceinline__ void v2_store_level(const Oct0Args &A, unsigned stage, int s, int x_tile, int y_tile)
{
#pragma unroll
  for (int b = 0; b < 2; b++) {
    if (A.keep_gauss)
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(A.out_maps + s), "r"(stage + b * V2_BOX_BYTES), "r"(x_tile + 32 * b), "r"(y_tile) : "memory");
    if (s > 0)
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(A.out_maps + A.nlev + s - 1), "r"(stage + (2 + b) * V2_BOX_BYTES), "r"(x_tile + 32 * b), "r"(y_tile) : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1)
oct0_ws_kernel(const double *__restrict__ weights, const __grid_constant__ Oct0Args A, const __grid_constant__ CUtensorMap tmap)
{
  extern __shared__ __align__(1024) double smem[];
  __shared__ __align__(8) unsigned long long bars[5];     // full[2], empty[2], source box landed
  double *S = smem;                                   // [49][49] source tile, fp64, clamp-to-edge applied
  double *Tv0 = smem + V2_T0_OFF;                     // 2 x [64][49] column-filtered rows, both row phases
  double *Tv1 = smem + V2_T1_OFF;
  double *Wt = smem + V2_W_OFF;                       // [nlev][V2_WSTRIDE]{w0, w1}
  const int tid_cta = threadIdx.x;
  // the issue arbiter favours high warp ids: the row warps (the critical path) take them
  const int tid = tid_cta - WS_COL_THREADS;           // row threads 0..255; negative for the column warps
  const unsigned full0 = v2_smem_u32(&bars[0]), empty0 = v2_smem_u32(&bars[2]), src_bar = v2_smem_u32(&bars[4]);

  if (tid_cta == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0), "r"(WS_COL_WARPS));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8), "r"(WS_COL_WARPS));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0), "r"(WS_ROW_WARPS));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8), "r"(WS_ROW_WARPS));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(src_bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid_cta; e < A.nlev * 2 * V2_WSTRIDE; e += WS_THREADS) Wt[e] = __ldg(weights + A.woff + e);
  if (tid_cta < V2_SP) S[V2_SR * V2_SP + tid_cta] = 0.0;                                          // slack row
  if (tid_cta < V2_SR) S[tid_cta * V2_SP + V2_SC] = 0.0;                                          // pad column
  __syncthreads();

  if (tid < 0) {
    // ------------------------------------------------------------------------------- column warps
    const int ctid = tid_cta;
    const unsigned raw_addr = v2_smem_u32(smem + V2_RAW_OFF);
    auto issue_source = [&](int t) {                  // one thread: arm the barrier, one box of the raw source
      const int ty = t / A.tiles_x, tx = t - ty * A.tiles_x;
      // the box starts on a 16-byte boundary of its row (a TMA requirement): u8 tiles take 16 columns of left
      // halo (box 64 x 48), f32 tiles the 8 they need (box 48 x 48)
      const unsigned bytes = A.dtype == SIFT_U8 ? V2_SR * 64u : V2_SR * V2_SC * 4u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the box was read through the generic proxy
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(src_bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
          ::"r"(raw_addr), "l"(&tmap), "r"(src_bar), "r"(tx * V2_SW - (A.dtype == SIFT_U8 ? 16 : V2_HALO)), "r"(ty * V2_SH - V2_HALO)
          : "memory");
    };
    if (A.use_tma && ctid == 0 && (int)blockIdx.x < A.n_tiles) issue_source(blockIdx.x);
    unsigned n = 0;                                    // levels produced so far (all tiles): Tv buffer n & 1, its use n >> 1
    int it = 0;
    for (int t = blockIdx.x; t < A.n_tiles; t += gridDim.x, it++) {
      const int ty = t / A.tiles_x, tx = t - ty * A.tiles_x;
      const int a_tile = tx * V2_SW, b_tile = ty * V2_SH;
      asm volatile("bar.sync 2, %0;" ::"n"(WS_COL_THREADS) : "memory");     // every column warp is done with the previous S
      static_assert(V2_SR * V2_SC == 12 * WS_COL_THREADS, "tile conversion assumes 12 samples per column thread");
      if (A.use_tma) {
        v2_mbar_wait(src_bar, (unsigned)(it & 1));
        // raw box -> S with clamp-to-edge (sift.js:116-119) by indexing: a staged element whose source coordinate
        // lies outside the image takes the raw element of the clamped coordinate (always inside the box)
        const unsigned char *raw8 = reinterpret_cast<const unsigned char *>(smem + V2_RAW_OFF);
        const float *raw32 = reinterpret_cast<const float *>(smem + V2_RAW_OFF);
#pragma unroll
        for (int i = 0; i < 12; i++) {
          const int e = ctid + i * WS_COL_THREADS;
          const int rr = e / V2_SC, cc = e - rr * V2_SC;
          const int ry = min(max(b_tile - V2_HALO + rr, 0), A.src_h - 1) - (b_tile - V2_HALO);
          const int rx = min(max(a_tile - V2_HALO + cc, 0), A.src_w - 1) - (a_tile - V2_HALO);
          S[rr * V2_SP + cc] = A.dtype == SIFT_U8 ? v2_u8_over_255(raw8[ry * 64 + rx + 8]) : (double)raw32[ry * V2_SC + rx];
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < 12; i++) {
          const int e = ctid + i * WS_COL_THREADS;
          const int rr = e / V2_SC, cc = e - rr * V2_SC;
          const int gy = min(max(b_tile - V2_HALO + rr, 0), A.src_h - 1);          // clamp-to-edge, sift.js:116-119
          const int gx = min(max(a_tile - V2_HALO + cc, 0), A.src_w - 1);
          const char *row = (const char *)A.src + (size_t)gy * A.src_pitch;
          double v;
          switch (A.dtype) {
            case SIFT_U8: v = v2_u8_over_255(__ldg((const unsigned char *)row + gx)); break;
            case SIFT_F32: v = (double)__ldg((const float *)row + gx); break;
            case SIFT_F64: v = __ldg((const double *)row + gx); break;
            default: {
              const uchar4 c = __ldg((const uchar4 *)row + gx);
              const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)c.x, 0.299), __dmul_rn((double)c.y, 0.587)),
                                         __dmul_rn((double)c.z, 0.114));                          // image-utils.js:107
              v = g / 255.0;
            }
          }
          S[rr * V2_SP + cc] = v;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(WS_COL_THREADS) : "memory");     // S complete, the raw box is free
      if (A.use_tma && ctid == 0 && t + (int)gridDim.x < A.n_tiles) issue_source(t + gridDim.x);
      for (int L = 0; L < A.nlev; L++, n++) {          // octave 0 blurs every level from the base (background.js:110)
        const unsigned b = n & 1, k = n >> 1;
        if (k >= 1) v2_mbar_wait(empty0 + 8 * b, (k - 1) & 1);      // the row warps have read this buffer's previous level
        v2_column_pass(S, b ? Tv1 : Tv0, Wt, L, A.radius[L], ctid);
        __syncwarp();
        if ((ctid & 31) == 0) v2_mbar_arrive(full0 + 8 * b);
      }
    }
    return;
  }

  // ----------------------------------------------------------------------------------- row warps
  Oct0Thread T;
  T.y = tid & (V2_TR - 1);
  {
    const int gq = tid >> 6;                          // 16-column group: box gq >> 1, chunks 4 (gq & 1) .. + 3 of its row
    T.stage = v2_smem_u32(smem + V2_STAGE_OFF) + (gq >> 1) * V2_BOX_BYTES + T.y * 128;
    T.swz = (unsigned)((4 * (gq & 1)) ^ (T.y & 7));
  }
  const unsigned stage0 = v2_smem_u32(smem + V2_STAGE_OFF);
  double a0[8], a1[8], b0[8], b1[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { b0[k] = 0.0; b1[k] = 0.0; }
  unsigned n = 0;
  for (int t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    const int ty = t / A.tiles_x, tx = t - ty * A.tiles_x;
    const int a_tile = tx * V2_SW, b_tile = ty * V2_SH;
    T.x0 = 2 * a_tile + 16 * (tid >> 6);
    T.yg = 2 * b_tile + T.y;
    T.row_ok = T.yg < A.oct.h && T.x0 < A.oct.w;
    T.seed_lane = A.has_next && (T.yg & 1) == 0;
#pragma unroll 1
    for (int L = 0; L < A.nlev; L++, n++) {
      const unsigned b = n & 1, k = n >> 1;
      v2_mbar_wait(full0 + 8 * b, k & 1);
      // the two accumulator sets swap roles every level (this level's values are the next level's "previous")
      if (L & 1) v2_row_pass(A, b ? Tv1 : Tv0, Wt, L, A.radius[L], tid, T, empty0 + 8 * b, b, b0, b1, a0, a1);
      else v2_row_pass(A, b ? Tv1 : Tv0, Wt, L, A.radius[L], tid, T, empty0 + 8 * b, b, a0, a1, b0, b1);
      // level n-1's boxes (the other staging buffer) must have been READ by the TMA unit before level n+1 stages:
      // the issuing thread checks before the barrier, so that everyone leaves the barrier knowing it
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(WS_ROW_THREADS) : "memory");
      if (tid == 0) v2_store_level(A, stage0 + b * V2_STAGE_BYTES, L, 2 * a_tile, 2 * b_tile);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the boxes must outlive the last read
}

// ------------------------------------------------------------------ host side
typedef CUresult (*Oct0EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool oct0_encode_map(CUtensorMap *m, const void *src, int dtype, size_t pitch, int w, int h)
{
  static Oct0EncodeTiledFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
        q == cudaDriverEntryPointSuccess)
      encode = (Oct0EncodeTiledFn)fn;
  }
  if (!encode) return false;
  const size_t es = dtype == SIFT_U8 ? 1 : 4;
  if ((pitch % 16) != 0 || ((uintptr_t)src % 16) != 0 || (size_t)w * es > pitch) return false;
  const cuuint64_t gdim[2] = { (cuuint64_t)w, (cuuint64_t)h };
  const cuuint64_t gstride[1] = { (cuuint64_t)pitch };
  const cuuint32_t box[2] = { dtype == SIFT_U8 ? 64u : (cuuint32_t)V2_SC, V2_SR };
  const cuuint32_t estride[2] = { 1, 1 };
  return encode(m, dtype == SIFT_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)src, gdim,
                gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool oct0_v2_supported(const LevelPlan *plans, int nlev)
{
  // Off by default: measured on the B200 it does not beat the tile kernel of blur_fused.cu (0.152 vs 0.146 ms per
  // 1080p frame; profiles/r02_oct0_experiments.md has the numbers and the micro-benchmarks that explain them).
  // SIFT_B200_OCT0_WS=1 selects it; tests/test_fallback_paths.py holds it to the same parity bars.
  static const bool on = getenv("SIFT_B200_OCT0_WS") != nullptr;
  if (!on) return false;
  for (int s = 0; s < nlev; s++)
    if (plans[s].radius < 1 || plans[s].radius > V2_MAXR) return false;
  return true;
}

// The merged tap table is blur_fused.cu's (fused0_merge_taps: [nlev][24]{w0, w1}, zero padded).
// Tensor maps of the octave's output planes for the TMA stores: [nlev] Gaussian then [nlev - 1] DoG.
size_t oct0_out_map_bytes(int nlev) { return (size_t)(2 * nlev - 1) * sizeof(CUtensorMap); }

bool oct0_build_out_maps(const OctaveDev &oct, int nlev, void *h_maps)
{
  CUtensorMap probe;
  float dummy[4];
  (void)dummy;
  static Oct0EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return false;
    encode = (Oct0EncodeTiledFn)fn;
  }
  (void)probe;
  CUtensorMap *maps = (CUtensorMap *)h_maps;
  for (int i = 0; i < 2 * nlev - 1; i++) {
    float *plane = i < nlev ? oct.gauss[i] : oct.dog[i - nlev];
    const cuuint64_t gdim[2] = { (cuuint64_t)oct.w, (cuuint64_t)oct.h };
    const cuuint64_t gstride[1] = { (cuuint64_t)oct.pitch * sizeof(float) };
    const cuuint32_t box[2] = { 32, V2_TR };
    const cuuint32_t estride[2] = { 1, 1 };
    if (encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)plane, gdim, gstride, box, estride,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

void launch_oct0_v2(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                    const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                    int poly_woff, int nlev, int spo, int keep_gauss, const void *d_out_maps)
{
  Oct0Args A;
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < SIFT_MAX_LEVELS; s++) A.radius[s] = s < nlev ? plans[s].radius : 0;
  A.woff = poly_woff;
  A.out_maps = (const CUtensorMap *)d_out_maps;
  static const bool no_tma = getenv("SIFT_B200_NO_TMA") != nullptr || getenv("SIFT_B200_NO_TMA_BLUR") != nullptr;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  A.use_tma = (!no_tma && (dtype == SIFT_U8 || dtype == SIFT_F32) && oct0_encode_map(&tmap, src, dtype, src_pitch, src_w, src_h)) ? 1 : 0;
  A.tiles_x = (src_w + V2_SW - 1) / V2_SW;
  A.n_tiles = A.tiles_x * ((src_h + V2_SH - 1) / V2_SH);
  // persistent: one CTA per SM, tiles dealt round-robin
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  static int sm_count[64];
  if (dev >= 0 && dev < 64 && sm_count[dev] == 0) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
  sms = (dev >= 0 && dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  static const char *env_ctas = getenv("SIFT_B200_OCT0_CTAS");
  if (env_ctas && atoi(env_ctas) > 0) sms = atoi(env_ctas);
  const int grid = A.n_tiles < sms ? A.n_tiles : sms;
  const size_t smem = V2_SMEM_BYTES;
  cudaFuncSetAttribute(oct0_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  oct0_ws_kernel<<<grid, WS_THREADS, smem, st>>>(d_weights, A, tmap);
}
