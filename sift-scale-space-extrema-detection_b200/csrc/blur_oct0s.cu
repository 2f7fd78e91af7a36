// blur_oct0s.cu -- octave 0 in one kernel, "small CTA" form: 64 x 32 output tiles, 256 threads of <= 64 registers,
// four CTAs (32 warps) per SM.
//
// Reference path restated (as blur_fused.cu / blur_oct0.cu): Matrix2D_linearResize(input, 0.5) (background.js:84,
// matrix2d.js:112-138), SIFT_blurMatrix2DChunk of that pixel-doubled base for every level (background.js:145-210,
// sift.js:72-149, clamp-to-edge sift.js:116-119), SIFT_subtractMatrix2DChunk between neighbouring levels
// (sift.js:154-188) and the rate-2.0 resize of level `spo` that seeds the next octave (background.js:114-130), in
// the polyphase form (R + 1 merged taps per output phase over SOURCE samples, both directions).
//
// Why small: tools/micro/dfma_warps.cu and dfma_operands.cu (profiles/r02_oct0_experiments.md) show that a tap loop
// keeps the fp64 pipe ~45 % busy with one warp per sub-partition, ~60 % with two and ~70 % with three or more (two
// fresh 64-bit register operands per DFMA cap it at 81 %), and that the fp64 -> fp32 conversions of the epilogue
// (XU pipe, a quarter of the DFMA rate) cost half as much time again unless other warps are in their tap loops
// meanwhile.  The 128-register tiles (16 outputs + 16 previous-level values per thread) allow 16 warps per SM, all
// walking through the same phases.  Here a thread owns 8 outputs (4 source positions x 2 phases): 16 + 16 + 8
// registers of state, so 32 warps fit, in four independent CTAs whose phases interleave.
//   column (y) pass first, on the un-doubled source columns: S -> Tv[32 rows][<= 48 columns], double-buffered;
//   row (x) pass last: a thread's 8 outputs are 32 contiguous bytes of one row, eight threads cover a 256-byte row
//   segment (two full lines), so plain 16-byte stores leave the SM as whole lines;
//   the previous level's unrounded values stay in registers (accumulator sets swap roles by level parity);
//   the source tile arrives as one TMA box for u8 / f32 images (clamp-to-edge by indexing in the conversion pass).
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "common.cuh"

#define OS_SW 32                       // source columns per tile
#define OS_SH 16                       // source rows per tile
#define OS_HALO 8                      // supports R <= 16
#define OS_SR (OS_SH + 2 * OS_HALO)    // 32 staged source rows (+ 1 slack row for the window prefetch)
#define OS_SC (OS_SW + 2 * OS_HALO)    // 48 staged source columns
#define OS_SP 49                       // S pitch (doubles)
#define OS_TP 49                       // Tv pitch (doubles): 48 columns + 1 slack
#define OS_TR (2 * OS_SH)              // 32 Tv rows (both row phases)
#define OS_THREADS 256
#define OS_MAXR 16
#define OS_WSTRIDE 24                  // merged tap table of blur_fused.cu: [nlev][24]{w0, w1}

#define OS_S_DOUBLES ((OS_SR + 1) * OS_SP)
#define OS_T_DOUBLES (OS_TR * OS_TP)
#define OS_T0_OFF ((OS_S_DOUBLES + 15) & ~15)
#define OS_T1_OFF (OS_T0_OFF + ((OS_T_DOUBLES + 15) & ~15))      // also the TMA landing zone of the raw source box
#define OS_W_OFF (OS_T1_OFF + ((OS_T_DOUBLES + 15) & ~15))

struct Oct0sArgs {
  const void *src;
  size_t src_pitch;
  int src_w, src_h, dtype;
  OctaveDev oct, next;
  int has_next, spo, keep_gauss, nlev;
  int radius[SIFT_MAX_LEVELS];
  int woff;
  int use_tma;
};

__device__ __forceinline__ unsigned os_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double os_u8_over_255(unsigned v)      // image-utils.js:114, exactly (see blur_fused.cu)
{
  const double r = 1.0 / 255.0;
  const double x = (double)v;
  const double q = x * r;
  return fma(fma(-q, 255.0, x), r, q);
}

// Two-phase sliding window over 4 neighbouring positions: a0[k] = sum_j w0[j] v[k+j], a1[k] = sum_j w1[j] v[k+j],
// v[p] = base[p * STRIDE].  The window rotates through 4 registers with compile-time indices; the first tap
// initialises the accumulators.  Positions up to np + 3 are read (np + 2 used).
#define OS_STEP(U, JJ)                                                                      \
  {                                                                                         \
    const double2 c = w2[(JJ)];                                                             \
    _Pragma("unroll") for (int k = 0; k < 4; k++) {                                         \
      a0[k] = fma(c.x, vw[(k + (U)) & 3], a0[k]);                                           \
      a1[k] = fma(c.y, vw[(k + (U)) & 3], a1[k]);                                           \
    }                                                                                       \
    vw[(U) & 3] = nxt[(JJ) * STRIDE];                                                       \
  }
template <int STRIDE>
__device__ __forceinline__ void os_window(const double *__restrict__ base, const double2 *__restrict__ w2, const int np,
                                          double (&a0)[4], double (&a1)[4])
{
  double vw[4];
#pragma unroll
  for (int k = 0; k < 4; k++) vw[k] = base[k * STRIDE];
  const double *nxt = base + 4 * STRIDE;
  {
    const double2 c = w2[0];
#pragma unroll
    for (int k = 0; k < 4; k++) { a0[k] = c.x * vw[k]; a1[k] = c.y * vw[k]; }
    vw[0] = nxt[0];
  }
  int j = 1;
  for (; j + 4 <= np; j += 4) { OS_STEP(1, j) OS_STEP(2, j + 1) OS_STEP(3, j + 2) OS_STEP(0, j + 3) }
  const int rem = np - j;                                    // 0..3, uniform over the CTA
  if (rem & 2) { OS_STEP(1, j) OS_STEP(2, j + 1) }
  if (rem & 1) {
    if (rem & 2) OS_STEP(3, j + 2)
    else OS_STEP(1, j)
  }
}

// Column pass of level s: one item = one staged column x 4 source rows x 2 phases -> Tv[2 r + phase][column]
__device__ __forceinline__ void os_column_pass(const double *__restrict__ S, double *__restrict__ Tv,
                                               const double *__restrict__ Wt, int s, int R, int tid)
{
  const int clo = -((R + 1) / 2);
  const int np = R + 1 + (R & 1);
  const int ncols = OS_SW + np - 1;                  // <= 48
  const int c_first = OS_HALO + clo;
  if (tid < 4 * ncols) {
    const int g = (tid >= ncols) + (tid >= 2 * ncols) + (tid >= 3 * ncols);
    const int cc = c_first + (tid - g * ncols);
    double a0[4], a1[4];
    os_window<OS_SP>(S + (OS_HALO + 4 * g + clo) * OS_SP + cc, reinterpret_cast<const double2 *>(Wt + s * 2 * OS_WSTRIDE), np, a0, a1);
    double *t = Tv + (8 * g) * OS_TP + cc;
#pragma unroll
    for (int k = 0; k < 4; k++) { t[(2 * k) * OS_TP] = a0[k]; t[(2 * k + 1) * OS_TP] = a1[k]; }
  }
}

struct Oct0sThread {
  int y, g, x0, yg;
  bool row_ok, full, seed_lane;
  size_t o_first;
};

// Row pass of level s + epilogue: G_s, D_{s-1} = G_{s-1} - G_s (sift.js:172) from the unrounded accumulators,
// pixel x0 + 2k + phase; the seed of the next octave at level spo.
__device__ __forceinline__ void os_row_pass(const Oct0sArgs &A, const double *__restrict__ Tv, const double *__restrict__ Wt,
                                            int s, int R, const Oct0sThread &T, double (&c0)[4], double (&c1)[4],
                                            const double (&p0)[4], const double (&p1)[4])
{
  const int clo = -((R + 1) / 2);
  const int np = R + 1 + (R & 1);
  os_window<1>(Tv + T.y * OS_TP + OS_HALO + 4 * T.g + clo, reinterpret_cast<const double2 *>(Wt + s * 2 * OS_WSTRIDE), np, c0, c1);
  if (!T.row_ok) return;
  float *gp = A.oct.gauss[s] + T.o_first;
  float *dp = A.oct.dog[s > 0 ? s - 1 : 0] + T.o_first;
  const bool wg = A.keep_gauss != 0, wd = s > 0;
  if (T.full) {
    if (wg) {
      reinterpret_cast<float4 *>(gp)[0] = make_float4((float)c0[0], (float)c1[0], (float)c0[1], (float)c1[1]);
      reinterpret_cast<float4 *>(gp)[1] = make_float4((float)c0[2], (float)c1[2], (float)c0[3], (float)c1[3]);
    }
    if (wd) {
      reinterpret_cast<float4 *>(dp)[0] = make_float4((float)(p0[0] - c0[0]), (float)(p1[0] - c1[0]), (float)(p0[1] - c0[1]), (float)(p1[1] - c1[1]));
      reinterpret_cast<float4 *>(dp)[1] = make_float4((float)(p0[2] - c0[2]), (float)(p1[2] - c1[2]), (float)(p0[3] - c0[3]), (float)(p1[3] - c1[3]));
    }
  } else {
    const int w = A.oct.w;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (T.x0 + 2 * k < w) {
        if (wg) gp[2 * k] = (float)c0[k];
        if (wd) dp[2 * k] = (float)(p0[k] - c0[k]);
      }
      if (T.x0 + 2 * k + 1 < w) {
        if (wg) gp[2 * k + 1] = (float)c1[k];
        if (wd) dp[2 * k + 1] = (float)(p1[k] - c1[k]);
      }
    }
  }
  if (s == A.spo && T.seed_lane) {                    // in[2a][2b] (matrix2d.js:129): even rows, even columns (phase 0)
    const int nr = (T.yg >> 1) + A.oct.seed_off;      // row of the next octave (strip-local)
    if (nr >= 0 && nr < A.next.h) {
      const int nc = T.x0 >> 1;
      double *sp = A.next.seed64 + (size_t)nr * A.next.w + nc;
      float *fp = A.next.gauss[0] + (size_t)nr * A.next.pitch + nc;
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (nc + k < A.next.w) { sp[k] = c0[k]; fp[k] = (float)c0[k]; }
    }
  }
}

__global__ void __launch_bounds__(OS_THREADS, 4)
oct0_small_kernel(const double *__restrict__ weights, const __grid_constant__ Oct0sArgs A, const __grid_constant__ CUtensorMap tmap)
{
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long bar;
  double *S = smem;
  double *Tv0 = smem + OS_T0_OFF;
  double *Tv1 = smem + OS_T1_OFF;
  double *Wt = smem + OS_W_OFF;
  const int tid = threadIdx.x;
  const int a_tile = blockIdx.x * OS_SW, b_tile = blockIdx.y * OS_SH;

  if (A.use_tma && tid == 0) {
    const unsigned b = os_smem_u32(&bar);
    // the box starts on a 16-byte boundary of its row: u8 tiles take 16 columns of left halo (box 64 x 32)
    const unsigned bytes = A.dtype == SIFT_U8 ? OS_SR * 64u : OS_SR * OS_SC * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(os_smem_u32(Tv1)), "l"(&tmap), "r"(b), "r"(a_tile - (A.dtype == SIFT_U8 ? 16 : OS_HALO)), "r"(b_tile - OS_HALO)
        : "memory");
  }
  for (int e = tid; e < A.nlev * 2 * OS_WSTRIDE; e += OS_THREADS) Wt[e] = __ldg(weights + A.woff + e);
  if (tid < OS_SP) S[OS_SR * OS_SP + tid] = 0.0;                                                  // slack row
  if (tid < OS_SR) S[tid * OS_SP + OS_SC] = 0.0;                                                  // pad column
  static_assert(OS_SR * OS_SC == 6 * OS_THREADS, "tile staging assumes 6 samples per thread");
  if (A.use_tma) {
    __syncthreads();
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(os_smem_u32(&bar)), "r"(0u) : "memory");
    }
    // raw box -> S with clamp-to-edge (sift.js:116-119) by indexing with clamped coordinates (always inside the box)
    const unsigned char *raw8 = reinterpret_cast<const unsigned char *>(Tv1);
    const float *raw32 = reinterpret_cast<const float *>(Tv1);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const int e = tid + i * OS_THREADS;
      const int rr = e / OS_SC, cc = e - rr * OS_SC;
      const int ry = min(max(b_tile - OS_HALO + rr, 0), A.src_h - 1) - (b_tile - OS_HALO);
      const int rx = min(max(a_tile - OS_HALO + cc, 0), A.src_w - 1) - (a_tile - OS_HALO);
      S[rr * OS_SP + cc] = A.dtype == SIFT_U8 ? os_u8_over_255(raw8[ry * 64 + rx + 8]) : (double)raw32[ry * OS_SC + rx];
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < 6; i++) {
      const int e = tid + i * OS_THREADS;
      const int rr = e / OS_SC, cc = e - rr * OS_SC;
      const int gy = min(max(b_tile - OS_HALO + rr, 0), A.src_h - 1);          // clamp-to-edge, sift.js:116-119
      const int gx = min(max(a_tile - OS_HALO + cc, 0), A.src_w - 1);
      const char *row = (const char *)A.src + (size_t)gy * A.src_pitch;
      double v;
      switch (A.dtype) {
        case SIFT_U8: v = os_u8_over_255(__ldg((const unsigned char *)row + gx)); break;
        case SIFT_F32: v = (double)__ldg((const float *)row + gx); break;
        case SIFT_F64: v = __ldg((const double *)row + gx); break;
        default: {
          const uchar4 c = __ldg((const uchar4 *)row + gx);
          const double g = __dadd_rn(__dadd_rn(__dmul_rn((double)c.x, 0.299), __dmul_rn((double)c.y, 0.587)),
                                     __dmul_rn((double)c.z, 0.114));                              // image-utils.js:107
          v = g / 255.0;
        }
      }
      S[rr * OS_SP + cc] = v;
    }
  }
  __syncthreads();

  Oct0sThread T;
  T.y = tid >> 3;                                     // tile row 0..31
  T.g = tid & 7;                                      // 4 source positions = 8 output columns
  T.x0 = 2 * a_tile + 8 * T.g;
  T.yg = 2 * b_tile + T.y;
  T.row_ok = T.yg < A.oct.h && T.x0 < A.oct.w;
  T.full = T.x0 + 8 <= A.oct.w;
  T.seed_lane = A.has_next && (T.yg & 1) == 0;
  T.o_first = (size_t)T.yg * A.oct.pitch + T.x0;

  // Step L: row pass (+ epilogue) of level L-1 from Tv[(L-1) & 1], then the column pass of level L into Tv[L & 1],
  // then one barrier.  Octave 0 blurs every level from the base (background.js:110).
  double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { b0[k] = 0.0; b1[k] = 0.0; }
#pragma unroll 1
  for (int L = 0; L <= A.nlev; L++) {
    if (L >= 1) {
      const double *Tr = ((L - 1) & 1) ? Tv1 : Tv0;
      if (L & 1) os_row_pass(A, Tr, Wt, L - 1, A.radius[L - 1], T, a0, a1, b0, b1);
      else os_row_pass(A, Tr, Wt, L - 1, A.radius[L - 1], T, b0, b1, a0, a1);
    }
    if (L < A.nlev) os_column_pass(S, (L & 1) ? Tv1 : Tv0, Wt, L, A.radius[L], tid);
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*Oct0sEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool oct0s_encode_map(CUtensorMap *m, const void *src, int dtype, size_t pitch, int w, int h)
{
  static Oct0sEncodeTiledFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
        q == cudaDriverEntryPointSuccess)
      encode = (Oct0sEncodeTiledFn)fn;
  }
  if (!encode) return false;
  const size_t es = dtype == SIFT_U8 ? 1 : 4;
  if ((pitch % 16) != 0 || ((uintptr_t)src % 16) != 0 || (size_t)w * es > pitch) return false;
  const cuuint64_t gdim[2] = { (cuuint64_t)w, (cuuint64_t)h };
  const cuuint64_t gstride[1] = { (cuuint64_t)pitch };
  const cuuint32_t box[2] = { dtype == SIFT_U8 ? 64u : (cuuint32_t)OS_SC, OS_SR };
  const cuuint32_t estride[2] = { 1, 1 };
  return encode(m, dtype == SIFT_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)src, gdim,
                gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool oct0_small_supported(const LevelPlan *plans, int nlev)
{
  // Off by default: 0.186 ms per 1080p frame against 0.146 for the tile kernel of blur_fused.cu -- 32 warps, but the
  // 4-position window needs a sample load per 8 DFMA and tops out near 65 % of the fp64 peak, and the kernel issues
  // 92 M warp instructions against 78 M (profiles/r02_oct0_experiments.md).  SIFT_B200_OCT0_SMALL=1 selects it.
  static const bool on = getenv("SIFT_B200_OCT0_SMALL") != nullptr;
  if (!on) return false;
  for (int s = 0; s < nlev; s++)
    if (plans[s].radius < 1 || plans[s].radius > OS_MAXR) return false;
  return true;
}

void launch_oct0_small(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                       const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                       int poly_woff, int nlev, int spo, int keep_gauss)
{
  Oct0sArgs A;
  A.src = src; A.src_pitch = src_pitch; A.src_w = src_w; A.src_h = src_h; A.dtype = dtype;
  A.oct = oct; A.next = next ? *next : oct; A.has_next = next ? 1 : 0;
  A.spo = spo; A.keep_gauss = keep_gauss; A.nlev = nlev;
  for (int s = 0; s < SIFT_MAX_LEVELS; s++) A.radius[s] = s < nlev ? plans[s].radius : 0;
  A.woff = poly_woff;
  static const bool no_tma = getenv("SIFT_B200_NO_TMA") != nullptr || getenv("SIFT_B200_NO_TMA_BLUR") != nullptr;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  A.use_tma = (!no_tma && (dtype == SIFT_U8 || dtype == SIFT_F32) && oct0s_encode_map(&tmap, src, dtype, src_pitch, src_w, src_h)) ? 1 : 0;
  dim3 grid((src_w + OS_SW - 1) / OS_SW, (src_h + OS_SH - 1) / OS_SH);
  const size_t smem = (size_t)(OS_W_OFF + nlev * 2 * OS_WSTRIDE) * sizeof(double);
  cudaFuncSetAttribute(oct0_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  oct0_small_kernel<<<grid, OS_THREADS, smem, st>>>(d_weights, A, tmap);
}
