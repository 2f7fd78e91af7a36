// scan.cu -- 3x3x3 strict extremum scan fused with the contrast pre-filter and a
// warp-aggregated compaction (ballot / popc, one atomic per warp per hit group).
//
// Restates SIFT_findExtremas (src/sift.js:212-316) over the DoG scales
// 1..nDoG-2 of an octave (background.js:374-377): a pixel is an extremum when all
// 26 neighbours are strictly greater (minimum, sift.js:261) or strictly smaller
// (maximum, sift.js:266); ties are never extrema.  abs(value) >= 0.8 * threshold
// (sift.js:285-294) sends it to the candidate list, otherwise to the low-contrast
// list (only materialised / counted on request: it feeds red UI markers only,
// background.js:408-413).
#include <cmath>
#include <cstring>
#include "common.cuh"

#define SC_BX 32
#define SC_BY 8

template <typename T>
__device__ __forceinline__ bool is_extremum(const T *__restrict__ p0, const T *__restrict__ p1,
                                            const T *__restrict__ p2, size_t pitch, int x, int y, T c)
{
  bool is_min = true, is_max = true;
  const T *pl[3] = { p0, p1, p2 };
#pragma unroll
  for (int p = 0; p < 3; p++) {
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const T *row = pl[p] + (size_t)(y + dy) * pitch + x;
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        if (p == 1 && dy == 0 && dx == 0) continue;
        const T v = row[dx];
        is_min = is_min && (v > c);
        is_max = is_max && (v < c);
      }
    }
    if (!(is_min || is_max)) return false;
  }
  return is_min || is_max;
}

// Warp-aggregated append: returns the slot for lanes with `hit`, -1 otherwise.
__device__ __forceinline__ int warp_append(bool hit, int *counter)
{
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (m == 0) return -1;
  const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return hit ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}


// ---- whole pyramid in one launch: 4 pixels per thread (one 16-byte load per scale), the 26
// neighbours are only fetched for pixels that pass the pre-filter, in-plane ring first.
#define SA_PX 4
#define SA_TW (32 * SA_PX)    // 128 pixels per tile row
#define SA_TH 8

struct ScanAllArgs {
  int n_oct, spo, count_low;
  int tile_start[SIFT_MAX_OCTAVES + 1];
  int tiles_x[SIFT_MAX_OCTAVES];
  double pix_threshold;
};

__device__ __forceinline__ bool is_extremum_ring_first(const float *__restrict__ p0, const float *__restrict__ p1,
                                                       const float *__restrict__ p2, size_t pitch, int x, int y,
                                                       float c)
{
  const float *r1 = p1 + (size_t)y * pitch + x;
  bool is_min = true, is_max = true;
  {
    const float a = r1[-1], b = r1[1];
    is_min = (a > c) && (b > c);
    is_max = (a < c) && (b < c);
    if (!(is_min || is_max)) return false;
  }
#pragma unroll
  for (int dy = -1; dy <= 1; dy += 2) {
    const float *row = r1 + (ptrdiff_t)dy * (ptrdiff_t)pitch;
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
      const float v = row[dx];
      is_min = is_min && (v > c);
      is_max = is_max && (v < c);
    }
    if (!(is_min || is_max)) return false;
  }
  const float *pl[2] = { p0, p2 };
#pragma unroll
  for (int p = 0; p < 2; p++) {
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const float *row = pl[p] + (size_t)(y + dy) * pitch + x;
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        const float v = row[dx];
        is_min = is_min && (v > c);
        is_max = is_max && (v < c);
      }
    }
    if (!(is_min || is_max)) return false;
  }
  return true;
}

__global__ void __launch_bounds__(32 * SA_TH)
scan_all_kernel(const OctaveDev *__restrict__ octs, ScanAllArgs A, sift_candidate *__restrict__ cand, int cand_cap,
                sift_candidate *__restrict__ low, int low_cap, Counters *ctr)
{
  // static indices only: a dynamically indexed kernel-parameter array is copied to local memory
  int o = 0, t0 = 0, ntx = A.tiles_x[0];
#pragma unroll
  for (int i = 1; i < SIFT_MAX_OCTAVES; i++)
    if (i < A.n_oct && (int)blockIdx.x >= A.tile_start[i]) { o = i; t0 = A.tile_start[i]; ntx = A.tiles_x[i]; }
  const int t = blockIdx.x - t0;
  const int ty = t / ntx, tx = t - ty * ntx;
  const OctaveDev &oc = octs[o];
  const int w = oc.w;
  const size_t pitch = oc.pitch;
  const int x0 = tx * SA_TW + threadIdx.x * SA_PX;
  const int y = ty * SA_TH + threadIdx.y;
  const int yg = y + oc.y_top;
  const bool row_ok = (yg >= 1 && yg < oc.gh - 1 && y >= oc.own0 && y < oc.own1) && x0 < w;   // sift.js:221
  for (int s = 1; s <= A.spo; s++) {                                                  // background.js:377
    const float *p1 = oc.dog[s];
    float c[SA_PX] = { 0.f, 0.f, 0.f, 0.f };
    if (row_ok) {
      const float4 v = *reinterpret_cast<const float4 *>(p1 + (size_t)y * pitch + x0);
      c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    }
    unsigned hit = 0, hit_low = 0;
#pragma unroll
    for (int i = 0; i < SA_PX; i++) {
      const int x = x0 + i;
      if (row_ok && x >= 1 && x < w - 1) {                                            // sift.js:222
        const bool strong = (double)fabsf(c[i]) >= A.pix_threshold;                   // sift.js:294
        if (strong || A.count_low) {
          if (is_extremum_ring_first(oc.dog[s - 1], p1, oc.dog[s + 1], pitch, x, y, c[i])) {
            if (strong) hit |= 1u << i; else hit_low |= 1u << i;
          }
        }
      }
    }
    if (__any_sync(0xffffffffu, hit != 0)) {
#pragma unroll
      for (int i = 0; i < SA_PX; i++) {
        const int slot = warp_append((hit >> i) & 1u, &ctr->n_cand);
        if (slot >= 0 && slot < cand_cap) {
          sift_candidate r; r.octave = o; r.scaleLevel = s; r.x = x0 + i; r.y = yg; r.value = c[i]; r.reserved0 = 0;
          cand[slot] = r;
        }
      }
    }
    if (A.count_low && __any_sync(0xffffffffu, hit_low != 0)) {
#pragma unroll
      for (int i = 0; i < SA_PX; i++) {
        const int slot = warp_append((hit_low >> i) & 1u, &ctr->n_low);
        if (low && slot >= 0 && slot < low_cap) {
          sift_candidate r; r.octave = o; r.scaleLevel = s; r.x = x0 + i; r.y = yg; r.value = c[i]; r.reserved0 = 0;
          low[slot] = r;
        }
      }
    }
  }
}

void launch_scan_all(cudaStream_t st, const OctaveDev *h_octs, const OctaveDev *d_octs, int n_oct, int spo,
                     double pix_threshold, int count_low, sift_candidate *cand, int cand_cap, sift_candidate *low,
                     int low_cap, Counters *ctr)
{
  ScanAllArgs A;
  A.n_oct = n_oct; A.spo = spo; A.count_low = count_low; A.pix_threshold = pix_threshold;
  int total = 0;
  for (int o = 0; o < n_oct; o++) {
    A.tile_start[o] = total;
    A.tiles_x[o] = (h_octs[o].w + SA_TW - 1) / SA_TW;
    total += A.tiles_x[o] * ((h_octs[o].h + SA_TH - 1) / SA_TH);
  }
  A.tile_start[n_oct] = total;
  if (total == 0) return;
  dim3 block(32, SA_TH);
  scan_all_kernel<<<total, block, 0, st>>>(d_octs, A, cand, cand_cap, low, low_cap, ctr);
}

// ---- whole pyramid in one launch, TMA-tiled -------------------------------------------------------
// One CTA = one 128 x 16 pixel tile of one octave.  An elected thread issues one TMA 2D tile load
// (cp.async.bulk.tensor) per DoG level -- box 136 x 18 floats: the tile plus the one-pixel ring the
// 26-neighbour test needs, widened to keep 16-byte alignment -- and the CTA waits on an mbarrier; out-of-
// range rows / columns are zero-filled by the TMA unit and never tested (sift.js:221-222 scans the
// interior only).  Every DoG value is read from HBM once per tile (+ halo) and all tests run from shared
// memory, branch-free: a thread owns 2 neighbouring pixels and marches down 8 rows keeping, per level, the
// horizontal 3-max / 3-min of the two previous rows in registers (see scan_tma_kernel for the test itself).
// Hits are compacted with ballot / popc and one atomic per warp.
#include <cuda.h>

#ifndef ST_PX
#define ST_PX 2                // pixels per thread along x (1 or 2)
#endif
#ifndef ST_TW
#define ST_TW 128
#endif
#ifndef ST_TH
#define ST_TH 16
#endif
#define ST_BW (ST_TW + 8)      // box: 4 columns left (alignment) + tile + 4 right
#define ST_BH (ST_TH + 2)
#define ST_PLANE ((ST_BH * ST_BW + 31) & ~31)   // floats per staged level: TMA destinations are 128-byte aligned
#ifndef ST_ROWS
#define ST_ROWS 8              // rows marched by one thread
#endif
#define ST_THREADS ((ST_TW / ST_PX) * (ST_TH / ST_ROWS))   // 128
#ifndef ST_CTAS_PER_SM
#define ST_CTAS_PER_SM 4
#endif

struct ScanTmaArgs {
  int n_oct, spo, ndog, count_low, total_tiles;
  int tile_start[SIFT_MAX_OCTAVES + 1];
  int tiles_x[SIFT_MAX_OCTAVES];
  int w[SIFT_MAX_OCTAVES], h[SIFT_MAX_OCTAVES];
  int y_top[SIFT_MAX_OCTAVES], gh[SIFT_MAX_OCTAVES], own0[SIFT_MAX_OCTAVES], own1[SIFT_MAX_OCTAVES];   // strips
  float thr_f;                   // smallest float >= the double threshold (see launch_scan_tma)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct ScanTile { int o, x_tile, y_tile, w, h, y_top, gh, own0, own1; };

__device__ __forceinline__ ScanTile scan_decode_tile(const ScanTmaArgs &A, int t)
{
  // A is a __grid_constant__ parameter: dynamic indices read the constant bank directly (no local copy)
  int o = 0;
  while (o + 1 < A.n_oct && t >= A.tile_start[o + 1]) o++;
  const int tt = t - A.tile_start[o], ntx = A.tiles_x[o];
  const int ty = tt / ntx, tx = tt - ty * ntx;
  ScanTile T;
  T.o = o; T.x_tile = tx * ST_TW; T.y_tile = ty * ST_TH; T.w = A.w[o]; T.h = A.h[o];
  T.y_top = A.y_top[o]; T.gh = A.gh[o]; T.own0 = A.own0[o]; T.own1 = A.own1[o];
  return T;
}

__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }   // one FMNMX3
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }

// One CTA = one tile.  Thread 0 decodes the tile, arms the mbarrier and issues one TMA box per DoG level; the
// four CTAs resident on an SM hide each other's load.  The min/max network runs on the half-rate ALU pipe, which
// bounds this kernel next to HBM, so the per-voxel test is kept off it: with m27 / n27 the max / min of the
// 3x3x3 block (centre included), t = (m27 - c) * (c - n27) is zero iff the centre equals one of them -- two
// FADD and one FMUL on the FMA pipe and a single compare.  Underflow can only produce false positives, and
// every positive (true extrema, ties, underflow) is settled by the exact strict 26-neighbour test.
template <int ND, bool COUNT_LOW>   // DoG levels per octave (spo + 2); also materialise the low-contrast list
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM)
scan_tma_kernel(const CUtensorMap *__restrict__ maps, const __grid_constant__ ScanTmaArgs A, sift_candidate *__restrict__ cand,
                int cand_cap, sift_candidate *__restrict__ low, int low_cap, Counters *ctr)
{
  extern __shared__ __align__(128) float tiles[];       // [ND][ST_PLANE], each level [ST_BH][ST_BW]
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  const ScanTile T = scan_decode_tile(A, blockIdx.x);

  if (tid == 0) {
    const unsigned b = smem_u32(&bar);
    const unsigned bytes = ND * ST_BH * ST_BW * sizeof(float);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
#pragma unroll
    for (int p = 0; p < ND; p++) {
      const CUtensorMap *m = maps + T.o * ND + p;
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
          ::"r"(smem_u32(tiles + p * ST_PLANE)), "l"(m), "r"(b), "r"(T.x_tile - 4), "r"(T.y_tile - 1)
          : "memory");
    }
  }
  __syncthreads();
  // thread -> ST_PX pixels (columns ST_PX*cx ...) x rows [8*ry, 8*ry + 8)
  const int cx = tid & (ST_TW / ST_PX - 1), ry = tid / (ST_TW / ST_PX);
  const int xl = ST_PX * cx;                              // tile-local column of pixel 0
  const int row_first = ry * ST_ROWS;                     // tile-local row of the first output row
  const int x0 = T.x_tile + xl;
  const float thr = A.thr_f;
  bool col_ok[ST_PX];                                                                  // sift.js:222
#pragma unroll
  for (int i = 0; i < ST_PX; i++) col_ok[i] = x0 + i >= 1 && x0 + i < T.w - 1;
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  const float *tile = tiles;
  // per level, three rolling rows (slot = row % 3, compile-time: the row loop advances by 3) of the horizontal
  // 3-max / 3-min, and the centre levels' own values.
  float hmax[3][ND][ST_PX], hmin[3][ND][ST_PX], cen[3][ND][ST_PX];
#pragma unroll 1
  for (int r3 = 0; r3 < ST_ROWS + 2; r3 += 3) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int r = r3 + j;
      if (r >= ST_ROWS + 2) break;
      const int cs = j, bs = (j + 2) % 3, as = (j + 1) % 3;          // slots of rows r, r-1, r-2
      // box row of tile-local row (row_first - 1 + r) is (row_first + r): the box starts one row above the tile
#pragma unroll
      for (int p = 0; p < ND; p++) {
        const float *rowp = tile + p * ST_PLANE + (row_first + r) * ST_BW + 4 + xl;
#if ST_PX == 2
        const float2 v = *reinterpret_cast<const float2 *>(rowp);
        const float L = rowp[-1], Rr = rowp[2];
        hmax[cs][p][0] = max3f(L, v.x, v.y); hmax[cs][p][1] = max3f(Rr, v.y, v.x);
        hmin[cs][p][0] = min3f(L, v.x, v.y); hmin[cs][p][1] = min3f(Rr, v.y, v.x);
        if (p >= 1 && p < ND - 1) { cen[cs][p][0] = v.x; cen[cs][p][1] = v.y; }
#else
        const float L = rowp[-1], v = rowp[0], Rr = rowp[1];
        hmax[cs][p][0] = max3f(L, v, Rr);
        hmin[cs][p][0] = min3f(L, v, Rr);
        if (p >= 1 && p < ND - 1) cen[cs][p][0] = v;
#endif
      }
      if (r >= 2) {
        float m9[ND][ST_PX], n9[ND][ST_PX];                  // 3x3 max / min per level (centre included)
#pragma unroll
        for (int p = 0; p < ND; p++)
#pragma unroll
          for (int i = 0; i < ST_PX; i++) {
            m9[p][i] = max3f(hmax[as][p][i], hmax[bs][p][i], hmax[cs][p][i]);
            n9[p][i] = min3f(hmin[as][p][i], hmin[bs][p][i], hmin[cs][p][i]);
          }
        float tz[ND][ST_PX];
        bool any = false;
#pragma unroll
        for (int s = 1; s < ND - 1; s++) {                   // background.js:377
#pragma unroll
          for (int i = 0; i < ST_PX; i++) {
            const float c = cen[bs][s][i];
            const float m27 = max3f(m9[s - 1][i], m9[s][i], m9[s + 1][i]);
            const float n27 = min3f(n9[s - 1][i], n9[s][i], n9[s + 1][i]);
            tz[s][i] = __fmul_rn(__fsub_rn(m27, c), __fsub_rn(c, n27));
            any = any || tz[s][i] == 0.f;
          }
        }
        if (__any_sync(0xffffffffu, any)) {                  // rare: a few per cent of the warp rows
          const int y = T.y_tile + row_first + r - 2;        // the middle row (slots as, bs, cs = rows y-1, y, y+1)
          const int yg = y + T.y_top;                        // row of the whole image (strips)
          const bool row_ok = yg >= 1 && yg < T.gh - 1 && y >= T.own0 && y < T.own1;   // sift.js:221
#pragma unroll
          for (int s = 1; s < ND - 1; s++)
#pragma unroll
            for (int i = 0; i < ST_PX; i++) {
              const float c = cen[bs][s][i];
              const bool strong = fabsf(c) >= thr;           // sift.js:294 (see launch_scan_tma)
              const bool maybe = tz[s][i] == 0.f && row_ok && col_ok[i] && (COUNT_LOW || strong);
              if (!__any_sync(0xffffffffu, maybe)) continue;
              bool e = false;
              if (maybe) {                                   // exact strict test of the 26 neighbours
                const float *ctr_px = tile + s * ST_PLANE + (row_first + r - 1) * ST_BW + 4 + xl + i;
                bool is_max = true, is_min = true;
#pragma unroll
                for (int dp = -1; dp <= 1; dp++)
#pragma unroll
                  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                      if (dp == 0 && dy == 0 && dx == 0) continue;
                      const float v = ctr_px[dp * ST_PLANE + dy * ST_BW + dx];
                      is_max = is_max && (v < c);            // sift.js:266
                      is_min = is_min && (v > c);            // sift.js:261
                    }
                e = is_max || is_min;
              }
              sift_candidate rec; rec.octave = T.o; rec.scaleLevel = s; rec.x = x0 + i; rec.y = yg; rec.value = c; rec.reserved0 = 0;
              int slot = warp_append(e && strong, &ctr->n_cand);
              if (slot >= 0 && slot < cand_cap) cand[slot] = rec;
              if (COUNT_LOW) {
                slot = warp_append(e && !strong, &ctr->n_low);
                if (low && slot >= 0 && slot < low_cap) low[slot] = rec;
              }
            }
        }
      }
    }
  }
}

// Host: one CUtensorMap per (octave, DoG level).  The driver entry point is fetched through the runtime
// (no link-time dependency on libcuda).  Returns 0 when TMA descriptors cannot be built.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

size_t scan_tma_map_bytes(int n_oct, int ndog) { return (size_t)n_oct * ndog * sizeof(CUtensorMap); }

bool scan_tma_supported(int ndog) { return ndog >= 3 && ndog <= 6; }

int scan_tma_build_maps(const OctaveDev *octs, int n_oct, int ndog, void *h_maps)
{
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return 0;
    encode = (EncodeTiledFn)fn;
  }
  CUtensorMap *maps = (CUtensorMap *)h_maps;
  for (int o = 0; o < n_oct; o++)
    for (int p = 0; p < ndog; p++) {
      const cuuint64_t gdim[2] = { (cuuint64_t)octs[o].w, (cuuint64_t)octs[o].h };
      const cuuint64_t gstride[1] = { (cuuint64_t)octs[o].pitch * sizeof(float) };
      const cuuint32_t box[2] = { ST_BW, ST_BH };
      const cuuint32_t estride[2] = { 1, 1 };
      if (encode(&maps[o * ndog + p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)octs[o].dog[p], gdim, gstride, box, estride,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    }
  return 1;
}

void launch_scan_tma(cudaStream_t st, const OctaveDev *h_octs, const void *d_maps, int n_oct, int spo,
                     double pix_threshold, int count_low, sift_candidate *cand, int cand_cap, sift_candidate *low,
                     int low_cap, Counters *ctr)
{
  ScanTmaArgs A;
  memset(&A, 0, sizeof A);
  A.n_oct = n_oct; A.spo = spo; A.ndog = spo + 2; A.count_low = count_low;
  // sift.js:294 compares abs(value) >= threshold in double.  The DoG values are floats, and for a float a,
  // (double)a >= T  <=>  a >= (the smallest float that is >= T): one float compare, exactly equivalent.
  float tf = (float)pix_threshold;
  if ((double)tf < pix_threshold) tf = nextafterf(tf, INFINITY);
  A.thr_f = tf;
  int total = 0;
  for (int o = 0; o < n_oct; o++) {
    A.tile_start[o] = total;
    A.tiles_x[o] = (h_octs[o].w + ST_TW - 1) / ST_TW;
    A.w[o] = h_octs[o].w; A.h[o] = h_octs[o].h;
    A.y_top[o] = h_octs[o].y_top; A.gh[o] = h_octs[o].gh; A.own0[o] = h_octs[o].own0; A.own1[o] = h_octs[o].own1;
    if (h_octs[o].w >= 3 && h_octs[o].h >= 3) total += A.tiles_x[o] * ((h_octs[o].h + ST_TH - 1) / ST_TH);
  }
  A.tile_start[n_oct] = total;
  A.total_tiles = total;
  if (total == 0) return;
  const int nd = spo + 2;
  const size_t smem = (size_t)nd * ST_PLANE * sizeof(float);
  const int grid = total;
  const CUtensorMap *maps = (const CUtensorMap *)d_maps;
#define LAUNCH_ND(N)                                                                                         \
  case N:                                                                                                    \
    if (count_low) {                                                                                         \
      cudaFuncSetAttribute(scan_tma_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      scan_tma_kernel<N, true><<<grid, ST_THREADS, smem, st>>>(maps, A, cand, cand_cap, low, low_cap, ctr);  \
    } else {                                                                                                 \
      cudaFuncSetAttribute(scan_tma_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      scan_tma_kernel<N, false><<<grid, ST_THREADS, smem, st>>>(maps, A, cand, cand_cap, low, low_cap, ctr); \
    }                                                                                                        \
    break;
  switch (nd) { LAUNCH_ND(3) LAUNCH_ND(4) LAUNCH_ND(5) LAUNCH_ND(6) }
#undef LAUNCH_ND
}

// ---- step function: SIFT_findExtremas on three Matrix2D (fp64) images ----------
__global__ void __launch_bounds__(SC_BX *SC_BY)
scan_f64_kernel(const double *__restrict__ d0, const double *__restrict__ d1, const double *__restrict__ d2,
                int rows, int cols, double pix_threshold, int32_t *__restrict__ cand_xy,
                double *__restrict__ cand_val, int cand_cap, int32_t *__restrict__ low_xy,
                double *__restrict__ low_val, int low_cap, int *counts)
{
  const int x = blockIdx.x * SC_BX + threadIdx.x;
  const int y = blockIdx.y * SC_BY + threadIdx.y;
  const bool inside = (x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1);
  bool hit = false, hit_low = false;
  double c = 0.0;
  if (inside) {
    c = d1[(size_t)y * cols + x];
    if (is_extremum<double>(d0, d1, d2, (size_t)cols, x, y, c)) {
      hit = fabs(c) >= pix_threshold;
      hit_low = !hit;
    }
  }
  int slot = warp_append(hit, &counts[0]);
  if (slot >= 0 && slot < cand_cap) { cand_xy[2 * slot] = x; cand_xy[2 * slot + 1] = y; cand_val[slot] = c; }
  slot = warp_append(hit_low, &counts[1]);
  if (slot >= 0 && slot < low_cap) { low_xy[2 * slot] = x; low_xy[2 * slot + 1] = y; low_val[slot] = c; }
}

void launch_scan_f64(cudaStream_t st, const double *d0, const double *d1, const double *d2, int rows, int cols,
                     double pix_threshold, int32_t *cand_xy, double *cand_val, int cand_cap,
                     int32_t *low_xy, double *low_val, int low_cap, int *counts)
{
  if (rows < 3 || cols < 3) return;
  dim3 block(SC_BX, SC_BY);
  dim3 grid((cols + SC_BX - 1) / SC_BX, (rows + SC_BY - 1) / SC_BY);
  scan_f64_kernel<<<grid, block, 0, st>>>(d0, d1, d2, rows, cols, pix_threshold, cand_xy, cand_val, cand_cap,
                                          low_xy, low_val, low_cap, counts);
}
