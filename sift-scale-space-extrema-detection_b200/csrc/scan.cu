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

__global__ void __launch_bounds__(SC_BX *SC_BY)
scan_octave_kernel(OctaveDev oct, int octave, int spo, double pix_threshold, int count_low,
                   sift_candidate *__restrict__ cand, int cand_cap, sift_candidate *__restrict__ low, int low_cap,
                   Counters *ctr)
{
  const int x = blockIdx.x * SC_BX + threadIdx.x;
  const int y = blockIdx.y * SC_BY + threadIdx.y;
  const bool inside = (x >= 1 && x < oct.w - 1 && y >= 1 && y < oct.h - 1);          // sift.js:221-222
  const size_t pitch = oct.pitch;
  for (int s = 1; s <= spo; s++) {                                                    // background.js:377
    bool hit = false, hit_low = false;
    float c = 0.f;
    if (inside) {
      c = oct.dog[s][(size_t)y * pitch + x];
      const bool strong = (double)fabsf(c) >= pix_threshold;                          // sift.js:294
      if (strong || count_low) {
        if (is_extremum<float>(oct.dog[s - 1], oct.dog[s], oct.dog[s + 1], pitch, x, y, c)) {
          hit = strong; hit_low = !strong;
        }
      }
    }
    int slot = warp_append(hit, &ctr->n_cand);
    if (slot >= 0 && slot < cand_cap) {
      sift_candidate r; r.octave = octave; r.scaleLevel = s; r.x = x; r.y = y; r.value = c; r.reserved0 = 0;
      cand[slot] = r;
    }
    if (count_low) {
      slot = warp_append(hit_low, &ctr->n_low);
      if (low && slot >= 0 && slot < low_cap) {
        sift_candidate r; r.octave = octave; r.scaleLevel = s; r.x = x; r.y = y; r.value = c; r.reserved0 = 0;
        low[slot] = r;
      }
    }
  }
}

void launch_scan_octave(cudaStream_t st, const OctaveDev &oct, int octave, int spo, double pix_threshold,
                          int count_low, sift_candidate *cand, int cand_cap, sift_candidate *low, int low_cap,
                          Counters *ctr)
{
  if (oct.w < 3 || oct.h < 3) return;
  dim3 block(SC_BX, SC_BY);
  dim3 grid((oct.w + SC_BX - 1) / SC_BX, (oct.h + SC_BY - 1) / SC_BY);
  scan_octave_kernel<<<grid, block, 0, st>>>(oct, octave, spo, pix_threshold, count_low, cand, cand_cap, low,
                                             low_cap, ctr);
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
  const int w = oc.w, h = oc.h;
  const size_t pitch = oc.pitch;
  const int x0 = tx * SA_TW + threadIdx.x * SA_PX;
  const int y = ty * SA_TH + threadIdx.y;
  const bool row_ok = (y >= 1 && y < h - 1) && x0 < w;                                // sift.js:221
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
          sift_candidate r; r.octave = o; r.scaleLevel = s; r.x = x0 + i; r.y = y; r.value = c[i]; r.reserved0 = 0;
          cand[slot] = r;
        }
      }
    }
    if (A.count_low && __any_sync(0xffffffffu, hit_low != 0)) {
#pragma unroll
      for (int i = 0; i < SA_PX; i++) {
        const int slot = warp_append((hit_low >> i) & 1u, &ctr->n_low);
        if (low && slot >= 0 && slot < low_cap) {
          sift_candidate r; r.octave = o; r.scaleLevel = s; r.x = x0 + i; r.y = y; r.value = c[i]; r.reserved0 = 0;
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
