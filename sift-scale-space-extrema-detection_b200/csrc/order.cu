// order.cu -- device-side ordering of the keypoint records into the reference's output order.
//
// The reference emits keypoints in candidate order: octave, DoG scale, row, column (background.js:468-471
// iterates octaves / scales, sift.js:221-222 scans rows then columns); refinement here appends them in
// whatever order the threads finish.  sift_detect (host output) orders on the host while the GPU works on the
// next frames; sift_detect_device(..., ordered = 1) orders on the device: one key per record, a radix sort of
// (key, index) pairs -- cub::DeviceRadixSort, library code, not on the measured hot path -- and a gather.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

__global__ void order_keys_kernel(const sift_keypoint *__restrict__ kp, const Counters *__restrict__ ctr, int n_sort,
                                  unsigned long long *__restrict__ keys, unsigned *__restrict__ idx)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sort) return;
  const int n = min(ctr->n_kp, n_sort);
  unsigned long long k = ~0ull;                        // padding sorts last
  if (i < n) {
    const sift_keypoint r = kp[i];
    k = ((unsigned long long)(unsigned)r.octave << 58) | ((unsigned long long)(unsigned)r.candScale << 52) |
        ((unsigned long long)(unsigned)r.candY << 26) | (unsigned long long)(unsigned)r.candX;
  }
  keys[i] = k;
  idx[i] = (unsigned)i;
}

__global__ void order_gather_kernel(const sift_keypoint *__restrict__ kp, const Counters *__restrict__ ctr, int n_sort,
                                    const unsigned *__restrict__ idx, sift_keypoint *__restrict__ out, int cap, int *d_count)
{
  const int n = min(min(ctr->n_kp, n_sort), cap);
  // 80-byte records = 5 x 16 bytes: five lanes per record
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int rec = t / 5, part = t - rec * 5;
  if (rec < n) reinterpret_cast<uint4 *>(out + rec)[part] = reinterpret_cast<const uint4 *>(kp + idx[rec])[part];
  if (t == 0 && d_count) *d_count = ctr->n_kp;
}

size_t order_scratch_bytes(int n_sort)
{
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const unsigned *)nullptr, (unsigned *)nullptr, n_sort, 0, 62);
  const size_t n = (size_t)n_sort;
  return 2 * n * sizeof(unsigned long long) + 2 * n * sizeof(unsigned) + ((temp + 255) & ~(size_t)255) + 1024;
}

// kp: unordered records (n_sort slots readable); out: cap ordered records.  Returns the number of kernel launches.
int launch_order_keypoints(cudaStream_t st, const sift_keypoint *kp, const Counters *ctr, int n_sort, void *scratch,
                           size_t scratch_bytes, sift_keypoint *out, int cap, int *d_count)
{
  const size_t n = (size_t)n_sort;
  char *p = (char *)scratch;
  unsigned long long *k0 = (unsigned long long *)p; p += n * sizeof(unsigned long long);
  unsigned long long *k1 = (unsigned long long *)p; p += n * sizeof(unsigned long long);
  unsigned *i0 = (unsigned *)p; p += n * sizeof(unsigned);
  unsigned *i1 = (unsigned *)p; p += n * sizeof(unsigned);
  p = (char *)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  size_t temp = scratch_bytes - (size_t)(p - (char *)scratch);
  order_keys_kernel<<<(n_sort + 255) / 256, 256, 0, st>>>(kp, ctr, n_sort, k0, i0);
  cub::DeviceRadixSort::SortPairs(p, temp, k0, k1, i0, i1, n_sort, 0, 62, st);
  const long long threads = 5LL * n_sort;
  order_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(kp, ctr, n_sort, i1, out, cap, d_count);
  return 2 + 6;
}
