// Minimal check: TMA 2D load of a u8 box with the tensor map passed as a __grid_constant__ kernel parameter
// (variant 0) or read from global memory (variant 1).  nvcc -arch=sm_100a tma_param_u8.cu -o tma_param_u8 -lcuda
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
struct Pad { int a[150]; };
__global__ void k(const __grid_constant__ Pad pad, const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, int variant, int x, int y, unsigned *out)
{
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(48u * 48u) : "memory");
    const CUtensorMap *m = variant == 0 ? &pm : gm;
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(m), "r"(b), "r"(x), "r"(y) : "memory");
  }
  __syncthreads();
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b), "r"(0u) : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < 48 * 48; i += blockDim.x) s += sm[i];
  atomicAdd(out, s + pad.a[0]);
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
  const int w = 1920, h = 1080;
  unsigned char *d; cudaMalloc(&d, w * h); cudaMemset(d, 1, w * h);
  void *fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m; memset(&m, 0, sizeof m);
  cuuint64_t gd[2] = { w, h }, gs[1] = { w };
  cuuint32_t box[2] = { 48, 48 }, es[2] = { 1, 1 };
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  CUtensorMap *gm; cudaMalloc(&gm, sizeof m); cudaMemcpy(gm, &m, sizeof m, cudaMemcpyHostToDevice);
  unsigned *out; cudaMalloc(&out, 4);
  Pad pad; memset(&pad, 0, sizeof pad);
  for (int variant = 1; variant >= 0; variant--)
    for (int x : { 64, -8 }) {
      cudaMemset(out, 0, 4);
      k<<<1, 256, 48 * 48>>>(pad, m, gm, variant, x, x, out);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned ho = 0; cudaMemcpy(&ho, out, 4, cudaMemcpyDeviceToHost);
      printf("variant %d x=%d: %s sum=%u (expect %d)\n", variant, x, cudaGetErrorString(e), ho, x < 0 ? 40 * 40 : 48 * 48);
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
