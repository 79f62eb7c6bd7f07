// standalone check of the TMA window load used by fpm_update_kernel
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
struct P { CUtensorMap tmap; float2* out; int x, y, z, nr, ocp; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  float2* dst = reinterpret_cast<float2*>(smem);
  const uint32_t bytes = p.nr * p.ocp * 8;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&p.tmap)), "r"(p.x), "r"(p.y), "r"(p.z), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  asm volatile("{\n\t.reg .pred q;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int t = threadIdx.x; t < p.nr * p.ocp; t += blockDim.x) p.out[t] = dst[t];
}
int main() {
  const int L = 384, T = 3, NR = 35, OCP = 36;
  std::vector<float2> h((size_t)T * L * L);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)(i % 100003), (float)(i / 1000));
  float2 *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, NR * OCP * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s fn=%p q=%d\n", cudaGetErrorString(e), fn, (int)q);
  P p; 
  const cuuint64_t gdim[3] = {2 * L, L, T}; const cuuint64_t gstr[2] = {2 * L * 4, (cuuint64_t)2 * L * L * 4};
  const cuuint32_t box[3] = {2 * OCP, NR, 1}; const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((enc)fn)(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  p.out = o; p.x = 2 * 172; p.y = 160; p.z = 2; p.nr = NR; p.ocp = OCP;
  k<<<1, 128, NR * OCP * 8 + 1024>>>(p);
  e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float2> ho(NR * OCP); cudaMemcpy(ho.data(), o, NR * OCP * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < NR; ++r2) for (int c = 0; c < OCP; ++c) {
    float2 w = h[((size_t)2 * L + 160 + r2) * L + 172 + c]; float2 g = ho[r2 * OCP + c];
    if (w.x != g.x || w.y != g.y) ++bad;
  }
  printf("mismatches: %d\n", bad);
  return 0;
}
