#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int x, int y, int n) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  float* dst = reinterpret_cast<float*>(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred q;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = dst[t];
}
int main(int argc, char** argv) {
  const int W = 768, Hh = 1152;
  int bx = argc > 1 ? atoi(argv[1]) : 64, by = argc > 2 ? atoi(argv[2]) : 32, x = argc > 3 ? atoi(argv[3]) : 0, y = argc > 4 ? atoi(argv[4]) : 0;
  std::vector<float> h((size_t)W * Hh);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, bx * by * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)Hh}; const cuuint64_t gstr[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by}; const cuuint32_t es[2] = {1, 1};
  int r = (int)((enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d at (%d,%d) enc %d ", bx, by, x, y, r);
  k<<<1, 128, bx * by * 4 + 1024>>>(m, o, x, y, bx * by);
  cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> ho(bx * by); cudaMemcpy(ho.data(), o, bx * by * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < by; ++r2) for (int c = 0; c < bx; ++c) if (ho[r2 * bx + c] != h[(size_t)(y + r2) * W + x + c]) ++bad;
    printf("mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
