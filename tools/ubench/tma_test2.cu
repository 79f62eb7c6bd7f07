#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// mode 0: mbarrier only; 1: 3-D TMA with tensor map as own param; 2: 2-D TMA; 3: 1-D bulk copy (no tensor map)
__global__ void k(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2, float2* src, float2* out, int mode, int x, int y, int z, int nr, int ocp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  float2* dst = reinterpret_cast<float2*>(smem);
  const uint32_t bytes = nr * ocp * 8;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    } else if (mode == 1) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    } else if (mode == 2) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap2)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    } else {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dst)), "l"(src), "r"(1024), "r"(smem_u32(&bar)) : "memory");
    }
  }
  asm volatile("{\n\t.reg .pred q;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int t = threadIdx.x; t < nr * ocp; t += blockDim.x) out[t] = dst[t];
}
#include <cstdlib>
int main(int argc, char** argv) {
  const int L = 384, T = 3, NR = 35, OCP = 36;
  std::vector<float2> h((size_t)T * L * L);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)(i % 100003), (float)(i / 1000));
  float2 *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, NR * OCP * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m3, m2;
  { const cuuint64_t gdim[3] = {2 * L, L, T}; const cuuint64_t gstr[2] = {2 * L * 4, (cuuint64_t)2 * L * L * 4};
    const cuuint32_t box[3] = {2 * OCP, NR, 1}; const cuuint32_t es[3] = {1, 1, 1};
    printf("enc3 %d\n", (int)((enc)fn)(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  { const cuuint64_t gdim[2] = {2 * L, (cuuint64_t)L * T}; const cuuint64_t gstr[1] = {2 * L * 4};
    const cuuint32_t box[2] = {2 * OCP, NR}; const cuuint32_t es[2] = {1, 1};
    printf("enc2 %d\n", (int)((enc)fn)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  int m0 = 0, m1 = 4;
  if (argc > 1) { m0 = atoi(argv[1]); m1 = m0 + 1; }
  for (int mode = m0; mode < m1; ++mode) {
    k<<<1, 128, NR * OCP * 8 + 1024>>>(m3, m2, d, o, mode, 2 * 172, mode == 2 ? 2 * L + 160 : 160, 2, NR, OCP);
    cudaError_t e = cudaDeviceSynchronize(); printf("mode %d kernel: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) break;
    if (mode == 1 || mode == 2) {
      std::vector<float2> ho(NR * OCP); cudaMemcpy(ho.data(), o, NR * OCP * 8, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < NR; ++r2) for (int c = 0; c < OCP; ++c) { float2 w = h[((size_t)2 * L + 160 + r2) * L + 172 + c]; float2 g = ho[r2 * OCP + c]; if (w.x != g.x || w.y != g.y) ++bad; }
      printf("  mismatches: %d\n", bad);
    }
  }
  return 0;
}
