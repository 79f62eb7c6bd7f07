// micro-benchmarks of warp primitives on sm_100a (single warp, dependent chains)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, float* sink) {
  __shared__ float sm[1024];
  int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float v = lane * 1.5f; unsigned u = lane;
  long long t0, t1;
  const int R = 64;
  // 1. dependent REDUX.MAX chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) u = __reduce_max_sync(0xffffffffu, u + lane);
  t1 = clock64(); if (threadIdx.x == 0) out[0] = (t1 - t0) / R;
  // 2. dependent SHFL xor chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)) + 1.f;
  t1 = clock64(); if (threadIdx.x == 0) out[1] = (t1 - t0) / R;
  // 3. dependent LDS chain
  int idx = lane;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) idx = ((int)sm[idx]) & 1023;
  t1 = clock64(); if (threadIdx.x == 0) out[2] = (t1 - t0) / R;
  // 4. STS + syncwarp + LDS chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) { sm[lane] = v; __syncwarp(); v = sm[(lane + 1) & 31] + 1.f; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) out[3] = (t1 - t0) / R;
  // 5. dependent FFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) v = fmaf(v, 1.0001f, 0.5f);
  t1 = clock64(); if (threadIdx.x == 0) out[4] = (t1 - t0) / R;
  // 6. __syncthreads with all warps
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) out[5] = (t1 - t0) / R;
  // 7. MUFU rsqrt chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); v = r + 1.f; }
  t1 = clock64(); if (threadIdx.x == 0) out[6] = (t1 - t0) / R;
  // 8. integer division chain
  int q = 1000003 + lane, dd = 35 + (lane & 1);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) q = q / dd + 1000003;
  t1 = clock64(); if (threadIdx.x == 0) out[7] = (t1 - t0) / R;
  // 9. half-mask REDUX
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) u = __reduce_max_sync((lane & 16) ? 0xffff0000u : 0x0000ffffu, u + lane);
  t1 = clock64(); if (threadIdx.x == 0) out[8] = (t1 - t0) / R;
  // 10. atomicMax shared same address all lanes
  unsigned* us = reinterpret_cast<unsigned*>(sm);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < R; ++i) atomicMax(&us[(lane >> 4)], u + i);
  __syncwarp();
  t1 = clock64(); if (threadIdx.x == 0) out[9] = (t1 - t0) / R;
  sink[threadIdx.x] = v + u + idx + q + sm[lane];
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 128); cudaMalloc(&s, 4096);
  for (int nt : {32, 512}) {
    cudaMemset(d, 0, 128);
    k<<<1, nt>>>(d, s); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("threads %3d: REDUX %lld | SHFL+max+add %lld | LDS chain %lld | STS+syncwarp+LDS %lld | FFMA %lld | syncthreads %lld | MUFU+add %lld | idiv %lld | halfmask REDUX %lld | ATOMS.max %lld\n",
           nt, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  }
  return 0;
}
