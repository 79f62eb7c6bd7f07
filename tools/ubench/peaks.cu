// peaks.cu -- measures the two on-chip ceilings the update kernels are judged against:
//   FP32 FMA throughput (plain FFMA and the packed FFMA2 of sm_100) and shared-memory load bandwidth (LDS.128,
//   conflict-free), whole GPU, best of several launches, CUDA events.  Prints one JSON line; bench.py runs it on the
//   box right before the timed legs and uses the numbers as the roofline denominators (a nominal
//   148 SM x 128 lanes x 2 x f_max is only the fallback).  Built by fpm-opencv_b200/Makefile into bin/fpm_peaks.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(1024) ffma_kernel(float* sink, float x, float y) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(1024) ffma2_kernel(float* sink, float x, float y) {
  float2 a[8];
  const float2 xx = make_float2(x, x), yy = make_float2(y, y);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], xx, yy);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(1024) lds_kernel(float* sink) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {            // consecutive lanes read consecutive 16-byte words: conflict-free LDS.128
      const float4 v = sm[(idx + i * 1024) & 4095];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + 32) & 4095;
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

template <typename F> static float best_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  if (cudaSetDevice(dev) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  float* sink; cudaMalloc(&sink, 64);
  const int ctas = p.multiProcessorCount * 2 * 8;       // 2 resident CTAs of 1024 threads per SM, 8 waves
  const double threads = (double)ctas * 1024;
  const float t1 = best_ms([&] { ffma_kernel<<<ctas, 1024>>>(sink, 1.0001f, 0.5f); });
  const float t2 = best_ms([&] { ffma2_kernel<<<ctas, 1024>>>(sink, 1.0001f, 0.5f); });
  cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const float t3 = best_ms([&] { lds_kernel<<<ctas, 1024, 65536>>>(sink); });
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  const double ffma = threads * 8 * ITERS * 2 / (t1 * 1e-3) / 1e12;
  const double ffma2 = threads * 16 * ITERS * 2 / (t2 * 1e-3) / 1e12;
  const double lds = threads * ITERS * 16 / (t3 * 1e-3) / 1e12;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("{\"fp32_ffma_tflops\": %.2f, \"fp32_ffma2_tflops\": %.2f, \"smem_lds128_tbs\": %.2f, \"sm_count\": %d, "
         "\"clock_rate_mhz\": %.0f, \"how\": \"8 independent FFMA / FFMA2 chains per thread, 2x1024 threads per SM, %d iterations; "
         "conflict-free LDS.128 from 64 KB; best of 5 launches, CUDA events\"}\n",
         ffma, ffma2, lds, p.multiProcessorCount, clk / 1000.0, ITERS);
  return 0;
}
