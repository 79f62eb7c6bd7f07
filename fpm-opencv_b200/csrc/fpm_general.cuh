// fpm_general.cuh -- the sub-aperture update for tile sizes the fused kernels do not cover: any even Np whose
// prime factors are 2, 3, 5 (the shipped dataset*.json use cropSizeX = 90, 100 and 200).
//
// Same arithmetic and conventions as fpm_update.cuh (one body of fpmMain.cpp:350-475 per update, centred
// spectrum, exact max|objF| from a grid of cell maxima), but unfused: one elementwise kernel per step of the
// loop body, batched over tiles (blockIdx.y), with line_fft_kernel (mixed radix 2/3/4/5) for the transforms.
// The field makes a round trip through L2 between the kernels; this is the coverage path, not the fast one.
//
//   gen_window_mul    Phi = O * P                          fpmMain.cpp:358-364
//   (IFFT rows, cols; scaled by 1/N^2)                     :365
//   gen_amplitude     psi' = sqrt(I) psi / |psi + eps|     :378-393
//   (FFT rows, cols)                                       :394
//   gen_object_update O += dPhi |P| P* / D_O ; Q = dPhi |O| O* / ((|O|^2+d1) + i k d1) * S     :406-447,459-472
//   gen_cells_update  cell maxima of |objF|^2 for the cells the window touches                  :460
//   gen_cells_max     max|objF|^2 over the grid                                                 :467
//   gen_pupil_update  P += Q / max|objF| ; max|P|^2 for the next update                         :470-475,415
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_regs.cuh"
#include "fpm_update.cuh"

namespace fpm {

struct GeneralParams {
  float2* objFc;            // [n_tiles][L][L] centred
  float2* pupil;            // [n_tiles][N][N] DC-at-corner
  const float* stack;       // [n_tiles][n_leds][N][N] 1/I, natural order
  const float* support;     // [N][N]
  const short2* crop;       // [n_leds]
  float2* field;            // [n_tiles][N][N]
  float2* q;                // [n_tiles][N][N]
  float* cells;             // [n_tiles][cgr][cgc] maxima of |objFc|^2 over 16x16-pixel cells (edge cells partial)
  float* scal;              // [n_tiles][4]: 0 = max|P|^2, 1 = max|objF|^2
  int N, L, n_leds, tile0, slot;
  int cgr, cgc;
  int ylo, xlo, nrb, ncb;   // bounding box of the pupil support: first wrapped row / column (in [-N/2, N/2)), extent
  float delta1, delta2, eps, kappa;
  int apply;                // gen_pupil_update: 0 = only reduce max|P|^2 (start of a launch sequence)
  int stack_r1;             // stack layout (stack_pos_offset): 0 = natural
};

__device__ __forceinline__ int wrap_half(int i, int N) { return (i < N / 2) ? i : i - N; }
// Stack layouts of the general path: R1 == 0 natural order [y][x]; R1 > 0 position-major [pos(y)][pos(x)] with
// pos(v) = R2*(v % R1) + v / R1, R2 = N / R1 -- the order in which the two-stage in-place transforms of
// fpm_pruned_fused.cuh leave the field, so that its amplitude stage reads 1/I coalesced.
__host__ __device__ __forceinline__ int stack_pos_offset(int y, int x, int N, int R1) {
  if (R1 == 0) return y * N + x;
  const int R2 = N / R1;
  return (R2 * (y % R1) + y / R1) * N + R2 * (x % R1) + x / R1;
}

__global__ void __launch_bounds__(256) gen_window_mul(const GeneralParams p) {
  const int tile = p.tile0 + blockIdx.y, N = p.N, L = p.L, H = N / 2;
  const short2 cr = p.crop[p.slot];
  const float2* O = p.objFc + (size_t)tile * L * L + (size_t)(cr.y + H) * L + (cr.x + H);
  const float2* P = p.pupil + (size_t)tile * N * N;
  float2* F = p.field + (size_t)tile * N * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x) {
    const int i = t / N, j = t - i * N;
    const int iw = wrap_half(i, N), jw = wrap_half(j, N);
    const bool in = (unsigned)(iw - p.ylo) < (unsigned)p.nrb && (unsigned)(jw - p.xlo) < (unsigned)p.ncb;
    F[t] = in ? cmul(O[iw * L + jw], P[t]) : make_float2(0.f, 0.f);            // P = 0 outside the support's box
  }
}

__global__ void __launch_bounds__(256) gen_amplitude(const GeneralParams p) {
  const int tile = p.tile0 + blockIdx.y, N = p.N;
  const float* inv_i = p.stack + ((size_t)tile * p.n_leds + p.slot) * N * N;
  float2* F = p.field + (size_t)tile * N * N;
  const float er = p.eps, ei = p.kappa * p.eps;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x) {
    const float2 v = F[t];
    const float2 tt = make_float2(v.x + er, v.y + ei);
    const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * inv_i[stack_pos_offset(t / N, t % N, N, p.stack_r1)]);   // sqrt(I)/|psi+eps|; I = 0 -> 0
    F[t] = make_float2(v.x * sc, v.y * sc);
  }
}

__global__ void __launch_bounds__(256) gen_object_update(const GeneralParams p) {
  const int tile = p.tile0 + blockIdx.y, N = p.N, L = p.L, H = N / 2;
  const short2 cr = p.crop[p.slot];
  float2* O = p.objFc + (size_t)tile * L * L + (size_t)(cr.y + H) * L + (cr.x + H);
  const float2* P = p.pupil + (size_t)tile * N * N;
  const float2* F = p.field + (size_t)tile * N * N;
  float2* Q = p.q + (size_t)tile * N * N;
  const float inv_pmax = rsqrt_fast(p.scal[(size_t)tile * 4 + 0]);
  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < p.nrb * p.ncb; b += gridDim.x * blockDim.x) {
    const int bi = b / p.ncb, iw = p.ylo + bi, jw = p.xlo + (b - bi * p.ncb);          // the box: nothing changes outside
    const int t = (iw < 0 ? iw + N : iw) * N + (jw < 0 ? jw + N : jw);
    float2* op = O + iw * L + jw;
    const float2 Ov = *op, Pv = P[t];
    const float2 d = csub(F[t], cmul(Ov, Pv));
    const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
    const float2 num = cmulc(d, Pv);
    const float A = pa2 + p.delta2;
    const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
    *op = make_float2(Ov.x + (num.x * A + num.y * kd2) * sc, Ov.y + (num.y * A - num.x * kd2) * sc);
    const float oa2 = fmaf(Ov.x, Ov.x, Ov.y * Ov.y);
    const float2 numq = cmulc(d, Ov);
    const float A1 = oa2 + p.delta1;
    const float sq = sqrt_fast(oa2) * p.support[t] * rcp_fast(fmaf(A1, A1, kd1 * kd1));
    Q[t] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
  }
}

// One CTA (16x16 threads) per cell.  mode 0: the cells touched by the window of p.slot (blockIdx.x enumerates
// them); mode 1: every cell of the grid (start of a launch sequence).
__global__ void __launch_bounds__(256) gen_cells_update(const GeneralParams p, int all) {
  const int tile = p.tile0 + blockIdx.y, L = p.L;
  int a, b;
  if (all) {
    a = blockIdx.x / p.cgc; b = blockIdx.x % p.cgc;
  } else {
    const short2 cr = p.crop[p.slot];
    const int H = p.N / 2;                                          // the part of the window inside the box
    const int a0 = (cr.y + H + p.ylo) >> 4, a1 = (cr.y + H + p.ylo + p.nrb - 1) >> 4;
    const int b0 = (cr.x + H + p.xlo) >> 4, b1 = (cr.x + H + p.xlo + p.ncb - 1) >> 4;
    const int nb = b1 - b0 + 1;
    a = a0 + blockIdx.x / nb; b = b0 + blockIdx.x % nb;
    if (a > a1) return;
  }
  const int r = (a << 4) + (threadIdx.x >> 4), c = (b << 4) + (threadIdx.x & 15);
  float m = 0.f;
  if (r < L && c < L) {
    const float2 o = p.objFc[(size_t)tile * L * L + (size_t)r * L + c];
    m = fmaf(o.x, o.x, o.y * o.y);
  }
  __shared__ float red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    p.cells[(size_t)tile * p.cgr * p.cgc + a * p.cgc + b] = m;
  }
}

// one CTA per tile: max over the cell grid -> scal[1]; max|P|^2 (consumed by gen_object_update) is cleared for
// the accumulation in gen_pupil_update
__global__ void __launch_bounds__(256) gen_cells_max(const GeneralParams p) {
  const int tile = p.tile0 + blockIdx.x;
  const float* U = p.cells + (size_t)tile * p.cgr * p.cgc;
  float m = 0.f;
  for (int t = threadIdx.x; t < p.cgr * p.cgc; t += blockDim.x) m = fmaxf(m, U[t]);
  __shared__ float red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    p.scal[(size_t)tile * 4 + 1] = m;
    if (p.apply) p.scal[(size_t)tile * 4 + 0] = 0.f;
  }
}

__global__ void __launch_bounds__(256) gen_pupil_update(const GeneralParams p) {
  const int tile = p.tile0 + blockIdx.y, N = p.N;
  float2* P = p.pupil + (size_t)tile * N * N;
  const float2* Q = p.q + (size_t)tile * N * N;
  const float inv = p.apply ? rsqrt_fast(p.scal[(size_t)tile * 4 + 1]) : 0.f;
  float m = 0.f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < p.nrb * p.ncb; b += gridDim.x * blockDim.x) {
    const int bi = b / p.ncb, iw = p.ylo + bi, jw = p.xlo + (b - bi * p.ncb);
    const int t = (iw < 0 ? iw + N : iw) * N + (jw < 0 ? jw + N : jw);
    float2 v = P[t];
    if (p.apply) {
      const float2 qv = Q[t];
      v.x = fmaf(qv.x, inv, v.x);
      v.y = fmaf(qv.y, inv, v.y);
      P[t] = v;
    }
    m = fmaxf(m, fmaf(v.x, v.x, v.y * v.y));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f)
    atomicMax(reinterpret_cast<unsigned*>(p.scal + (size_t)tile * 4 + 0), __float_as_uint(m));
}

// uint16 -> 1/I (0 -> +inf: rsqrt(inf) = 0 = sqrt(0))
__global__ void __launch_bounds__(256) stack_convert_general(float* stack, const uint16_t* raw, long long first_elem, long long n,
                                                             int N, int R1) {
  const int NN = N * N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long e = first_elem + t, img = e / NN;
    const int rem = (int)(e - img * NN);
    stack[img * NN + stack_pos_offset(rem / N, rem % N, N, R1)] = 1.0f / (float)raw[e];
  }
}

// sqrt(I) of the init slot -> complex scratch [tile][N][N]   (fpmMain.cpp:319-322)
__global__ void gen_init_amp(float2* scratch, const float* stack, int n_leds, int slot, int tile0, int N, int R1) {
  const int tile = tile0 + blockIdx.y;
  const float* img = stack + ((size_t)tile * n_leds + slot) * N * N;
  float2* out = scratch + (size_t)blockIdx.y * N * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x)
    out[t] = make_float2(sqrtf(1.0f / img[stack_pos_offset(t / N, t % N, N, R1)]), 0.f);
}

}  // namespace fpm
