// fpm_fft2d.cuh -- planned line FFT for the spectrum-sized transforms (sm_100a).
//
// objCrop = IDFT(objF) (fpmMain.cpp:481) is an Nlarge x Nlarge transform per tile and iteration batch: 384, 512, 600,
// 1024, 1536 ... (2,3,5-smooth, not powers of two).  line_fft_kernel (fpm_kernels.cuh) takes run-time radices 2..5 and
// spends five or six shared-memory passes per line; here the length is a compile-time product R0 * R1 * R2 of radices up
// to 16 (composite ones as register Cooley-Tukey splits, fft_regs.cuh), three in-place decimation-in-frequency stages per
// line, twiddles from a shared-memory table, and the fftShift between the centred device spectrum and the reference's
// DC-at-corner layout (fpmMain.cpp:358 convention) folded into the load addresses of the row pass -- the separate
// shift_copy_kernel pass over the spectrum disappears.
//
// One CTA transforms LINES = 16 lines.  Butterfly work items are (line, index) with the line fastest: the 16 lanes of a
// half-warp touch 16 different lines at the same in-line position, the line pitch is odd, so every 64-bit shared-memory
// access is conflict-free for any radix.  In place: position p = k0*(M0+1) + k1*M1 + k2 (M0 = L/R0, M1 = M0/R1; one pad
// element per R0-block so that the row pass's coalesced store, lanes along k, is also conflict-free) ends up holding
// X[k0 + R0*k1 + R0*R1*k2]; the permutation is undone by the store addresses.
#pragma once
#include <cuda_runtime.h>
#include "fft_regs.cuh"

namespace fpm {

struct PlanFFTParams {
  const float2* src;        // batch of L x L images (may equal dst)
  float2* dst;
  const float2* tw;         // [L] exp(-2*pi*i*k/L)
  long long src_stride, dst_stride;   // elements between images
  int shift;                // rows pass only: read src through the fftShift (src is centred, dst DC-at-corner)
  int cols;                 // 0: lines are rows (contiguous), 1: lines are columns
  float scale;
};

template <int R0, int R1, int R2> struct PlanShape {
  static constexpr int L = R0 * R1 * R2, M0 = L / R0, M1 = M0 / R1;
  static constexpr int LP = L + R0;                 // padded line: block k0 starts at k0 * (M0 + 1)
  static constexpr int PITCH = LP | 1;              // odd line pitch
  static constexpr int LINES = 16;
  static constexpr size_t smem = sizeof(float2) * ((size_t)LINES * PITCH + L);
};

template <int R0, int R1, int R2, bool INV>
__global__ void __launch_bounds__(256) plan_fft_kernel(const __grid_constant__ PlanFFTParams p) {
  using P = PlanShape<R0, R1, R2>;
  constexpr int L = P::L, M0 = P::M0, M1 = P::M1, PITCH = P::PITCH, LINES = P::LINES, NT = 256;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  float2* tw = buf + (size_t)LINES * PITCH;
  const int tid = threadIdx.x;
  const int line0 = blockIdx.x * LINES;
  const float2* src = p.src + (size_t)blockIdx.y * p.src_stride;
  float2* dst = p.dst + (size_t)blockIdx.y * p.dst_stride;
  constexpr int h = L / 2;

  for (int t = tid; t < L; t += NT) tw[t] = p.tw[t];
  // element n of line l sits at l*PITCH + n + n/M0.  Rows: lanes along the line (coalesced); columns: lanes along 16
  // adjacent columns (16 x 8 bytes contiguous per row).  LB global loads of a thread are in flight together.
  constexpr int LB = 8, NLOAD = (LINES * L + NT - 1) / NT;
  for (int b = 0; b < NLOAD; b += LB) {
    float2 v[LB];
    int so[LB];
#pragma unroll
    for (int k = 0; k < LB; ++k) {
      const int t = tid + (b + k) * NT;
      int l, n;
      if (!p.cols) { l = t / L; n = t - l * L; } else { l = t % LINES; n = t / LINES; }
      const bool in = (b + k < NLOAD) && t < LINES * L && line0 + l < L;
      so[k] = (b + k < NLOAD && t < LINES * L) ? l * PITCH + n + n / M0 : -1;
      size_t g;
      if (!p.cols) {
        int r = line0 + l, c = n;
        if (p.shift) { r += h; if (r >= L) r -= L; c += h; if (c >= L) c -= L; }
        g = (size_t)r * L + c;
      } else g = (size_t)n * L + line0 + l;
      v[k] = in ? __ldcs(src + g) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < LB; ++k)
      if (so[k] >= 0) buf[so[k]] = v[k];
  }
  __syncthreads();

  // stage 0: R0-point butterflies over x[m + q*M0], twiddle W_L^(m*k0)
  for (int t = tid; t < LINES * M0; t += NT) {
    const int l = t % LINES, m = t / LINES;
    float2* x = buf + l * PITCH + m;
    float2 v[R0];
#pragma unroll
    for (int q = 0; q < R0; ++q) v[q] = x[q * (M0 + 1)];
    fft_reg<R0, INV>(v);
    static_for<0, R0>([&](auto I) {
      constexpr int i = decltype(I)::value, k0 = radix_out<R0>(i);
      x[k0 * (M0 + 1)] = (k0 == 0) ? v[i] : twmul<INV>(v[i], tw[m * k0]);
    });
  }
  __syncthreads();
  // stage 1: inside every R0-block, R1-point butterflies over x[m' + q*M1], twiddle W_M0^(m'*k1) = W_L^(R0*m'*k1)
  for (int t = tid; t < LINES * R0 * M1; t += NT) {
    const int l = t % LINES, rest = t / LINES;
    const int b = rest / M1, m = rest - b * M1;
    float2* x = buf + l * PITCH + b * (M0 + 1) + m;
    float2 v[R1];
#pragma unroll
    for (int q = 0; q < R1; ++q) v[q] = x[q * M1];
    fft_reg<R1, INV>(v);
    static_for<0, R1>([&](auto I) {
      constexpr int i = decltype(I)::value, k1 = radix_out<R1>(i);
      x[k1 * M1] = (k1 == 0 || R2 == 1) ? v[i] : twmul<INV>(v[i], tw[R0 * m * k1]);
    });
  }
  __syncthreads();
  // stage 2: R2-point butterflies over the M1 = R2 consecutive elements of every (k0, k1) block
  if constexpr (R2 > 1) {
    for (int t = tid; t < LINES * R0 * R1; t += NT) {
      const int l = t % LINES, rest = t / LINES;
      const int b = rest / R1, k1 = rest - b * R1;
      float2* x = buf + l * PITCH + b * (M0 + 1) + k1 * M1;
      float2 v[R2];
#pragma unroll
      for (int q = 0; q < R2; ++q) v[q] = x[q];
      fft_reg<R2, INV>(v);
      static_for<0, R2>([&](auto I) {
        constexpr int i = decltype(I)::value;
        x[radix_out<R2>(i)] = v[i];
      });
    }
    __syncthreads();
  }

  // X[k] sits at (k % R0)*(M0+1) + ((k / R0) % R1)*M1 + k / (R0*R1)
  auto pos = [](int k) { return (k % R0) * (M0 + 1) + ((k / R0) % R1) * M1 + k / (R0 * R1); };
  if (!p.cols) {
    for (int t = tid; t < LINES * L; t += NT) {
      const int l = t / L, k = t - l * L;
      if (line0 + l < L) {
        const float2 v = buf[l * PITCH + pos(k)];
        dst[(size_t)(line0 + l) * L + k] = make_float2(v.x * p.scale, v.y * p.scale);
      }
    }
  } else {
    for (int t = tid; t < LINES * L; t += NT) {
      const int l = t % LINES, k = t / LINES;
      if (line0 + l < L) {
        const float2 v = buf[l * PITCH + pos(k)];
        dst[(size_t)k * L + line0 + l] = make_float2(v.x * p.scale, v.y * p.scale);
      }
    }
  }
}

}  // namespace fpm
