// fpm_kernels.cuh -- device code other than the fused update kernel (fpm_update.cuh):
// mixed-radix line FFT (spectrum seed, final objCrop), initialisation and layout helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_regs.cuh"
#include "fpm_update.cuh"
#include "fpm_update_phased.cuh"
#include "fpm_update_cluster.cuh"
#include "fpm_general.cuh"
#include "fpm_general_fused.cuh"
#include "fpm_pruned_fused.cuh"
#include "fpm_fov.cuh"
#include "fpm_fft2d.cuh"

namespace fpm {

// ------------------------------------------------------------------------------------------
// Generic mixed-radix (2,3,4,5) Stockham FFT of `lines` strided lines of length n, used for
//   * the N x N forward transform of the spectrum initialisation (fpmMain.cpp:325) and
//   * the Nlarge x Nlarge inverse transform objCrop = IDFT(objF) (fpmMain.cpp:481; Nlarge = 384,
//     600, 1536 ... are not powers of two).
// One CTA transforms LINES adjacent lines (adjacent in memory along `line_stride`), staging
// them in shared memory; `tw` = exp(-2*pi*i*k/n).
// ------------------------------------------------------------------------------------------
struct LineFFTParams {
  float2* data;             // batch of images
  const float2* tw;         // [n]
  long long batch_stride;   // elements between images
  int n;                    // transform length
  int n_lines;              // lines per image
  int line_first, n_sel;    // the lines transformed: line_first, line_first+1, ... (mod n_lines), n_sel of them
  long long elem_stride;    // elements between consecutive samples of a line
  long long line_stride;    // elements between consecutive lines
  int nrad;                 // number of stages
  int rad[12];
  float scale;
};

template <bool INV, int LINES>
__global__ void __launch_bounds__(256) line_fft_kernel(const __grid_constant__ LineFFTParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* bufA = reinterpret_cast<float2*>(smem_raw);
  const int n = p.n, np = n + 1;                // odd pitch (n is even): lanes over adjacent lines hit distinct banks
  float2* bufB = bufA + (size_t)LINES * np;
  const int line0 = blockIdx.x * LINES;
  float2* img = p.data + (size_t)blockIdx.y * p.batch_stride;
  // load: when lines are adjacent in memory (line_stride==1) let lanes run along lines
  for (int t = threadIdx.x; t < LINES * n; t += blockDim.x) {
    int l, k;
    if (p.line_stride == 1) { l = t % LINES; k = t / LINES; } else { k = t % n; l = t / n; }
    float2 v = make_float2(0.f, 0.f);
    if (line0 + l < p.n_sel) {
      int ln = p.line_first + line0 + l;
      if (ln >= p.n_lines) ln -= p.n_lines;
      v = img[(size_t)ln * p.line_stride + (size_t)k * p.elem_stride];
    }
    bufA[l * np + k] = v;
  }
  __syncthreads();
  float2* src = bufA;
  float2* dst = bufB;
  int Ns = 1;
  for (int s = 0; s < p.nrad; ++s) {
    const int R = p.rad[s];
    const int T = n / R;
    const int tstep = n / (Ns * R);
    for (int t = threadIdx.x; t < LINES * T; t += blockDim.x) {
      const int l = t / T, j = t - l * T;
      const float2* x = src + l * np;
      float2* y = dst + l * np;
      const int k = j % Ns;
      const int j0 = (j - k) * R + k;
      float2 v[5];
#pragma unroll
      for (int r = 0; r < 5; ++r)
        if (r < R) {
          float2 a = x[j + r * T];
          if (r > 0) a = twmul<INV>(a, p.tw[(r * k * tstep) % n]);
          v[r] = a;
        }
      if (R == 2) fft2<INV>(v[0], v[1]);
      else if (R == 3) fft3<INV>(v[0], v[1], v[2]);
      else if (R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
      else fft5<INV>(v[0], v[1], v[2], v[3], v[4]);
#pragma unroll
      for (int r = 0; r < 5; ++r)
        if (r < R) y[j0 + r * Ns] = v[r];
    }
    __syncthreads();
    Ns *= R;
    float2* tmp = src; src = dst; dst = tmp;
  }
  for (int t = threadIdx.x; t < LINES * n; t += blockDim.x) {
    int l, k;
    if (p.line_stride == 1) { l = t % LINES; k = t / LINES; } else { k = t % n; l = t / n; }
    if (line0 + l < p.n_sel) {
      int ln = p.line_first + line0 + l;
      if (ln >= p.n_lines) ln -= p.n_lines;
      float2 v = src[l * np + k];
      img[(size_t)ln * p.line_stride + (size_t)k * p.elem_stride] = make_float2(v.x * p.scale, v.y * p.scale);
    }
  }
}

// amplitude image of the init slot -> complex scratch [tile][N][N]   (fpmMain.cpp:319-322)
// (the stack is in the permuted device layout of stack_offset<N>)
template <int N>
__global__ void init_amp_kernel(float2* scratch, const float* stack, int n_leds, int slot, int tile0) {
  const int tile = tile0 + blockIdx.y;
  const float* img = stack + ((size_t)tile * n_leds + slot) * N * N;      // 1/I
  float2* out = scratch + (size_t)blockIdx.y * N * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x)
    out[t] = make_float2(sqrtf(1.0f / img[stack_offset<N>(t / N, t % N)]), 0.f);   // sqrt(I); 1/inf = 0
}

// objFc = 0; centre block <- fftShift(F * support); pupil = support   (fpmMain.cpp:312-313,326-343)
__global__ void init_place_kernel(float2* objFc, float2* pupil, const float2* scratch, const float* support,
                                  int N, int L, int tile0) {
  const int tile = tile0 + blockIdx.y;
  float2* O = objFc + (size_t)tile * L * L;
  float2* P = pupil + (size_t)tile * N * N;
  const float2* F = scratch + (size_t)blockIdx.y * N * N;
  const int o = L / 2 - N / 2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L * L; t += gridDim.x * blockDim.x) {
    const int r = t / L, c = t % L;
    float2 v = make_float2(0.f, 0.f);
    if (r >= o && r < o + N && c >= o && c < o + N) {
      const int i = (r - o + N / 2) % N, j = (c - o + N / 2) % N;   // undo the fftShift
      const float s = support[i * N + j];
      const float2 f = F[i * N + j];
      v = make_float2(f.x * s, f.y * s);
    }
    O[t] = v;
    if (t < N * N) P[t] = make_float2(support[t], 0.f);
  }
}

// dst[(r+L/2)%L][(c+L/2)%L] = src[r][c]  (fftShift between centred and DC-at-corner layouts)
__global__ void shift_copy_kernel(float2* dst, const float2* src, int L, long long dst_stride, long long src_stride) {
  float2* d = dst + (size_t)blockIdx.y * dst_stride;
  const float2* s = src + (size_t)blockIdx.y * src_stride;
  const int h = L / 2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L * L; t += gridDim.x * blockDim.x) {
    const int r = t / L, c = t % L;
    int r2 = r + h; if (r2 >= L) r2 -= L;
    int c2 = c + h; if (c2 >= L) c2 -= L;
    d[(size_t)r2 * L + c2] = s[t];
  }
}

}  // namespace fpm
