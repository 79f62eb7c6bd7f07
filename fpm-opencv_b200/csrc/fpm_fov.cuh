// fpm_fov.cuh -- the callers either side of the reconstruction loop for a full field of view:
//   * frame ingest: every tile's Np x Np ROI of one LED frame is cut, dark-field divided, background-subtracted
//     on the device (fpmMain.cpp:109-144 does this once per tile AND frame on the host: O(tiles) redundant I/O);
//   * mosaic: the tiles' objCrop (fpmMain.cpp:481) are blended into one high-resolution amplitude image
//     (the reference reconstructs one tile per process and has no tile loop: fpmMain.cpp:519,532-533).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fpm_update.cuh"

namespace fpm {

// Background estimate of one frame (fpmMain.cpp:131-140): cv::mean of two Np x Np ROIs = sum * (1/Np^2) in double,
// averaged, clamped at bgThreshold, rounded half away from zero, stored as int16.  One CTA.
__global__ void __launch_bounds__(1024) ingest_bg_kernel(const uint16_t* frame, int w, int Np, int bk1x, int bk1y, int bk2x,
                                                         int bk2y, int bg_threshold, int* bg_out) {
  unsigned long long s1 = 0, s2 = 0;
  for (int t = threadIdx.x; t < Np * Np; t += blockDim.x) {
    const int y = t / Np, x = t - y * Np;
    s1 += frame[(size_t)(bk1y + y) * w + bk1x + x];
    s2 += frame[(size_t)(bk2y + y) * w + bk2x + x];
  }
  __shared__ unsigned long long r1[32], r2[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += r1[k]; b += r2[k]; }
    const double inv = 1.0 / ((double)Np * (double)Np);
    const double bk1 = (double)a * inv, bk2 = (double)b * inv;
    double bg = (bk2 + bk1) / 2;
    if (bg > (double)bg_threshold) bg = (double)bg_threshold;
    *bg_out = (int)(short)(int)round(bg);
  }
}

// ROI cut + cv::divide (round half to even, x/0 -> 0) + saturating background subtraction for every tile
// (blockIdx.y) of one LED frame; writes the uint16 result and 1/I in the update kernels' stack layout.
// perm = 1: stack_offset<N> order of the fused kernels (R1 x R2 factorisation), 2: position-major order of
// fpm_pruned_fused.cuh (stack_pos_offset), 0: natural order.
__global__ void __launch_bounds__(256) ingest_tiles_kernel(const uint16_t* frame, int w, const int2* origin, int tile0,
                                                           uint16_t* raw, float* stack, int n_leds, int slot, int Np, int R1,
                                                           int perm, int divisor, const int* bg, int bg_imm) {
  const int tile = tile0 + blockIdx.y;
  const int2 o = origin[tile];
  const int bgv = bg ? *bg : bg_imm;                      // measured on the device (ingest_bg_kernel) or given by the host
  const size_t base = ((size_t)tile * n_leds + slot) * Np * Np;
  const int R2 = Np / R1;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Np * Np; t += gridDim.x * blockDim.x) {
    const int y = t / Np, x = t - y * Np;
    int p = frame[(size_t)(o.y + y) * w + o.x + x];
    if (divisor != 1) {
      const double q = divisor == 0 ? 0.0 : rint((double)p / (double)divisor);
      p = q < 0 ? 0 : q > 65535 ? 65535 : (int)q;
    }
    int v = p - bgv;
    v = v < 0 ? 0 : v > 65535 ? 65535 : v;
    raw[base + t] = (uint16_t)v;
    int off = t;
    if (perm == 1) {
      const int pos = R2 * (y % R1) + y / R1;
      off = ((x % R1) * Np + pos) * R2 + x / R1;
    } else if (perm == 2) {
      off = (R2 * (y % R1) + y / R1) * Np + R2 * (x % R1) + x / R1;
    }
    stack[base + off] = 1.0f / (float)v;
  }
}

__global__ void set_bg_kernel(int* bg_out, int v) { *bg_out = v; }          // keeps fpmb200_ingest_bg's table complete

// Feathered mosaic of |objCrop| over a regular tile grid.  Tile (ix, iy) covers low-res pixels
// [x0 + ix*step, +Np) x [y0 + iy*step, +Np), i.e. f = L/Np times that in the mosaic; where tiles overlap the
// weights ramp linearly over the overlap (separable), and the sum is normalised -- a gather, so the result does
// not depend on any accumulation order.
struct MosaicParams {
  const float2* tiles;      // [nx*ny][L][L] objCrop
  float* out;               // [Hm][Wm]
  int L, Np, step, nx, ny, Wm, Hm;
};
__device__ __forceinline__ float ramp_weight(int u, int L, int ov) {   // u = coordinate inside the tile, ov = overlap (hi-res)
  if (ov <= 0) return 1.f;
  const int d = min(u, L - 1 - u);
  return d >= ov ? 1.f : (float)(d + 1) / (float)(ov + 1);
}
__global__ void __launch_bounds__(256) mosaic_kernel(const MosaicParams p) {
  const int f = p.L / p.Np, sh = p.step * f, ov = p.L - sh;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)p.Wm * p.Hm; t += (long long)gridDim.x * blockDim.x) {
    const int Y = (int)(t / p.Wm), X = (int)(t - (long long)Y * p.Wm);
    // tiles whose span [i*sh, i*sh + L) contains the coordinate
    const int ix1 = min(X / sh, p.nx - 1), iy1 = min(Y / sh, p.ny - 1);
    float acc = 0.f, wsum = 0.f;
    for (int iy = iy1; iy >= 0 && Y - iy * sh < p.L; --iy)
      for (int ix = ix1; ix >= 0 && X - ix * sh < p.L; --ix) {
        const int u = X - ix * sh, v = Y - iy * sh;
        const float wgt = ramp_weight(u, p.L, ov) * ramp_weight(v, p.L, ov);
        const float2 o = p.tiles[((size_t)(iy * p.nx + ix) * p.L + v) * p.L + u];
        acc = fmaf(wgt, sqrtf(fmaf(o.x, o.x, o.y * o.y)), acc);
        wsum += wgt;
      }
    p.out[t] = wsum > 0.f ? acc / wsum : 0.f;
  }
}

}  // namespace fpm
