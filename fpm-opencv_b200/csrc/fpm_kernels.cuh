// fpm_kernels.cuh -- device code of the fused sub-aperture update (sm_100a).
//
// One CTA per tile, persistent over the whole (iteration x LED) sequence: the low-res
// exit-wave field lives in shared memory for its entire life (crop*pupil -> IFFT ->
// amplitude replacement -> FFT -> object/pupil update) and never round-trips HBM.
// Restates fpmMain.cpp:350-475 (one loop body = one "update"), SURVEY.md appendix A.
//
// Index conventions
//   window index (i,j)   DC-at-corner, what the reference calls Objfcrop (fpmMain.cpp:361)
//   wrapped  (iw,jw)     iw = i < N/2 ? i : i-N  in [-N/2, N/2)
//   absolute (r,c)       centred spectrum objFc: r = ys + N/2 + iw, c = xs + N/2 + jw
//                        (= the two fftShifts of fpmMain.cpp:358,361 folded into index math)
//   support bbox         wrapped ranges [ylo,yhi] x [xlo,xhi] that contain every non-zero
//                        pupilSupport pixel.  P == 0 outside it for ever (P starts as the
//                        support and every increment is masked, fpmMain.cpp:313,472), so
//                        O*P, dO and dP vanish there: column passes, window traffic and the
//                        epilogue are restricted to the bbox with bit-identical results.
//
// 2-D FFT = separable, N = R1*R2 per dimension, four in-register radix stages per
// transform with the field exchanged through shared memory:
//   IFFT (DIF, natural in -> digit-scrambled out):  S1 cols-A, S2 cols-B, S3 rows-A, S4 rows-B
//   FFT  (DIT, scrambled in -> natural out):        S4 rows-B', S5 rows-A', S6 cols-B', S7 cols-A'
// S4 does the last inverse stage, the amplitude replacement and the first forward stage
// on the same registers; scrambled position p = R2*k1 + k2 holds index k1 + R1*k2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_regs.cuh"

#ifdef FPM_STAGE_TIMING
#define FPM_TICK(k) do { if (tid == 0 && blockIdx.x == 0) { long long t_ = clock64(); p.stage_clk[k] += t_ - tprev_; tprev_ = t_; } } while (0)
#else
#define FPM_TICK(k) do {} while (0)
#endif

namespace fpm {

struct UpdateParams {
  float2* objFc;            // [n_tiles][L][L]   centred spectrum
  float2* pupil;            // [n_tiles][N][N]   DC-at-corner
  const uint16_t* stack;    // [n_tiles][n_leds][N][N]
  const float* support;     // [N][N]
  const short2* crop;       // [n_leds] (x = cropXStart, y = cropYStart)
  const float2* tw;         // [N] exp(-2*pi*i*n/N)
  float2* field_gmem;       // [n_tiles][N][PITCH] scratch when the field does not fit SMEM
  int L, n_leds;
  int tile0;                // first tile of this launch
  int slot_begin, n_updates;
  float delta1, delta2, eps, kappa;
  int ylo, yhi, xlo, xhi;   // support bbox (wrapped)
  int bs;                   // log2 of the block-max grid cell edge (3,4,5)
  long long* stage_clk;     // [16] per-stage cycle totals of CTA 0 (only with -DFPM_STAGE_TIMING)
};

template <int N> struct Radix {
  static constexpr int R1 = (N == 64) ? 8 : 16;
  static constexpr int R2 = N / R1;
  static constexpr int PITCH = N + 8;     // float2 per row; == 8 (mod 16) -> row-pass conflict-free
};

// in-row XOR swizzle (bijective on aligned 128-blocks): makes the stride-1 radix stage
// (each lane owns 8 consecutive complex) hit 16 distinct 8-byte bank pairs per half-warp
__device__ __forceinline__ int swz(int j) { return j ^ ((j >> 3) & 15); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int N, int NT, int MINB, bool FIELD_SMEM, bool P_SMEM>
__global__ void __launch_bounds__(NT, MINB) fpm_update_kernel(const UpdateParams p) {
  constexpr int R1 = Radix<N>::R1, R2 = Radix<N>::R2, PITCH = Radix<N>::PITCH;
  constexpr int H = N / 2;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = p.tile0 + blockIdx.x;
  const int L = p.L;
  const int NR = p.yhi - p.ylo + 1, NC = p.xhi - p.xlo + 1;   // bbox rows / cols
  const int gdim = L >> p.bs;

  // ---- shared memory carve-up -------------------------------------------------------------
  unsigned char* sp = smem_raw;
  float2* fld;
  if constexpr (FIELD_SMEM) { fld = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * N * PITCH; }
  else fld = p.field_gmem + (size_t)tile * N * PITCH;
  float2* Pc = nullptr;                       // compact pupil [NR][NC]
  if constexpr (P_SMEM) { Pc = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * NR * NC; }
  float2* twA = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * N;   // [b*R2 + a] = W^(a*b)
  float2* twB = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * N;   // [a*R1 + b] = W^(a*b)
  float* red = reinterpret_cast<float*>(sp);  sp += sizeof(float) * 64;    // [0..31] objF partials, [32..63] pupil partials
  float* G1 = reinterpret_cast<float*>(sp);                                // [gdim][gdim] block maxima of |objFc|^2

  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* Pg = p.pupil + (size_t)tile * N * N;
  const uint16_t* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * N * N;

  auto Pref = [&](int iw, int jw) -> float2& {
    if constexpr (P_SMEM) return Pc[(iw - p.ylo) * NC + (jw - p.xlo)];
    else return Pg[(iw & (N - 1)) * N + (jw & (N - 1))];
  };

  // ---- prologue: twiddle tables, pupil -> SMEM, max|P|^2, block-max grid of |objFc|^2 --------
  for (int t = tid; t < N; t += NT) {
    int b = t / R2, a = t % R2;               // twA index t = b*R2 + a
    twA[t] = p.tw[a * b];
    int a2 = t / R1, b2 = t % R1;             // twB index t = a*R1 + b
    twB[t] = p.tw[a2 * b2];
  }
  float pmax2 = 0.f;
  for (int t = tid; t < NR * NC; t += NT) {
    int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
    float2 v = Pg[(iw & (N - 1)) * N + (jw & (N - 1))];
    if constexpr (P_SMEM) Pc[t] = v;
    pmax2 = fmaxf(pmax2, fmaf(v.x, v.x, v.y * v.y));
  }
  pmax2 = warp_max(pmax2);
  if (lane == 0) red[32 + warp] = pmax2;
  {
    // one warp per (cell-row, 32-column strip); lanes along columns (coalesced)
    const int B = 1 << p.bs, strips = L >> 5;
    for (int it = warp; it < gdim * strips; it += NW) {
      int br = it / strips, c = ((it % strips) << 5) + lane;
      float m = 0.f;
      for (int rr = 0; rr < B; ++rr) {
        float2 o = objFc[(size_t)(br * B + rr) * L + c];
        m = fmaxf(m, fmaf(o.x, o.x, o.y * o.y));
      }
      for (int o = 1; o < B && o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((lane & (B - 1)) == 0) G1[br * gdim + (c >> p.bs)] = m;
    }
  }
  __syncthreads();

  const float inv_n2 = 1.0f / (float)(N * N);
  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  const float epsr = p.eps, epsi = p.kappa * p.eps;

#ifdef FPM_STAGE_TIMING
  long long tprev_ = clock64();
#endif
  for (int u = 0; u < p.n_updates; ++u) {
    const int slot = (p.slot_begin + u) % p.n_leds;
    const short2 cr = p.crop[slot];
    const int xs = cr.x, ys = cr.y;
    const uint16_t* __restrict__ img = stack + (size_t)slot * N * N;
    if (tid == 0) {   // pull this LED's intensity tile towards L2 while S1-S3 run
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(img), "r"((unsigned)(N * N * 2)) : "memory");
    }
    float2* wbase = objFc + (size_t)(ys + H) * L + (xs + H);   // absolute (iw=0,jw=0)

    // ================= S1: Phi = O*P, cols stage A (inverse) =================
    for (int g = tid; g < R2 * NC; g += NT) {
      const int i0 = g / NC, jc = g - i0 * NC;
      const int jw = p.xlo + jc, j = jw & (N - 1);
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) {
        const int i = i0 + R2 * m;
        const int iw = (i < H) ? i : i - N;
        if (iw >= p.ylo && iw <= p.yhi) v[m] = cmul(wbase[iw * L + jw], Pref(iw, jw));
        else v[m] = make_float2(0.f, 0.f);
      }
      fftR<R1, true>(v);
      const int js = swz(j);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1)
        fld[(i0 + R2 * k1) * PITCH + js] = twmul<true>(v[k1], twA[k1 * R2 + i0]);
    }
    if (NC < N) {   // columns outside the bbox are zero
      const int NZ = N - NC;
      for (int t = tid; t < N * NZ; t += NT) {
        const int i = t / NZ, j = (p.xhi + 1 + (t - i * NZ)) & (N - 1);
        fld[i * PITCH + swz(j)] = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();
    FPM_TICK(1);
    // ================= S2: cols stage B (inverse) =================
    for (int g = tid; g < R1 * NC; g += NT) {
      const int k1 = g / NC, jc = g - k1 * NC;
      const int js = swz((p.xlo + jc) & (N - 1));
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = fld[(R2 * k1 + a) * PITCH + js];
      fftR<R2, true>(v);
#pragma unroll
      for (int a = 0; a < R2; ++a) fld[(R2 * k1 + a) * PITCH + js] = v[a];
    }
    __syncthreads();
    FPM_TICK(2);
    // ================= S3: rows stage A (inverse) =================
    for (int g = tid; g < N * R2; g += NT) {
      const int row = g / R2, j0 = g % R2;
      float2* rp = fld + row * PITCH;
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) v[m] = rp[swz(j0 + R2 * m)];
      fftR<R1, true>(v);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) rp[swz(j0 + R2 * k1)] = twmul<true>(v[k1], twA[k1 * R2 + j0]);
    }
    __syncthreads();
    FPM_TICK(3);
    // ===== S4: rows stage B (inverse) + amplitude replacement + rows stage B' (forward) =====
    for (int g = tid; g < N * R1; g += NT) {
      const int row = g / R1, k1 = g % R1;
      const int y = (row / R2) + R1 * (row % R2);          // spatial row held by this physical row
      const uint16_t* ip = img + y * N + k1;                // x = k1 + R1*k2
      float amp[R2];
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) amp[k2] = (float)__ldg(ip + R1 * k2);
      float2* rp = fld + row * PITCH;
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = rp[swz(R2 * k1 + a)];
      fftR<R2, true>(v);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) {
        // psi' = sqrt(I) * psi / |psi + eps|   (fpmMain.cpp:378-393)
        const float px = v[k2].x * inv_n2, py = v[k2].y * inv_n2;
        const float tx = px + epsr, ty = py + epsi;
        const float s = sqrt_fast(amp[k2]) * rsqrt_fast(fmaf(tx, tx, ty * ty));
        v[k2] = make_float2(px * s, py * s);
      }
      fftR<R2, false>(v);
#pragma unroll
      for (int q = 0; q < R2; ++q) rp[swz(R2 * k1 + q)] = twmul<false>(v[q], twB[q * R1 + k1]);
    }
    __syncthreads();
    FPM_TICK(4);
    // ================= S5: rows stage A' (forward) =================
    for (int g = tid; g < N * R2; g += NT) {
      const int row = g / R2, q = g % R2;
      float2* rp = fld + row * PITCH;
      float2 v[R1];
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) v[k1] = rp[swz(R2 * k1 + q)];
      fftR<R1, false>(v);
#pragma unroll
      for (int r = 0; r < R1; ++r) rp[swz(R2 * r + q)] = v[r];
    }
    __syncthreads();
    FPM_TICK(5);
    // ================= S6: cols stage B' (forward) =================
    for (int g = tid; g < R1 * NC; g += NT) {
      const int k1 = g / NC, jc = g - k1 * NC;
      const int js = swz((p.xlo + jc) & (N - 1));
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = fld[(R2 * k1 + a) * PITCH + js];
      fftR<R2, false>(v);
#pragma unroll
      for (int q = 0; q < R2; ++q) fld[(R2 * k1 + q) * PITCH + js] = twmul<false>(v[q], twB[q * R1 + k1]);
    }
    __syncthreads();
    FPM_TICK(6);
    // ================= S7: cols stage A' (forward) -> Phi' in natural order =================
    for (int g = tid; g < R2 * NC; g += NT) {
      const int q = g / NC, jc = g - q * NC;
      const int js = swz((p.xlo + jc) & (N - 1));
      float2 v[R1];
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) v[k1] = fld[(R2 * k1 + q) * PITCH + js];
      fftR<R1, false>(v);
#pragma unroll
      for (int r = 0; r < R1; ++r) fld[(R2 * r + q) * PITCH + js] = v[r];
    }
    __syncthreads();
    FPM_TICK(7);
    // ================= S8: object update + block maxima (fpmMain.cpp:406-447) =================
    float pm2 = red[32];
#pragma unroll
    for (int w = 1; w < NW; ++w) pm2 = fmaxf(pm2, red[32 + w]);
    const float pupil_abs_max = sqrtf(pm2);
    {
      const int B = 1 << p.bs;
      const int r0 = ys + H + p.ylo, r1 = ys + H + p.yhi;      // absolute update rectangle (inclusive)
      const int c0 = xs + H + p.xlo, c1 = xs + H + p.xhi;
      const int br0 = r0 >> p.bs, nbr = (r1 >> p.bs) - br0 + 1;
      const int ct0 = c0 >> 5, nct = (c1 >> 5) - ct0 + 1;
      for (int it = warp; it < nbr * nct; it += NW) {
        const int br = br0 + it / nct, c = ((ct0 + it % nct) << 5) + lane;
        const int cb0 = (c >> p.bs) << p.bs;                    // first column of this lane's cell
        const bool cell_on = (cb0 + B - 1 >= c0) && (cb0 <= c1);
        const bool col_in = (c >= c0) && (c <= c1);
        const int jw = c - xs - H, j = jw & (N - 1);
        float m = 0.f;
        if (cell_on) {
          for (int rr = br << p.bs; rr < ((br + 1) << p.bs); ++rr) {
            float2* gp = objFc + (size_t)rr * L + c;
            float2 O = *gp;
            float a2 = fmaf(O.x, O.x, O.y * O.y);
            if (col_in && rr >= r0 && rr <= r1) {
              const int iw = rr - ys - H, i = iw & (N - 1);
              float2* fp = fld + i * PITCH + swz(j);
              const float2 Pv = Pref(iw, jw);
              const float2 Phi = cmul(O, Pv);
              const float2 Phi2 = *fp;
              const float2 d = csub(Phi2, Phi);
              // dO = d * |P| conj(P) / (max|P| * ((|P|^2 + delta2) + i*kappa*delta2))
              const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y), pa = sqrtf(pa2);
              const float2 num = cmulc(d, Pv);                     // d * conj(P)
              const float A = pa2 + p.delta2;
              const float sc = pa / (pupil_abs_max * fmaf(A, A, kd2 * kd2));
              const float2 dO = make_float2((num.x * A + num.y * kd2) * sc, (num.y * A - num.x * kd2) * sc);
              // Q = d * |O| conj(O) / ((|O|^2 + delta1) + i*kappa*delta1) * support   (fpmMain.cpp:459-472)
              const float oa = sqrtf(a2);
              const float2 numq = cmulc(d, O);
              const float A1 = a2 + p.delta1;
              const float sq = oa * __ldg(p.support + i * N + j) / fmaf(A1, A1, kd1 * kd1);
              *fp = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
              O = cadd(O, dO);
              *gp = O;
              a2 = fmaf(O.x, O.x, O.y * O.y);
            }
            m = fmaxf(m, a2);
          }
        }
        for (int o = 1; o < B && o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (cell_on && (lane & (B - 1)) == 0) G1[br * gdim + (c >> p.bs)] = m;
      }
    }
    __syncthreads();
    FPM_TICK(8);
    // ---- max |objF| over the whole updated spectrum (fpmMain.cpp:460,467) ----
    {
      float m = 0.f;
      for (int t = tid; t < gdim * gdim; t += NT) m = fmaxf(m, G1[t]);
      m = warp_max(m);
      if (lane == 0) red[warp] = m;
    }
    __syncthreads();
    FPM_TICK(9);
    float om2 = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) om2 = fmaxf(om2, red[w]);
    const float inv_objf_max = 1.0f / sqrtf(om2);
    // ================= S9: pupil update  P += Q / max|objF|  (fpmMain.cpp:470-475) =================
    float pnew = 0.f;
    for (int t = tid; t < NR * NC; t += NT) {
      const int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
      const float2 Q = fld[(iw & (N - 1)) * PITCH + swz(jw & (N - 1))];
      float2& pr = Pref(iw, jw);
      float2 v = pr;
      v.x = fmaf(Q.x, inv_objf_max, v.x);
      v.y = fmaf(Q.y, inv_objf_max, v.y);
      pr = v;
      pnew = fmaxf(pnew, fmaf(v.x, v.x, v.y * v.y));
    }
    pnew = warp_max(pnew);
    if (lane == 0) red[32 + warp] = pnew;   // last read of red[32..] was before two barriers (S8)
    __syncthreads();
    FPM_TICK(10);
  }

  if constexpr (P_SMEM) {
    for (int t = tid; t < NR * NC; t += NT) {
      int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
      Pg[(iw & (N - 1)) * N + (jw & (N - 1))] = Pc[t];
    }
  }
}

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
  long long elem_stride;    // elements between consecutive samples of a line
  long long line_stride;    // elements between consecutive lines
  int nrad;                 // number of stages
  int rad[12];
  float scale;
};

template <bool INV, int LINES>
__global__ void __launch_bounds__(256) line_fft_kernel(const LineFFTParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* bufA = reinterpret_cast<float2*>(smem_raw);
  float2* bufB = bufA + (size_t)LINES * p.n;
  const int n = p.n;
  const int line0 = blockIdx.x * LINES;
  float2* img = p.data + (size_t)blockIdx.y * p.batch_stride;
  // load: when lines are adjacent in memory (line_stride==1) let lanes run along lines
  for (int t = threadIdx.x; t < LINES * n; t += blockDim.x) {
    int l, k;
    if (p.line_stride == 1) { l = t % LINES; k = t / LINES; } else { k = t % n; l = t / n; }
    float2 v = make_float2(0.f, 0.f);
    if (line0 + l < p.n_lines) v = img[(size_t)(line0 + l) * p.line_stride + (size_t)k * p.elem_stride];
    bufA[l * n + k] = v;
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
      const float2* x = src + l * n;
      float2* y = dst + l * n;
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
    if (line0 + l < p.n_lines) {
      float2 v = src[l * n + k];
      img[(size_t)(line0 + l) * p.line_stride + (size_t)k * p.elem_stride] = make_float2(v.x * p.scale, v.y * p.scale);
    }
  }
}

// amplitude image of the init slot -> complex scratch [tile][N][N]   (fpmMain.cpp:319-322)
__global__ void init_amp_kernel(float2* scratch, const uint16_t* stack, int N, int n_leds, int slot, int tile0) {
  const int tile = tile0 + blockIdx.y;
  const uint16_t* img = stack + ((size_t)tile * n_leds + slot) * N * N;
  float2* out = scratch + (size_t)blockIdx.y * N * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x)
    out[t] = make_float2(sqrtf((float)img[t]), 0.f);
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
