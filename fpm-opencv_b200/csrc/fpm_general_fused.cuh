// fpm_general_fused.cuh -- fused sub-aperture update for tile sizes that are not powers of two but whose field fits
// one SM twice (even Np <= 112 with prime factors 2, 3, 5: the shipped dataset*.json use cropSizeX = 90 and 100).
//
// One CTA per tile, persistent over all updates of the launch (the reference's sequential LED order,
// fpmMain.cpp:350-475, is kept inside the CTA; parallelism is across tiles and inside the Np x Np transform):
//
//   A   pending pupil update P += Q / max|objF| (of the previous LED, fpmMain.cpp:470-475) fused with the window fetch
//       and Phi = O * P (:358-364); max|P|^2 reduced on the way (:415)
//   I   inverse 2-D transform, mixed-radix (4,2,3,5) Stockham stages between two shared-memory copies of the field,
//       rows then columns, unscaled (the 1/Np^2 of ifft2 cancels in psi/|psi+eps|; eps is scaled instead)   (:365)
//   M   psi' = psi * rsqrt(|psi+eps|^2 * (1/I))                                                             (:378-393)
//   F   forward 2-D transform                                                                                (:394)
//   C   dPhi = Phi' - O P;  O += dPhi |P| P* / D_O written to the spectrum;  Q = dPhi |O| O* / D_P * S kept in the
//       second field buffer (it is the transform's scratch and free between F and the next I)      (:406-447,459-472)
//   D   exact max|objF|: the 16x16-pixel cells of the grid of cell maxima that the window touches are rebuilt from
//       the spectrum, the grid (shared memory) is scanned                                                  (:460,467)
//
// The field never leaves shared memory; lanes run over LINES in every transform stage (rows: odd pitch, columns:
// adjacent addresses), so all 64-bit accesses of a half-warp fall into distinct bank pairs for any Np, and the
// butterfly index -- hence the twiddle -- is uniform over (nearly) the whole warp.
// Same arithmetic and conventions as fpm_general.cuh (the unfused path, which stays for tiles too large for this
// kernel, e.g. cropSizeX = 200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "fft_regs.cuh"
#include "fpm_update.cuh"
#include "fpm_general.cuh"

namespace fpm {

struct GeneralFusedParams {
  float2* objFc;            // [n_tiles][L][L] centred
  float2* pupil;            // [n_tiles][N][N] DC-at-corner
  const float* stack;       // [n_tiles][n_leds][N][N] 1/I, natural order
  const float* support;     // [N][N]
  const short2* crop;       // [n_leds]
  const float2* tw;         // [N] exp(-2*pi*i*k/N)
  int N, L, n_leds, tile0;
  int slot_begin, n_updates;
  int cgr, cgc;             // grid of 16x16-pixel max-cells over the spectrum (edge cells partial)
  int nrad, rad[8];         // transform stages
  float delta1, delta2, eps, kappa;
};

__host__ __device__ inline size_t general_fused_smem_bytes(int N, int cgr, int cgc) {
  const size_t fld = (sizeof(float2) * (size_t)N * (N + 1) + 15) / 16 * 16;
  return 2 * fld + sizeof(float2) * N + sizeof(float) * ((size_t)cgr * cgc + 64) + 32;
}

// What the fused column stages need besides the field (MODE 1: amplitude replacement, MODE 2: object / pupil increments)
struct StageExtra {
  const float* inv_i;                 // MODE 1: 1/I of this LED, [N][N]
  float2* O;                          // MODE 2: window origin in the centred spectrum (wrapped indices -H .. H-1)
  const float2* P;                    // MODE 2: pupil [N][N]
  const float* support;               // MODE 2: [N][N]
  int L;
  float epsr, epsi, delta1, delta2, kd1, kd2, inv_pmax;
};

// One Stockham stage of a two-stage plan N = R1 * R2 over all N lines, src -> dst, radix R at compile time: each work
// item is one R-point transform in registers (fft_reg, composite radices 6 / 9 / 10 included).  FIRST: radix R1, no
// twiddles, outputs contiguous (j*R + k); otherwise radix R2, inputs twiddled by W_N^(r*j), outputs at j + k*R1.
// Work item t = j * N + l: lanes over lines.  MODE 1 / 2 (column stages only: es = pitch, ls = 1, so lanes run over
// columns and every global access below is coalesced) fuse the pointwise step that follows the transform into the
// stores: M (amplitude replacement, fpmMain.cpp:378-393) resp. C (object update written to the spectrum, pupil increment
// Q left in dst in place of Phi', :406-447, 459-472); their global operands are requested before the butterflies.
template <int NT, int R, bool INV, bool FIRST, int MODE>
__device__ __forceinline__ void plan_stage(const float2* __restrict__ src, float2* __restrict__ dst,
                                           const float2* __restrict__ tws, int N, int es, int ls, int tid,
                                           const StageExtra& x) {
  const int T = N / R, total = N * T;
  const int qNT = NT / N, rNT = NT % N;
  int j = tid / N, l = tid % N;
  for (int t = tid; t < total; t += NT) {
    float2 v[R];
    float ii[MODE == 1 ? R : 1];
    float2 Ov[MODE == 2 ? R : 1], Pv[MODE == 2 ? R : 1];
    int oo[MODE == 2 ? R : 1];
    if constexpr (MODE == 1) {
      static_for<0, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        ii[i] = __ldg(x.inv_i + (j + radix_out<R>(i) * T) * N + l);
      });
    }
    if constexpr (MODE == 2) {
      const int wl = wrap_half(l, N);
      static_for<0, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const int row = j + radix_out<R>(i) * T;
        oo[i] = wrap_half(row, N) * x.L + wl;
        Ov[i] = x.O[oo[i]];
        Pv[i] = x.P[row * N + l];
      });
    }
    const float2* s = src + l * ls + j * es;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = s[r * T * es];
    if constexpr (!FIRST) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = twmul<INV>(v[r], tws[r * j]);
    }
    fft_reg<R, INV>(v);
    float2* d = dst + l * ls + (FIRST ? j * R : j) * es;
    static_for<0, R>([&](auto I) {
      constexpr int i = decltype(I)::value, ko = radix_out<R>(i);
      float2 val = v[i];
      if constexpr (MODE == 1) {
        const float2 tt = make_float2(val.x + x.epsr, val.y + x.epsi);
        const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * ii[i]);          // sqrt(I)/|psi+eps|; I = 0 -> 0
        val = make_float2(val.x * sc, val.y * sc);
      }
      if constexpr (MODE == 2) {
        const float2 O0 = Ov[i], P0 = Pv[i];
        const float sup = __ldg(x.support + (j + ko * T) * N + l);
        const float2 dd = csub(val, cmul(O0, P0));
        const float pa2 = fmaf(P0.x, P0.x, P0.y * P0.y);
        const float2 num = cmulc(dd, P0);
        const float A = pa2 + x.delta2;
        const float sc = __fdividef(sqrt_fast(pa2) * x.inv_pmax, fmaf(A, A, x.kd2 * x.kd2));
        x.O[oo[i]] = make_float2(O0.x + (num.x * A + num.y * x.kd2) * sc, O0.y + (num.y * A - num.x * x.kd2) * sc);
        const float oa2 = fmaf(O0.x, O0.x, O0.y * O0.y);
        const float2 numq = cmulc(dd, O0);
        const float A1 = oa2 + x.delta1;
        const float sq = __fdividef(sqrt_fast(oa2) * sup, fmaf(A1, A1, x.kd1 * x.kd1));
        val = make_float2((numq.x * A1 + numq.y * x.kd1) * sq, (numq.y * A1 - numq.x * x.kd1) * sq);
      }
      d[(FIRST ? ko : ko * T) * es] = val;
    });
    j += qNT; l += rNT;
    if (l >= N) { l -= N; ++j; }
  }
  __syncthreads();
}

// R1 * R2 == Np: two-stage plan with compile-time radices (M and C fused into the column stages, the pupil increment Q
// shares the field buffer); R1 == 0: radices from p.rad at run time, every step its own pass.
template <int NT, int R1, int R2>
__global__ void __launch_bounds__(NT, 1) fpm_update_general_kernel(const __grid_constant__ GeneralFusedParams p) {
  constexpr bool PLAN = R1 > 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N, L = p.L, H = N / 2, PITCH = N + 1, NN = N * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  constexpr int UN = 4;                                                  // elements per thread in flight in the pointwise passes
  const size_t fld_bytes = (sizeof(float2) * (size_t)N * PITCH + 15) / 16 * 16;
  float2* bufF = reinterpret_cast<float2*>(smem_raw);                    // the field
  float2* bufQ = reinterpret_cast<float2*>(smem_raw + fld_bytes);        // transform scratch / pupil increment Q
  float2* tws = reinterpret_cast<float2*>(smem_raw + 2 * fld_bytes);
  float* red = reinterpret_cast<float*>(tws + N);                        // [0..31] |P|^2 partials, [32..63] |objF|^2
  float* U = red + 64;                                                   // [cgr][cgc] cell maxima of |objFc|^2
  float2* qbuf = PLAN ? bufF : bufQ;                                     // where the pupil increment Q waits for max|objF|

  const int tile = p.tile0 + blockIdx.x;
  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* P = p.pupil + (size_t)tile * NN;
  const float* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * NN;
  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  const float epsr = p.eps * (float)NN, epsi = p.kappa * epsr;

  // element index stepping t -> (i, j) = (t / N, t % N) without dividing inside the loops
  const int qNT = NT / N, rNT = NT % N, ti0 = tid / N, tj0 = tid % N;
  auto step = [&](int& q, int& r) { q += qNT; r += rNT; if (r >= N) { r -= N; ++q; } };

  // per-lane partial maximum of the 16x16 cell (a, b) from the spectrum, one warp: lanes = 2 rows x 16 columns per load
  auto cell_part = [&](int a, int b) -> float {
    const int c = (b << 4) + (lane & 15);
    float m = 0.f;
    if (c < L) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int r = (a << 4) + 2 * rr + (lane >> 4);
        if (r < L) {
          const float2 o = __ldcg(objFc + (size_t)r * L + c);
          m = fmaxf(m, fmaf(o.x, o.x, o.y * o.y));
        }
      }
    }
    return m;
  };
  auto grid_max = [&]() -> float {                   // scan of U; all threads return the maximum (two barriers)
    float m = 0.f;
    for (int t = tid; t < p.cgr * p.cgc; t += NT) m = fmaxf(m, U[t]);
    m = warp_max(m);
    if (lane == 0) red[32 + warp] = m;
    __syncthreads();
    m = red[32 + (lane % NW)];
    m = warp_max(m);
    __syncthreads();
    return m;
  };
  // One 1-D transform of all N lines, src -> dst per stage; the result ends in `a` when nrad is even, in `b` otherwise.
  // es / ls: element / line stride in float2.
  auto lines_fft = [&](auto inv_tag, float2*& a, float2*& b, int es, int ls) {
    constexpr bool INV = decltype(inv_tag)::value;
    int Ns = 1;
    for (int s = 0; s < p.nrad; ++s) {
      const int R = p.rad[s], T = N / R, tstep = N / (Ns * R);
      int j = ti0, l = tj0;                                   // work item t = j * N + l: lanes over lines
      for (int t = tid; t < N * T; t += NT, step(j, l)) {
        const int k = j % Ns, j0 = (j - k) * R + k;
        const float2* x = a + l * ls + j * es;
        float2* y = b + l * ls + j0 * es;
        float2 v[5];
#pragma unroll
        for (int r = 0; r < 5; ++r)
          if (r < R) {
            float2 w = x[r * T * es];
            if (r > 0 && Ns > 1) w = twmul<INV>(w, tws[r * k * tstep]);      // r*k*tstep < N always
            v[r] = w;
          }
        if (R == 2) fft2<INV>(v[0], v[1]);
        else if (R == 3) fft3<INV>(v[0], v[1], v[2]);
        else if (R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
        else fft5<INV>(v[0], v[1], v[2], v[3], v[4]);
#pragma unroll
        for (int r = 0; r < 5; ++r)
          if (r < R) y[r * Ns * es] = v[r];
      }
      __syncthreads();
      Ns *= R;
      float2* tmp = a; a = b; b = tmp;
    }
  };

  // ---- prologue: twiddles, the grid of cell maxima, max|objF|^2 ----
  for (int t = tid; t < N; t += NT) tws[t] = p.tw[t];
  for (int t = warp; t < p.cgr * p.cgc; t += 2 * NW) {           // two cells per warp in flight
    const int t1 = t + NW;
    const float m0 = cell_part(t / p.cgc, t % p.cgc);
    const float m1 = t1 < p.cgr * p.cgc ? cell_part(t1 / p.cgc, t1 % p.cgc) : 0.f;
    const float r0 = warp_max(m0), r1 = warp_max(m1);
    if (lane == 0) { U[t] = r0; if (t1 < p.cgr * p.cgc) U[t1] = r1; }
  }
  __syncthreads();
  float omax2 = grid_max();
  bool pending = false;                                       // Q of the previous update not yet added to P

  int slot = p.slot_begin % p.n_leds;
  for (int u = 0; u < p.n_updates; ++u) {
    const short2 cr = p.crop[slot];
    float2* O = objFc + (size_t)(cr.y + H) * L + (cr.x + H);           // window origin: wrapped indices -H .. H-1
    float2 *fa = bufF, *fb = bufQ;

    // ---- A: (P += Q / max|objF|), Phi = O * P, max|P|^2 ----
    {
      const float inv_omax = pending ? rsqrt_fast(omax2) : 0.f;
      float pm2 = 0.f;
      int i = ti0, j = tj0;
      constexpr int UA = PLAN ? 8 : UN;
      for (int t = tid; t < NN; t += UA * NT) {               // UA elements in flight: their global loads are issued together
        float2 pv[UA], ov[UA];
        int fo[UA];
#pragma unroll
        for (int k = 0; k < UA; ++k) {
          fo[k] = i * PITCH + j;
          if (t + k * NT < NN) { pv[k] = P[t + k * NT]; ov[k] = O[wrap_half(i, N) * L + wrap_half(j, N)]; }
          step(i, j);
        }
#pragma unroll
        for (int k = 0; k < UA; ++k)
          if (t + k * NT < NN) {
            if (pending) {
              const float2 qv = qbuf[fo[k]];
              pv[k].x = fmaf(qv.x, inv_omax, pv[k].x);
              pv[k].y = fmaf(qv.y, inv_omax, pv[k].y);
              P[t + k * NT] = pv[k];
            }
            pm2 = fmaxf(pm2, fmaf(pv[k].x, pv[k].x, pv[k].y * pv[k].y));
            bufF[fo[k]] = cmul(ov[k], pv[k]);
          }
      }
      pm2 = warp_max(pm2);
      if (lane == 0) red[warp] = pm2;
    }
    __syncthreads();
    const float inv_pmax = rsqrt_fast(warp_max(red[lane % NW]));

    if constexpr (PLAN) {
      StageExtra x;
      x.inv_i = stack + (size_t)slot * NN; x.O = O; x.P = P; x.support = p.support; x.L = L;
      x.epsr = epsr; x.epsi = epsi; x.delta1 = p.delta1; x.delta2 = p.delta2; x.kd1 = kd1; x.kd2 = kd2; x.inv_pmax = inv_pmax;
      // ---- I: inverse transform, rows then columns; M fused into the last column stage ----
      plan_stage<NT, R1, true, true, 0>(bufF, bufQ, tws, N, 1, PITCH, tid, x);
      plan_stage<NT, R2, true, false, 0>(bufQ, bufF, tws, N, 1, PITCH, tid, x);
      plan_stage<NT, R1, true, true, 0>(bufF, bufQ, tws, N, PITCH, 1, tid, x);
      plan_stage<NT, R2, true, false, 1>(bufQ, bufF, tws, N, PITCH, 1, tid, x);
      // ---- F: forward transform; C fused into the last column stage (Q replaces Phi' in bufF) ----
      plan_stage<NT, R1, false, true, 0>(bufF, bufQ, tws, N, 1, PITCH, tid, x);
      plan_stage<NT, R2, false, false, 0>(bufQ, bufF, tws, N, 1, PITCH, tid, x);
      plan_stage<NT, R1, false, true, 0>(bufF, bufQ, tws, N, PITCH, 1, tid, x);
      plan_stage<NT, R2, false, false, 2>(bufQ, bufF, tws, N, PITCH, 1, tid, x);
    } else {
      // ---- I: inverse transform, rows then columns (2 * nrad stages: the result is back in bufF) ----
      lines_fft(std::true_type{}, fa, fb, 1, PITCH);
      lines_fft(std::true_type{}, fa, fb, PITCH, 1);

      // ---- M: amplitude replacement ----
      {
        const float* __restrict__ inv_i = stack + (size_t)slot * NN;
        int i = ti0, j = tj0;
        for (int t = tid; t < NN; t += UN * NT) {
          float ii[UN];
          int fo[UN];
  #pragma unroll
          for (int k = 0; k < UN; ++k) {
            fo[k] = i * PITCH + j;
            if (t + k * NT < NN) ii[k] = __ldg(inv_i + t + k * NT);
            step(i, j);
          }
  #pragma unroll
          for (int k = 0; k < UN; ++k)
            if (t + k * NT < NN) {
              const float2 v = fa[fo[k]];
              const float2 tt = make_float2(v.x + epsr, v.y + epsi);
              const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * ii[k]);     // sqrt(I)/|psi+eps|; I = 0 -> 0
              fa[fo[k]] = make_float2(v.x * sc, v.y * sc);
            }
        }
      }
      __syncthreads();

      // ---- F: forward transform ----
      lines_fft(std::false_type{}, fa, fb, 1, PITCH);
      lines_fft(std::false_type{}, fa, fb, PITCH, 1);

      // ---- C: object update (old pupil), Q from the old window ----
      {
        int i = ti0, j = tj0;
        for (int t = tid; t < NN; t += UN * NT) {
          float2 Ovs[UN], Pvs[UN];
          float sup[UN];
          int fo[UN], oo[UN];
  #pragma unroll
          for (int k = 0; k < UN; ++k) {
            fo[k] = i * PITCH + j;
            oo[k] = wrap_half(i, N) * L + wrap_half(j, N);
            if (t + k * NT < NN) { Ovs[k] = O[oo[k]]; Pvs[k] = P[t + k * NT]; sup[k] = __ldg(p.support + t + k * NT); }
            step(i, j);
          }
  #pragma unroll
          for (int k = 0; k < UN; ++k)
            if (t + k * NT < NN) {
              const float2 Ov = Ovs[k], Pv = Pvs[k];
              const float2 d = csub(fa[fo[k]], cmul(Ov, Pv));
              const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
              const float2 num = cmulc(d, Pv);
              const float A = pa2 + p.delta2;
              const float sc = __fdividef(sqrt_fast(pa2) * inv_pmax, fmaf(A, A, kd2 * kd2));
              O[oo[k]] = make_float2(Ov.x + (num.x * A + num.y * kd2) * sc, Ov.y + (num.y * A - num.x * kd2) * sc);
              const float oa2 = fmaf(Ov.x, Ov.x, Ov.y * Ov.y);
              const float2 numq = cmulc(d, Ov);
              const float A1 = oa2 + p.delta1;
              const float sq = __fdividef(sqrt_fast(oa2) * sup[k], fmaf(A1, A1, kd1 * kd1));
              fb[fo[k]] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
            }
        }
      }
    }
    __syncthreads();                                           // the window's new values are visible to the whole CTA

    // ---- D: touched cells, max|objF|^2 ----
    {
      const int a0 = cr.y >> 4, a1 = (cr.y + N - 1) >> 4, b0 = cr.x >> 4, b1 = (cr.x + N - 1) >> 4;
      const int nb = b1 - b0 + 1, nc = (a1 - a0 + 1) * nb;
      for (int t = warp; t < nc; t += 2 * NW) {                // two cells per warp in flight
        const int t1 = t + NW;
        const int ca = a0 + t / nb, cb = b0 + t % nb, ca1 = a0 + t1 / nb, cb1 = b0 + t1 % nb;
        const float m0 = cell_part(ca, cb);
        const float m1 = t1 < nc ? cell_part(ca1, cb1) : 0.f;
        const float r0 = warp_max(m0), r1 = warp_max(m1);
        if (lane == 0) { U[ca * p.cgc + cb] = r0; if (t1 < nc) U[ca1 * p.cgc + cb1] = r1; }
      }
    }
    __syncthreads();
    omax2 = grid_max();
    pending = true;
    if (++slot == p.n_leds) slot = 0;
  }

  // ---- epilogue: the last pupil update ----
  if (pending) {
    const float inv_omax = rsqrt_fast(omax2);
    int i = ti0, j = tj0;
    for (int t = tid; t < NN; t += NT, step(i, j)) {
      float2 pv = P[t];
      const float2 qv = qbuf[i * PITCH + j];
      pv.x = fmaf(qv.x, inv_omax, pv.x);
      pv.y = fmaf(qv.y, inv_omax, pv.y);
      P[t] = pv;
    }
  }
}

}  // namespace fpm
