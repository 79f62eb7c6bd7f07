// fpm_general_fused.cuh -- fused sub-aperture update for tile sizes that are not powers of two but whose field fits
// one SM twice (even Np <= 112 with prime factors 2, 3, 5: the shipped dataset*.json use cropSizeX = 90 and 100).
//
// One CTA per tile, persistent over all updates of the launch (the reference's sequential LED order,
// fpmMain.cpp:350-475, is kept inside the CTA; parallelism is across tiles and inside the Np x Np transform):
//
//   A   pending pupil update P += Q / max|objF| (of the previous LED, fpmMain.cpp:470-475) fused with the window fetch
//       and Phi = O * P (:358-364); max|P|^2 reduced on the way (:415)
//   I   inverse 2-D transform, Stockham stages between two shared-memory copies of the field, rows then columns,
//       unscaled (the 1/Np^2 of ifft2 cancels in psi/|psi+eps|; eps is scaled instead).  Np = R1 * R2 with both
//       radices compiled in (plan_stage: 90 = 10 x 9, 100 = 10 x 10 ...), else radices 4,2,3,5 at run time    (:365)
//   M   psi' = psi * rsqrt(|psi+eps|^2 * (1/I))                                                             (:378-393)
//   F   forward 2-D transform                                                                                (:394)
//   C   dPhi = Phi' - O P;  O += dPhi |P| P* / D_O written to the spectrum;  Q = dPhi |O| O* / D_P * S kept on chip
//       (plans: in place of Phi' in the field buffer; run-time radices: in the scratch buffer)     (:406-447,459-472)
//   D   exact max|objF|: the 16x16-pixel cells of the grid of cell maxima that the window touches are rebuilt from
//       the spectrum, the grid (shared memory) is scanned                                                  (:460,467)
//
// The field never leaves shared memory; lanes run over LINES in every transform stage (rows: odd pitch, columns:
// adjacent addresses), so all 64-bit accesses of a half-warp fall into distinct bank pairs for any Np, and the
// butterfly index -- hence the twiddle -- is uniform over (nearly) the whole warp.
// With a compiled plan everything is pruned to the bounding box of the pupil support (P, Q and the object increment
// vanish outside it): A, C and D touch the box only, the inverse row stages run the box's rows (the other rows of
// O * P are zero), the forward column stages the box's columns.
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
  int ylo, yhi, xlo, xhi;   // bounding box of the pupil support, wrapped indices in [-N/2, N/2)
  float delta1, delta2, eps, kappa;
  long long* stage_clk;     // [16] per-stage cycle totals of CTA 0 (only with -DFPM_STAGE_TIMING)
};

__host__ __device__ inline size_t general_fused_smem_bytes(int N, int cgr, int cgc) {
  const size_t fld = (sizeof(float2) * (size_t)N * (N + 1) + 15) / 16 * 16;
  return 2 * fld + sizeof(float2) * N + sizeof(float) * ((size_t)cgr * cgc + 64) + 32;
}

// What the column stages with an epilogue need besides the field
struct StageExtra {
  const float* inv_i;                 // MODE 1: 1/I of this LED, [N][N]
  float epsr, epsi;                   // MODE 1
};

// One Stockham stage of a two-stage plan N = R1 * R2 over all N lines, src -> dst, radix R at compile time: each work
// item is one R-point transform in registers (fft_reg, composite radices 6 / 9 / 10 included).  FIRST: radix R1, no
// twiddles, outputs contiguous (j*R + k); otherwise radix R2, inputs twiddled by W_N^(r*j), outputs at j + k*R1.
// Work item t = j * nl + li over the nl lines l0 .. l0+nl-1 (wrapped indices; all lines: l0 = 0, nl = N): lanes over
// lines.  Lines outside the bounding box of the pupil support are skipped where they are known to be zero (row
// transforms of O * P) or not needed (column transforms of Phi'); ZIN: samples known to be zero are not read either
// (columns outside the box in the first inverse row stage, rows outside it in the first inverse column stage), so
// nothing outside the box ever has to be cleared.  MODE 1 (column stage: es = pitch, ls = 1, lanes run over columns
// and the global loads are coalesced) fuses M (amplitude replacement, fpmMain.cpp:378-393) into the stores, its 1/I
// operands requested before the butterflies.  (C runs as its own pass over the box: fused into the last forward
// stage it was issue-bound on an unbalanced 1.2 rounds of work items, 13 k cycles against 2 k + 3 k separately.)
template <int NT, int R, bool INV, bool FIRST, int MODE, bool ZIN = false>
__device__ __forceinline__ void plan_stage(const float2* __restrict__ src, float2* __restrict__ dst,
                                           const float2* __restrict__ tws, int N, int es, int ls, int tid,
                                           int l0, int nl, const StageExtra& x, int z0 = 0, int nz = 0) {
  // (padding nl to a multiple of 16 in the work-item index, so that no half-warp straddles two values of j, removes
  // the bank-conflict replays -- 19 % of the wavefronts -- but the idle lanes cost more: -4 % measured)
  const int T = N / R, total = nl * T;
  const int qNT = NT / nl, rNT = NT % nl;
  int j = tid / nl, li = tid % nl;
  for (int t = tid; t < total; t += NT) {
    int l = l0 + li;                                     // wrapped line index -> natural
    if (l < 0) l += N;
    float2 v[R];
    float ii[MODE == 1 ? R : 1];
    if constexpr (MODE == 1) {
      static_for<0, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        ii[i] = __ldg(x.inv_i + (j + radix_out<R>(i) * T) * N + l);
      });
    }
    const float2* s = src + l * ls + j * es;
    if constexpr (ZIN) {
      // samples whose (wrapped) index along the line lies outside [z0, z0 + nz) are zero by construction (O * P outside
      // the box): not read -- whatever the buffer holds there is stale
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool in = (unsigned)(wrap_half(j + r * T, N) - z0) < (unsigned)nz;
        v[r] = make_float2(0.f, 0.f);
        if (in) v[r] = s[r * T * es];
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = s[r * T * es];
    }
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
      d[(FIRST ? ko : ko * T) * es] = val;
    });
    j += qNT; li += rNT;
    if (li >= nl) { li -= nl; ++j; }
  }
  __syncthreads();
}

// R1 * R2 == Np: two-stage plan with compile-time radices (M fused into its column stage, pointwise passes and the
// row / column transforms pruned to the bounding box of the pupil support, the pupil increment Q shares the field
// buffer); R1 == 0: radices from p.rad at run time, every step its own pass.
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

  // bounding box of the pupil support (PLAN: P, Q and the object increment are zero outside it and never touched)
  const int NRb = p.yhi - p.ylo + 1, NCb = p.xhi - p.xlo + 1, nbox = NRb * NCb;
  // element index stepping t -> (i, j) = (t / N, t % N) without dividing inside the loops
  const int qNT = NT / N, rNT = NT % N, ti0 = tid / N, tj0 = tid % N;
  auto step = [&](int& q, int& r) { q += qNT; r += rNT; if (r >= N) { r -= N; ++q; } };

  // per-lane partial maximum of the 16x16 cell (a, b) from the spectrum, one warp: lanes = 2 rows x 16 columns per load
  auto cell_part = [&](int a, int b) -> float {
    // cells cut by the spectrum border: indices clamped (a pixel counted twice does not change the maximum), so the
    // eight loads are unconditional and in flight together
    const int c = min((b << 4) + (lane & 15), L - 1);
    float2 o[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = min((a << 4) + 2 * rr + (lane >> 4), L - 1);
      o[rr] = __ldcg(objFc + (size_t)r * L + c);
    }
    float m = 0.f;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) m = fmaxf(m, fmaf(o[rr].x, o[rr].x, o[rr].y * o[rr].y));
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
#ifdef FPM_STAGE_TIMING
  long long tacc_[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) tacc_[k] = 0;
  long long tprev_ = clock64();
#endif
  for (int u = 0; u < p.n_updates; ++u) {
    const short2 cr = p.crop[slot];
    float2* O = objFc + (size_t)(cr.y + H) * L + (cr.x + H);           // window origin: wrapped indices -H .. H-1
    float2 *fa = bufF, *fb = bufQ;

    // ---- A: (P += Q / max|objF|), Phi = O * P, max|P|^2 ----
    {
      const float inv_omax = pending ? rsqrt_fast(omax2) : 0.f;
      float pm2 = 0.f;
      if constexpr (PLAN) {
        // batches of 8 elements per thread in flight (their global loads are issued together) while more than 4 per
        // thread remain, then batches of 4: 61 x 61 box = one batch of 8, 75 x 75 = 8 + 4
        const int qb = NT / NCb, rb = NT % NCb;
        int bi = tid / NCb, bj = tid % NCb;
        auto batch = [&](auto ua_tag, int t) {
          constexpr int UA = decltype(ua_tag)::value;
          float2 pv[UA], ov[UA];
          int fo[UA], pe[UA];
#pragma unroll
          for (int k = 0; k < UA; ++k) {
            const bool in = t + k * NT < nbox;                  // past the end: the box's first element, loaded and dropped
            const int iw = p.ylo + (in ? bi : 0), jw = p.xlo + (in ? bj : 0);
            const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
            fo[k] = in ? i * PITCH + j : -1;
            pe[k] = i * N + j;
            pv[k] = P[pe[k]];
            ov[k] = O[iw * L + jw];
            bi += qb; bj += rb;
            if (bj >= NCb) { bj -= NCb; ++bi; }
          }
#pragma unroll
          for (int k = 0; k < UA; ++k)
            if (fo[k] >= 0) {
              if (pending) {
                const float2 qv = bufF[fo[k]];
                pv[k].x = fmaf(qv.x, inv_omax, pv[k].x);
                pv[k].y = fmaf(qv.y, inv_omax, pv[k].y);
                P[pe[k]] = pv[k];
              }
              pm2 = fmaxf(pm2, fmaf(pv[k].x, pv[k].x, pv[k].y * pv[k].y));
              bufF[fo[k]] = cmul(ov[k], pv[k]);
            }
        };
        int t = tid;
        for (; nbox - (t - tid) > 4 * NT; t += 8 * NT) batch(std::integral_constant<int, 8>{}, t);
        for (; t < nbox; t += 4 * NT) batch(std::integral_constant<int, 4>{}, t);
      } else {
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
      }
      pm2 = warp_max(pm2);
      if (lane == 0) red[warp] = pm2;
      FPM_TICK(11);
    }
    __syncthreads();
    const float inv_pmax = rsqrt_fast(warp_max(red[lane % NW]));
    FPM_TICK(1);

    if constexpr (PLAN) {
      StageExtra x;
      x.inv_i = stack + (size_t)slot * NN; x.epsr = epsr; x.epsi = epsi;
      // ---- I: inverse transform, rows (only those of the bounding box: the others are zero) then columns; M fused
      //         into the last column stage ----
      plan_stage<NT, R1, true, true, 0, true>(bufF, bufQ, tws, N, 1, PITCH, tid, p.ylo, NRb, x, p.xlo, NCb); FPM_TICK(2);
      plan_stage<NT, R2, true, false, 0>(bufQ, bufF, tws, N, 1, PITCH, tid, p.ylo, NRb, x); FPM_TICK(3);
      plan_stage<NT, R1, true, true, 0, true>(bufF, bufQ, tws, N, PITCH, 1, tid, 0, N, x, p.ylo, NRb); FPM_TICK(4);
      plan_stage<NT, R2, true, false, 1>(bufQ, bufF, tws, N, PITCH, 1, tid, 0, N, x); FPM_TICK(5);
      // ---- F: forward transform, rows then the columns of the bounding box (outside the box bufF is stale from here
      //         on and not read until the next M rewrites it) ----
      plan_stage<NT, R1, false, true, 0>(bufF, bufQ, tws, N, 1, PITCH, tid, 0, N, x); FPM_TICK(6);
      plan_stage<NT, R2, false, false, 0>(bufQ, bufF, tws, N, 1, PITCH, tid, 0, N, x); FPM_TICK(7);
      plan_stage<NT, R1, false, true, 0>(bufF, bufQ, tws, N, PITCH, 1, tid, p.xlo, NCb, x); FPM_TICK(8);
      plan_stage<NT, R2, false, false, 0>(bufQ, bufF, tws, N, PITCH, 1, tid, p.xlo, NCb, x); FPM_TICK(9);
      // ---- C: object update (old pupil) written to the spectrum, Q from the old window in place of Phi' ----
      {
        const int qb = NT / NCb, rb = NT % NCb;
        int bi = tid / NCb, bj = tid % NCb;
        auto batch = [&](auto uc_tag, int t) {
          constexpr int UC = decltype(uc_tag)::value;
          float2 Ovs[UC], Pvs[UC];
          float sup[UC];
          int fo[UC], oo[UC];
#pragma unroll
          for (int k = 0; k < UC; ++k) {
            const bool in = t + k * NT < nbox;
            const int iw = p.ylo + (in ? bi : 0), jw = p.xlo + (in ? bj : 0);
            const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
            fo[k] = in ? i * PITCH + j : -1;
            oo[k] = iw * L + jw;
            Ovs[k] = O[oo[k]];
            Pvs[k] = P[i * N + j];
            sup[k] = __ldg(p.support + i * N + j);
            bi += qb; bj += rb;
            if (bj >= NCb) { bj -= NCb; ++bi; }
          }
#pragma unroll
          for (int k = 0; k < UC; ++k)
            if (fo[k] >= 0) {
              const float2 Ov = Ovs[k], Pv = Pvs[k];
              const float2 d = csub(bufF[fo[k]], cmul(Ov, Pv));
              const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
              const float2 num = cmulc(d, Pv);
              const float A = pa2 + p.delta2;
              const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
              O[oo[k]] = make_float2(Ov.x + (num.x * A + num.y * kd2) * sc, Ov.y + (num.y * A - num.x * kd2) * sc);
              const float oa2 = fmaf(Ov.x, Ov.x, Ov.y * Ov.y);
              const float2 numq = cmulc(d, Ov);
              const float A1 = oa2 + p.delta1;
              const float sq = sqrt_fast(oa2) * sup[k] * rcp_fast(fmaf(A1, A1, kd1 * kd1));
              bufF[fo[k]] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
            }
        };
        int t = tid;
        for (; nbox - (t - tid) > 4 * NT; t += 8 * NT) batch(std::integral_constant<int, 8>{}, t);
        for (; t < nbox; t += 4 * NT) batch(std::integral_constant<int, 4>{}, t);
      }
      FPM_TICK(14);
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
              const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
              O[oo[k]] = make_float2(Ov.x + (num.x * A + num.y * kd2) * sc, Ov.y + (num.y * A - num.x * kd2) * sc);
              const float oa2 = fmaf(Ov.x, Ov.x, Ov.y * Ov.y);
              const float2 numq = cmulc(d, Ov);
              const float A1 = oa2 + p.delta1;
              const float sq = sqrt_fast(oa2) * sup[k] * rcp_fast(fmaf(A1, A1, kd1 * kd1));
              fb[fo[k]] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
            }
        }
      }
    }
    __syncthreads();                                           // the window's new values are visible to the whole CTA

    // ---- D: touched cells, max|objF|^2 ----
    {
      // rows / columns of the spectrum this update wrote: the window (PLAN: its part inside the bounding box)
      const int wy0 = PLAN ? cr.y + H + p.ylo : cr.y, wy1 = PLAN ? cr.y + H + p.yhi : cr.y + N - 1;
      const int wx0 = PLAN ? cr.x + H + p.xlo : cr.x, wx1 = PLAN ? cr.x + H + p.xhi : cr.x + N - 1;
      const int a0 = wy0 >> 4, a1 = wy1 >> 4, b0 = wx0 >> 4, b1 = wx1 >> 4;
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
    FPM_TICK(12);
    __syncthreads();
    FPM_TICK(13);
    omax2 = grid_max();
    pending = true;
    if (++slot == p.n_leds) slot = 0;
    FPM_TICK(10);
  }
#ifdef FPM_STAGE_TIMING
  if (tid == 0 && blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p.stage_clk[k] += tacc_[k];
  }
#endif

  // ---- epilogue: the last pupil update ----
  if (pending) {
    const float inv_omax = rsqrt_fast(omax2);
    if constexpr (PLAN) {
      for (int t = tid; t < NRb * NCb; t += NT) {
        const int iw = p.ylo + t / NCb, jw = p.xlo + t % NCb;
        const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
        float2 pv = P[i * N + j];
        const float2 qv = bufF[i * PITCH + j];
        pv.x = fmaf(qv.x, inv_omax, pv.x);
        pv.y = fmaf(qv.y, inv_omax, pv.y);
        P[i * N + j] = pv;
      }
    } else {
      int i = ti0, j = tj0;
      for (int t = tid; t < NN; t += NT, step(i, j)) {
        float2 pv = P[t];
        const float2 qv = bufQ[i * PITCH + j];
        pv.x = fmaf(qv.x, inv_omax, pv.x);
        pv.y = fmaf(qv.y, inv_omax, pv.y);
        P[t] = pv;
      }
    }
  }
}

}  // namespace fpm
