// fpm_pruned_fused.cuh -- fused sub-aperture update for tiles whose FIELD does not fit one SM but whose pupil box does:
// Np = R1 * R2 with both radices compiled in (the shipped dataset_dogStomach.json: cropSizeX = 200 = 20 x 10, Nlarge 600,
// pupil box 53 x 53 -- the one size the reference's own profile was taken at, output.svg:498).
//
// One CTA per tile, persistent over all updates of the launch (sequential LED order of fpmMain.cpp:350-475 inside the
// CTA).  The Np x Np field (320 KB at Np = 200) is NEVER materialised:
//   * Phi = O * P is zero outside the bounding box of the pupil support (NRb x NCb), so the inverse ROW transform
//     runs on the NRb box rows only                                   -> X [NRb][Np]   (85 KB)
//   * the inverse COLUMN transform, the amplitude replacement (pointwise) and the forward COLUMN transform touch one
//     column at a time: a batch of CB columns goes X -> S [Np][CB] -> X through shared memory, and only the box's rows
//     of the forward result are kept (the object / pupil updates need Phi' on the box only)
//   * the forward ROW transform runs on the NRb rows again and keeps the box's columns.
// Transforms are two-stage in place: decimation in frequency for the inverse (natural order in, digit-scrambled
// position p = R2*k1 + k2 <-> index k1 + R1*k2 out), decimation in time for the forward (scrambled in, natural out), so
// no reordering pass exists; the stage in the middle of the column phase does inverse radix-R2, amplitude replacement
// and forward radix-R2 on the same registers.  1/I is stored position-major ([pos(y)][pos(x)], stack_convert_general
// with R1 > 0) so that the amplitude stage reads it coalesced.
//
//   A    pending P += Q / max|objF| (fpmMain.cpp:470-475 of the previous LED) + window fetch + Phi = O * P (:358-364)
//   IR   inverse rows, stages A (radix R1, inputs outside the box taken as zero) and B (radix R2)          (:365)
//   per batch of CB columns:
//     IC-A   inverse columns stage A: X (box rows) -> S                                                  (:365)
//     MID    inverse stage B, psi' = psi * rsqrt(|psi+eps|^2 * (1/I)), forward stage B'                  (:378-394)
//     FC-A'  forward columns stage A': S -> X (box rows only)                                            (:394)
//   FR   forward rows, stages B' and A' (box columns kept)                                               (:394)
//   C    dPhi = Phi' - O P;  O += dPhi |P| P* / D_O to the spectrum;  Q = dPhi |O| O* / D_P * S in X      (:406-447,459-472)
//   D    exact max|objF|: the 16x16 cells the box touched are rebuilt from the spectrum, grid scan        (:460,467)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "fft_regs.cuh"
#include "fpm_update.cuh"
#include "fpm_general.cuh"

namespace fpm {

struct PrunedParams {
  float2* objFc;            // [n_tiles][L][L] centred
  float2* pupil;            // [n_tiles][N][N] DC-at-corner
  const float* stack;       // [n_tiles][n_leds][N][N] 1/I, position-major: [pos(y)][pos(x)], pos(v) = R2*(v%R1) + v/R1
  const float* support;     // [N][N]
  const short2* crop;       // [n_leds]
  const float2* tw;         // [N] exp(-2*pi*i*k/N)
  int L, n_leds, tile0;
  int slot_begin, n_updates;
  int cgr, cgc;             // grid of 16x16-pixel max-cells over the spectrum (edge cells partial)
  int ylo, yhi, xlo, xhi;   // bounding box of the pupil support, wrapped indices in [-N/2, N/2)
  int cb;                   // columns per batch of the column phase
  float delta1, delta2, eps, kappa;
  long long* stage_clk;     // [16] per-stage cycle totals of CTA 0 (only with -DFPM_STAGE_TIMING)
};

__host__ __device__ inline size_t pruned_fused_smem_bytes(int N, int nrb, int cb, int cgr, int cgc) {
  return sizeof(float2) * ((size_t)nrb * (N + 1) + (size_t)N * cb + N) + sizeof(float) * ((size_t)cgr * cgc + 64) + 64;
}

#ifndef FPM_PRUNED_ILP
#define FPM_PRUNED_ILP 1      // work items per thread in flight in the B / B' / MID stages.  Measured on B200 (Np = 200, 148
                              // tiles): 2 items in flight spill at the 96-register cap of 17-20 warps and lose 15-20 %
                              // (640 threads: 4.62 M updates/s with 1, 3.59 M with 2; 512 threads: 4.44 M / 4.06 M)
#endif

template <int N> __device__ __forceinline__ int wrap_half_c(int i) { return (i < N / 2) ? i : i - N; }

// One in-place stage of the two-stage transforms over `nl` lines (lanes run over lines: element stride es, line
// stride ls, in float2).  Work item t = j * nl + li.
//   KIND 0  DIF stage A (inverse): radix R1 over src[n2 + R2*n1] (n2 = j), twiddle W^(n2*k1), -> dst[n2 + R2*k1]
//           ZSRC: src holds only the samples whose wrapped index lies in [z0, z0+nz), compacted (index - z0); the
//           others are zero (rows of X in the column phase) -- else src is full and samples outside [z0, z0+nz) are
//           stale and taken as zero (columns of X in the row phase)
//   KIND 1  DIF stage B (inverse): radix R2 over [R2*k1 + n2] (k1 = j), in place
//   KIND 2  DIT stage B' (forward): radix R2 over [R2*k1 + k2], twiddle W^(k1*q2), in place
//   KIND 3  DIT stage A' (forward): radix R1 over src[q2 + R2*k1] (q2 = j) -> natural index q2 + R2*q1, stored only
//           where the wrapped index lies in [z0, z0+nz); ZDST: compacted (index - z0)
//   KIND 4  KIND 1 + amplitude replacement + KIND 2 on the same registers (the middle of the column phase);
//           inv_i points at the 1/I row of line 0: element (position row r, line li) at inv_i[r * N + li]
// NARROW (R1 = 20): the box [z0, z0+nz) lies within +-3*R2 of the origin, so a stage-A butterfly has at most the six
// samples n1 = 0, 1, 2, R1-3, R1-2, R1-1 inside it: KIND 0 loads those six (unconditionally, clamped + masked) and
// runs fft_reg_in6, KIND 3 computes those six outputs only (fft_reg_out6).
// Stages B / B' / MID can run two work items per thread at a time (FPM_PRUNED_ILP = 2: loads issued together, two
// independent butterfly streams); off by default, see above.
template <int NT, int R1, int R2, int KIND, bool ZSRC, bool ZDST, bool NARROW>
__device__ __forceinline__ void pruned_stage(const float2* __restrict__ src, float2* __restrict__ dst,
                                             const float2* __restrict__ tws, int es_src, int es_dst, int ls, int tid, int nl,
                                             int z0, int nz, const float* __restrict__ inv_i, float epsr, float epsi) {
  constexpr int N = R1 * R2;
  constexpr bool STAGE_A = (KIND == 0 || KIND == 3);
  constexpr int R = STAGE_A ? R1 : R2;
  constexpr int J = STAGE_A ? R2 : R1;               // butterflies per line
  constexpr bool INV = (KIND == 0 || KIND == 1 || KIND == 4);
  const int total = nl * J;
  const int qNT = NT / nl, rNT = NT % nl;
  int j = tid / nl, li = tid % nl;
  auto advance = [&](int& jj, int& ll) { jj += qNT; ll += rNT; if (ll >= nl) { ll -= nl; ++jj; } };
  if constexpr (STAGE_A) {
    for (int t = tid; t < total; t += NT) {
      float2 v[R];
      if constexpr (KIND == 0) {
        const float2* s = src + li * ls;
        if constexpr (NARROW) {
          float2 w6[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const int n1 = q < 3 ? q : R - 6 + q;
            const int w = wrap_half_c<N>(j + R2 * n1) - z0;
            const bool in = (unsigned)w < (unsigned)nz;
            const float2 x = s[(ZSRC ? (in ? w : 0) : j + R2 * n1) * es_src];         // always a valid address
            w6[q] = in ? x : make_float2(0.f, 0.f);
          }
          fft_reg_in6<R, INV>(w6, v);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int w = wrap_half_c<N>(j + R2 * r) - z0;
            v[r] = make_float2(0.f, 0.f);
            if ((unsigned)w < (unsigned)nz) v[r] = s[(ZSRC ? w : j + R2 * r) * es_src];
          }
          fft_reg<R, INV>(v);
        }
        float2* d = dst + li * ls + j * es_dst;
        static_for<0, R>([&](auto I) {
          constexpr int i = decltype(I)::value, k1 = radix_out<R>(i);
          float2 val = v[i];
          if constexpr (k1 > 0) val = twmul<INV>(val, tws[j * k1]);
          d[R2 * k1 * es_dst] = val;
        });
      } else {                                          // KIND 3
        const float2* s = src + li * ls + j * es_src;
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = s[R2 * r * es_src];
        float2* d = dst + li * ls;
        if constexpr (NARROW) {
          float2 o6[6];
          fft_reg_out6<R, false>(v, o6);
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const int q1 = q < 3 ? q : R - 6 + q;
            const int w = wrap_half_c<N>(j + R2 * q1) - z0;
            if ((unsigned)w < (unsigned)nz) d[(ZDST ? w : j + R2 * q1) * es_dst] = o6[q];
          }
        } else {
          fft_reg<R, false>(v);
          static_for<0, R>([&](auto I) {
            constexpr int i = decltype(I)::value, q1 = radix_out<R>(i);
            const int w = wrap_half_c<N>(j + R2 * q1) - z0;
            if ((unsigned)w < (unsigned)nz) d[(ZDST ? w : j + R2 * q1) * es_dst] = v[i];
          });
        }
      }
      advance(j, li);
    }
  } else {
    // two items per thread in flight: (j, li) and the item NT further on
    auto load = [&](int jj, int ll, float2 (&v)[R], float (&ii)[R]) {
      const float2* d = dst + ll * ls + R2 * jj * es_dst;
      if constexpr (KIND == 4) {
#pragma unroll
        for (int r = 0; r < R; ++r) ii[r] = __ldg(inv_i + (R2 * jj + r) * N + ll);    // position row R2*k1 + k2, coalesced over lines
      }
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = d[r * es_dst];
    };
    auto compute_store = [&](int jj, int ll, float2 (&v)[R], float (&ii)[R]) {
      float2* d = dst + ll * ls + R2 * jj * es_dst;
      fft_reg<R, INV>(v);
      if constexpr (KIND == 4) {
        float2 w[R];
        static_for<0, R>([&](auto I) {
          constexpr int i = decltype(I)::value, k2 = radix_out<R>(i);
          const float2 val = v[i];
          const float2 tt = make_float2(val.x + epsr, val.y + epsi);
          const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * ii[k2]);        // sqrt(I)/|psi+eps|; I = 0 -> 0
          w[k2] = make_float2(val.x * sc, val.y * sc);
        });
        fft_reg<R, false>(w);
        static_for<0, R>([&](auto I) {
          constexpr int i = decltype(I)::value, q2 = radix_out<R>(i);
          float2 val = w[i];
          if constexpr (q2 > 0) val = twmul<false>(val, tws[jj * q2]);
          d[q2 * es_dst] = val;
        });
      } else {
        static_for<0, R>([&](auto I) {
          constexpr int i = decltype(I)::value, k2 = radix_out<R>(i);
          float2 val = v[i];
          if constexpr (KIND == 2 && k2 > 0) val = twmul<false>(val, tws[jj * k2]);
          d[k2 * es_dst] = val;
        });
      }
    };
    if constexpr (FPM_PRUNED_ILP == 2) {
      for (int t = tid; t < total; t += 2 * NT) {
        int j2 = j, li2 = li;
        advance(j2, li2);
        const bool two = t + NT < total;
        float2 va[R], vb[R];
        float ia[R], ib[R];
        load(j, li, va, ia);
        if (two) load(j2, li2, vb, ib);
        compute_store(j, li, va, ia);
        if (two) compute_store(j2, li2, vb, ib);
        j = j2; li = li2;
        advance(j, li);
      }
    } else {
      for (int t = tid; t < total; t += NT) {
        float2 va[R];
        float ia[R];
        load(j, li, va, ia);
        compute_store(j, li, va, ia);
        advance(j, li);
      }
    }
  }
  __syncthreads();
}

template <int NT, int R1, int R2, bool NARROW>
__global__ void __launch_bounds__(NT, 1) fpm_update_pruned_kernel(const __grid_constant__ PrunedParams p) {
  constexpr int N = R1 * R2, H = N / 2, PX = N + 1, NN = N * N, NW = NT / 32;
  static_assert(NT % 32 == 0 && N % 2 == 0, "whole warps, even tile edge");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = p.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NRb = p.yhi - p.ylo + 1, NCb = p.xhi - p.xlo + 1, nbox = NRb * NCb, CB = p.cb;
  float2* X = reinterpret_cast<float2*>(smem_raw);                        // [NRb][PX]: box rows of the field / Phi' / Q
  float2* S = X + (size_t)NRb * PX;                                       // [N][CB]: one batch of columns
  float2* tws = S + (size_t)N * CB;
  float* red = reinterpret_cast<float*>(tws + N);                         // [0..31] |P|^2 partials, [32..63] |objF|^2
  float* U = red + 64;                                                    // [cgr][cgc] cell maxima of |objFc|^2

  const int tile = p.tile0 + blockIdx.x;
  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* P = p.pupil + (size_t)tile * NN;
  const float* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * NN;
  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  const float epsr = p.eps * (float)NN, epsi = p.kappa * epsr;

  // per-lane partial maximum of the 16x16 cell (a, b) from the spectrum, one warp: lanes = 2 rows x 16 columns per load
  auto cell_part = [&](int a, int b) -> float {
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
  const int qb = NT / NCb, rb = NT % NCb;
  for (int u = 0; u < p.n_updates; ++u) {
    const short2 cr = p.crop[slot];
    float2* O = objFc + (size_t)(cr.y + H) * L + (cr.x + H);           // window origin: wrapped indices -H .. H-1
    // the next LED's 1/I image towards L2 while this update computes (the amplitude stage reads it with plain loads)
    if (tid < 8 && u + 1 < p.n_updates) {
      const int nslot = slot + 1 == p.n_leds ? 0 : slot + 1;
      constexpr unsigned chunk = (unsigned)(NN * 4 / 8) & ~15u;        // eight pieces, 16-byte multiples (the tail is not prefetched)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(stack + (size_t)nslot * NN) + (size_t)tid * chunk),
                   "r"(chunk) : "memory");
    }

    // ---- A: (P += Q / max|objF|), Phi = O * P on the box -> X, max|P|^2 ----
    {
      const float inv_omax = pending ? rsqrt_fast(omax2) : 0.f;
      float pm2 = 0.f;
      int bi = tid / NCb, bj = tid % NCb;
      auto batch = [&](auto ua_tag, int t) {
        constexpr int UA = decltype(ua_tag)::value;
        float2 pv[UA], ov[UA];
        int fo[UA], pe[UA];
#pragma unroll
        for (int k = 0; k < UA; ++k) {
          const bool in = t + k * NT < nbox;                  // past the end: the box's first element, loaded and dropped
          const int br = in ? bi : 0, bc = in ? bj : 0;
          const int iw = p.ylo + br, jw = p.xlo + bc;
          const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
          fo[k] = in ? br * PX + j : -1;
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
              const float2 qv = X[fo[k]];
              pv[k].x = fmaf(qv.x, inv_omax, pv[k].x);
              pv[k].y = fmaf(qv.y, inv_omax, pv[k].y);
              P[pe[k]] = pv[k];
            }
            pm2 = fmaxf(pm2, fmaf(pv[k].x, pv[k].x, pv[k].y * pv[k].y));
            X[fo[k]] = cmul(ov[k], pv[k]);
          }
      };
      int t = tid;
      for (; nbox - (t - tid) > 4 * NT; t += 8 * NT) batch(std::integral_constant<int, 8>{}, t);
      for (; t < nbox; t += 4 * NT) batch(std::integral_constant<int, 4>{}, t);
      pm2 = warp_max(pm2);
      if (lane == 0) red[warp] = pm2;
    }
    __syncthreads();
    const float inv_pmax = rsqrt_fast(warp_max(red[lane % NW]));
    FPM_TICK(1);

    // ---- IR: inverse rows of the box (lanes over rows: element stride 1, line stride PX) ----
    pruned_stage<NT, R1, R2, 0, false, false, NARROW>(X, X, tws, 1, 1, PX, tid, NRb, p.xlo, NCb, nullptr, 0.f, 0.f); FPM_TICK(2);
    pruned_stage<NT, R1, R2, 1, false, false, NARROW>(X, X, tws, 1, 1, PX, tid, NRb, 0, 0, nullptr, 0.f, 0.f); FPM_TICK(3);

    // ---- column phase, CB column positions at a time (lanes over columns: line stride 1) ----
    const float* __restrict__ inv_led = stack + (size_t)slot * NN;
    for (int c0 = 0; c0 < N; c0 += CB) {
      const int ncb = min(CB, N - c0);
      pruned_stage<NT, R1, R2, 0, true, false, NARROW>(X + c0, S, tws, PX, CB, 1, tid, ncb, p.ylo, NRb, nullptr, 0.f, 0.f); FPM_TICK(4);
      pruned_stage<NT, R1, R2, 4, false, false, NARROW>(S, S, tws, CB, CB, 1, tid, ncb, 0, 0, inv_led + c0, epsr, epsi); FPM_TICK(5);
      pruned_stage<NT, R1, R2, 3, false, true, NARROW>(S, X + c0, tws, CB, PX, 1, tid, ncb, p.ylo, NRb, nullptr, 0.f, 0.f); FPM_TICK(6);
    }

    // ---- FR: forward rows of the box, box columns kept ----
    pruned_stage<NT, R1, R2, 2, false, false, NARROW>(X, X, tws, 1, 1, PX, tid, NRb, 0, 0, nullptr, 0.f, 0.f); FPM_TICK(7);
    pruned_stage<NT, R1, R2, 3, false, false, NARROW>(X, X, tws, 1, 1, PX, tid, NRb, p.xlo, NCb, nullptr, 0.f, 0.f); FPM_TICK(8);

    // ---- C: object update (old pupil) written to the spectrum, Q from the old window in place of Phi' ----
    {
      int bi = tid / NCb, bj = tid % NCb;
      auto batch = [&](auto uc_tag, int t) {
        constexpr int UC = decltype(uc_tag)::value;
        float2 Ovs[UC], Pvs[UC];
        float sup[UC];
        int fo[UC], oo[UC];
#pragma unroll
        for (int k = 0; k < UC; ++k) {
          const bool in = t + k * NT < nbox;
          const int br = in ? bi : 0, bc = in ? bj : 0;
          const int iw = p.ylo + br, jw = p.xlo + bc;
          const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
          fo[k] = in ? br * PX + j : -1;
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
            const float2 d = csub(X[fo[k]], cmul(Ov, Pv));
            const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
            const float2 num = cmulc(d, Pv);
            const float A = pa2 + p.delta2;
            const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
            O[oo[k]] = make_float2(Ov.x + (num.x * A + num.y * kd2) * sc, Ov.y + (num.y * A - num.x * kd2) * sc);
            const float oa2 = fmaf(Ov.x, Ov.x, Ov.y * Ov.y);
            const float2 numq = cmulc(d, Ov);
            const float A1 = oa2 + p.delta1;
            const float sq = sqrt_fast(oa2) * sup[k] * rcp_fast(fmaf(A1, A1, kd1 * kd1));
            X[fo[k]] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
          }
      };
      int t = tid;
      for (; nbox - (t - tid) > 4 * NT; t += 8 * NT) batch(std::integral_constant<int, 8>{}, t);
      for (; t < nbox; t += 4 * NT) batch(std::integral_constant<int, 4>{}, t);
    }
    FPM_TICK(9);
    __syncthreads();                                           // the window's new values are visible to the whole CTA

    // ---- D: touched cells, max|objF|^2 ----
    {
      const int wy0 = cr.y + H + p.ylo, wy1 = cr.y + H + p.yhi, wx0 = cr.x + H + p.xlo, wx1 = cr.x + H + p.xhi;
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
    __syncthreads();
    FPM_TICK(10);
    omax2 = grid_max();
    pending = true;
    if (++slot == p.n_leds) slot = 0;
    FPM_TICK(11);
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
    for (int t = tid; t < nbox; t += NT) {
      const int br = t / NCb, bc = t % NCb;
      const int iw = p.ylo + br, jw = p.xlo + bc;
      const int i = iw < 0 ? iw + N : iw, j = jw < 0 ? jw + N : jw;
      float2 pv = P[i * N + j];
      const float2 qv = X[br * PX + j];
      pv.x = fmaf(qv.x, inv_omax, pv.x);
      pv.y = fmaf(qv.y, inv_omax, pv.y);
      P[i * N + j] = pv;
    }
  }
}

}  // namespace fpm
