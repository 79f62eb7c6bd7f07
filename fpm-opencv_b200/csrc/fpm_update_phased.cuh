// fpm_update_phased.cuh -- the fused sub-aperture update, three phases per update (sm_100a).
//
// Same algorithm, data layout and results as fpm_update_kernel<N, 512, 1, true, true, true, .> (fpm_update.cuh; loop body
// fpmMain.cpp:350-475) for tiles whose field, pupil, pupil increment and windows all live in shared memory with one-row
// max-cells: N = 128 = 16 x 8 with a narrow pupil (box within +-23: six-sample stage-A butterflies, the shipped 128-pixel
// configurations and the bench workload) and N = 64 = 8 x 8 (any box).  Restructured around what the older kernel's stage
// table showed: half of its cycles went to stages with too little work for 16 warps, separated by eight block-wide
// barriers.  Here an update is three phases:
//
//   A   threads < R2*NC:  pending pupil update P += Q / max|objF|, Phi = O*P from the TMA-staged window, inverse
//                          column stage A
//       the other warps:   the max-cells of the PREVIOUS update's rectangle are rebuilt from W (off the critical path)
//   --- block barrier ---
//   B   per row block (32 rows, named barriers only): inverse column stage B of the block's k1 (the warps that own a k1
//       belong to the row block that consumes it), inverse row stages A, B, amplitude replacement, forward row stages
//       B', A', forward column stage B' of the same k1.
//       Before it, a thread per spectrum row takes the maximum of the cells this update's rectangle does NOT touch:
//       one load of the row maximum Rm for most rows, a re-read of the row of U (refreshing Rm) for the rows of this and
//       of the previous rectangle.
//   --- block barrier ---
//   C   one bbox element per lane: the last forward column stage as a direct R1-term DFT of exactly the bbox outputs,
//       fused with the object update, the pupil increment Q, the forward of the new values into the next LED's window,
//       |O_new|^2 -> W and the running maximum over rectangle + edge pixels.
//   --- block barrier ---   (TMA store of the window; max|objF| = max(untouched cells, rectangle, edges))
//
// max|objF| is exact as before (a maximum does not depend on the order it is taken in): the results equal the older
// kernel's except for the rounding of the last column stage (direct sum instead of a butterfly).
//
// Measured on the bench workload (592 tiles of 128 x 128, box 35 x 35, 157 LEDs x 10): 15.0 M updates/s against 14.5 M
// (update kernel alone).  Also measured, and not adopted: 1024 threads of 64 registers (13.6 M), twiddles of the row
// stages from constant memory instead of shared memory (14.6 - 14.8 M), the next window requested after S2 or S3
// instead of S4 (14.9 M), the two elements a thread processes back to back in phase C explicitly interleaved (all loads
// and arithmetic of both before the stores of either: 14.94 M against 14.98 M -- phase C is issue-bound, not latency-bound).
#pragma once
#include "fpm_update.cuh"

// the thread that issues the window's TMA store, waits for it and issues the next load (bulk async-groups are per
// thread): a thread of the last warp, which has the least work in phases A and C
#define FPM_TMA_TID (NT - 1)
#ifndef FPM_TICK_TID
#define FPM_TICK_TID 0        // the thread whose stage clocks the timing build reports
#endif

namespace fpm {

// acc + a * w for a table entry t = (w.x, w.y, -w.y, w.x)
__device__ __forceinline__ float2 cfma4(float2 acc, float2 a, float4 t) {
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(t.z, t.w), __ffma2_rn(make_float2(a.x, a.x), make_float2(t.x, t.y), acc));
}

template <int N> struct PhasedShape {
  static constexpr int R1 = Shape<N>::R1, R2 = Shape<N>::R2;
  static constexpr bool SIX = (R1 == 16);   // stage-A butterflies see 6 of 16 samples (box within +-(3*R2-1))
  static constexpr int TK = R1 / 2;         // entries per row of the W_R1^(r*k) table of phase C ...
  static constexpr int TP = TK + 1;         // ... padded against bank conflicts
  // Dynamic shared memory.  No static shared memory exists in the kernel, so the dynamic segment starts 1024-byte
  // aligned and every offset below is an absolute alignment.  The first part is fixed at compile time; the offsets that
  // depend on the box reach the kernel through UpdateParams::noff (constant bank: an address is `base + c[..]`, not a
  // chain of size arithmetic the register allocator has to rematerialise).
  enum { WIN0, WIN1, PC, QC, SC, WPIX, UCELL, RMAX, TOTAL, NOFF };
  static constexpr size_t off_fld = 0, off_twA = sizeof(float2) * N * (N + 1), off_twB = off_twA + sizeof(float4) * N,
                          off_red = off_twB + sizeof(float4) * N, off_T = off_red + sizeof(float) * 128,
                          off_var = off_T + sizeof(float4) * R1 * TP;
  static_assert(off_var % 128 == 0, "TMA destinations are 128-byte aligned");
  static void layout(int NR, int NC, int ocp, int L, int* off) {
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t b = off_var;
    off[WIN0] = (int)b; b += up(sizeof(float2) * (size_t)NR * ocp, 128);
    off[WIN1] = (int)b; b += up(sizeof(float2) * (size_t)NR * ocp, 128);
    off[PC] = (int)b; b += sizeof(float2) * (size_t)NR * ocp;            // P, Q, S share the window's (row, column) index
    off[QC] = (int)b; b += sizeof(float2) * (size_t)NR * ocp;
    off[SC] = (int)b; b += up(sizeof(float) * (size_t)NR * ocp, 16);
    const int tmc = (NC >> 4) + 2;
    int wsh = 0; while ((1 << wsh) < (tmc << 4)) ++wsh;
    off[WPIX] = (int)b; b += sizeof(float) * ((size_t)NR << wsh);
    off[UCELL] = (int)b; b += sizeof(float) * (size_t)L * (L >> 4);
    off[RMAX] = (int)b; b += sizeof(float) * (size_t)L;
    off[TOTAL] = (int)b;
  }
  // the boxes this kernel takes (besides: field, P, Q in shared memory, one-row max-cells, Nlarge a multiple of 64)
  static bool box_ok(int ylo, int yhi, int xlo, int xhi) {
    const int lim = 3 * R2 - 1;
    if (SIX) return ylo >= -lim && yhi <= lim && xlo >= -lim && xhi <= lim;
    return xhi - xlo + 1 <= 64 && yhi - ylo + 1 <= 64;
  }
};

template <int N, int NT>
__global__ void __launch_bounds__(NT, 1) fpm_update_phased_kernel(const __grid_constant__ UpdateParams p) {
  using S = Shape<N>;
  using PS = PhasedShape<N>;
  constexpr int R1 = S::R1, R2 = S::R2, PITCH = S::PITCH, CH = S::CH;
  constexpr int H = N / 2, NW = NT / 32, WPB = NT / N;      // WPB warps own a block of 32 rows
  constexpr bool SIX = PS::SIX;
  constexpr int NIN = SIX ? 6 : R1;                         // samples a stage-A butterfly reads / a stage-A' butterfly keeps
  constexpr int TK = PS::TK, TP = PS::TP;
  static_assert((R1 == 16 || R1 == 8) && R2 == 8, "N = 128 = 16 x 8 or N = 64 = 8 x 8");
  constexpr int WK = NW / R1;                               // warps per k1 in the column B stages
  static_assert(NW % R1 == 0 && WK * (32 / R2) == WPB, "column stage B: the warps of k1 belong to the row block that consumes it");
  static_assert(R2 % WPB == 0 && R1 % WPB == 0, "row work items per warp");
  // sample k of a stage-A butterfly is input m(k) of the radix-R1 transform, at line index (j0 or i0) + R2*m, wrapped
  auto m_of = [](int k) constexpr { return SIX ? ((k < 3) ? k : R1 - 6 + k) : k; };
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = warp / WPB, gsub = warp % WPB;
  // balanced passes (fpmb200_run): tile from a list; bit 30 of the entry = the cell maxima this kernel left in p.ucache at
  // the end of the tile's previous pass are current
  const int tile_entry = p.tile_list ? p.tile_list[blockIdx.x] : p.tile0 + blockIdx.x;
  const int tile = tile_entry & 0x3fffffff;
  const bool u_cached = p.ucache && (tile_entry & 0x40000000);
  const int L = p.L;
  const int NR = p.yhi - p.ylo + 1, NC = p.xhi - p.xlo + 1;
  const int gc = L >> 4, gc4 = gc >> 2;                    // max-cells are 1 row x 16 columns (cs == 0); float4 per U row
  const int tq = tid / NC, tr = tid - tq * NC;             // phase A work item (i0, jc)
  const int nA = R2 * NC;                                  // threads of phase A's butterflies (NC <= 64: <= NT)
  // the max-cell rebuild of phase A: the threads without a butterfly when there are enough of them, else everybody
  const bool rb_own = (NT - nA) >= 64;
  const int rb_tid = rb_own ? tid - nA : tid, rb_n = rb_own ? NT - nA : NT;
  // "side warps": the warps without a phase-A butterfly.  With at least four of them they do the whole max bookkeeping
  // during phase A -- rebuild of the previous rectangle's cells, then (named barrier among themselves) the maximum of
  // the cells this update does not touch -- and phase B starts with the transforms right away.
  const int sw0 = (nA + 31) >> 5;
  const bool side = !SIX && (NW - sw0) >= 4;               // (measured: +1.7 % at N = 64, -2.1 % at N = 128, where phase B's scan stays)
  const int side_tid = tid - 32 * sw0, side_n = 32 * (NW - sw0);

  // ---- shared memory carve-up (PS::layout) ----
  float2* const fld = reinterpret_cast<float2*>(smem_raw + PS::off_fld);
  float4* const twA = reinterpret_cast<float4*>(smem_raw + PS::off_twA);
  float4* const twB = reinterpret_cast<float4*>(smem_raw + PS::off_twB);
  float* const red = reinterpret_cast<float*>(smem_raw + PS::off_red);    // per-warp partial maxima of |.|^2:
  float* const redC = red, *const redU = red + 32, *const redP = red + 64;         //   rectangle + edges, untouched cells, pupil
  float2* const sink2 = reinterpret_cast<float2*>(red + 96);               // where the stores of switched-off lanes go
  float* const sink1 = red + 100;
  uint64_t* const wbar = reinterpret_cast<uint64_t*>(red + 104);           // completion barrier of the window TMA
  float4* const TW = reinterpret_cast<float4*>(smem_raw + PS::off_T);      // TW[r*TP + k] = W_R1^(r*k) as a cfma4 operand
  const int OCP = p.ocp;                                                   // pitch of the window AND of P, Q, S (box-relative)
  float2* const Ocb0 = reinterpret_cast<float2*>(smem_raw + p.noff[PS::WIN0]);
  float2* const Ocb1 = reinterpret_cast<float2*>(smem_raw + p.noff[PS::WIN1]);
  float2* const Pc = reinterpret_cast<float2*>(smem_raw + p.noff[PS::PC]);
  float2* const Qc = reinterpret_cast<float2*>(smem_raw + p.noff[PS::QC]);
  float* const Sc = reinterpret_cast<float*>(smem_raw + p.noff[PS::SC]);
  float* const W = reinterpret_cast<float*>(smem_raw + p.noff[PS::WPIX]);     // |.|^2 of every pixel of the touched cells
  float* const U = reinterpret_cast<float*>(smem_raw + p.noff[PS::UCELL]);    // [L][gc] exact cell maxima of |objFc|^2
  float* const Rm = reinterpret_cast<float*>(smem_raw + p.noff[PS::RMAX]);    // [L] row maxima of U (see phase B)
  const uint32_t win_bytes = (uint32_t)(sizeof(float2) * NR * OCP);

  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* Pg = p.pupil + (size_t)tile * N * N;
  const float* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * N * N;

  const int rb_row = 32 * grp + lane, rb_sub = gsub;
  auto row_block_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPB * 32) : "memory"); };

  // column stage B items of this thread: (k1 = warp, column 32*cf + lane) for every full group of 32 bbox columns, and
  // at most one of the row block's left-over items (WPB k1 x NC % 32 columns, packed over the block's warps)
  const int cb_nfull = NC >> 5, cb_nl = NC & 31;
  const int lo_i = gsub * 32 + lane;
  const bool lo_valid = lo_i < WPB * cb_nl;
  const int lo_k1 = grp * WPB + (lo_valid ? lo_i / max(cb_nl, 1) : 0), lo_jc = (cb_nfull << 5) + (lo_valid ? lo_i % max(cb_nl, 1) : 0);
  auto col_items_B = [&](auto&& body) {
    if constexpr (WK == 1) {
      for (int cf = 0; cf < cb_nfull; ++cf) body(warp, (cf << 5) + lane);
      if (lo_valid) body(lo_k1, lo_jc);
    } else {                                                // several warps per k1: 32 columns each (NC <= 47 <= 32 * WK)
      const int jc = ((warp % WK) << 5) + lane;
      if (jc < NC) body(warp / WK, jc);
    }
  };
  // phase C work items: one warp per (bbox row, 32 columns), the left-over columns packed 32 elements per warp
  const int fsh = cb_nfull >> 1;                           // (cb_nfull <= 2: full item `it` is row it >> fsh, segment it & fsh)
  const int nC_full = NR * cb_nfull, nC_items = nC_full + ((NR * cb_nl + 31) >> 5);
  const unsigned nl_mul = (65536u + (unsigned)max(cb_nl, 1) - 1u) / (unsigned)max(cb_nl, 1);   // e / cb_nl = (e * nl_mul) >> 16 (e < 2^11)

  // ---- prologue ----
  for (int t = tid; t < N; t += NT) {
    const int b = t / R2, a = t % R2;
    const float2 wa = p.tw[a * b];
    twA[t] = make_float4(wa.x, -wa.y, wa.y, wa.x);
    const int a2 = t / R1, b2 = t % R1;
    const float2 wb = p.tw[a2 * b2];
    twB[t] = make_float4(wb.x, wb.y, -wb.y, wb.x);
  }
  for (int t = tid; t < R1 * TK; t += NT) {
    const int r = t / TK, k = t % TK;
    const float2 w = p.tw[((r * k) % R1) * (N / R1)];
    TW[r * TP + k] = make_float4(w.x, w.y, -w.y, w.x);
  }
  for (int t = tid; t < NR * OCP; t += NT) {
    const int ir = t / OCP, jc = t - ir * OCP;
    const int gi = ((p.ylo + ir) & (N - 1)) * N + ((p.xlo + jc) & (N - 1));
    const bool in = jc < NC;
    Pc[t] = in ? Pg[gi] : make_float2(0.f, 0.f);
    Sc[t] = in ? p.support[gi] : 0.f;
    Qc[t] = make_float2(0.f, 0.f);
  }
  if (u_cached) {                                           // U as the previous pass over this tile left it
    const float4* src = reinterpret_cast<const float4*>(p.ucache + (size_t)tile * L * gc);
    for (int t = tid; t < L * gc4; t += NT) reinterpret_cast<float4*>(U)[t] = __ldcg(src + t);
  } else
  for (int it = warp; it < L * (L >> 5); it += NW) {       // U from the spectrum: one warp per pair of cells
    const int row = it / (L >> 5), seg = it % (L >> 5);
    const float2 o = objFc[(size_t)row * L + (seg << 5) + lane];
    const float cm = half_warp_max(fmaf(o.x, o.x, o.y * o.y), lane);
    if ((lane & 15) == 0) U[row * gc + 2 * seg + (lane >> 4)] = cm;
  }
  short2 cr_a = p.crop[p.slot_begin % p.n_leds], cr_b = p.crop[(p.slot_begin + 1) % p.n_leds];
  uint32_t wphase = 0;
  if (tid == 0) {
    mbar_init(wbar, 1);
    asm volatile("fence.proxy.async;" ::: "memory");
    mbar_expect_tx(wbar, 2 * win_bytes);
    tma_load_window(Ocb0, &p.tmap, 2 * ((cr_a.x + H + p.xlo) & ~1), cr_a.y + H + p.ylo, tile, wbar);
    tma_load_window(Ocb1, &p.tmap, 2 * ((cr_b.x + H + p.xlo) & ~1), cr_b.y + H + p.ylo, tile, wbar);
  }
  if (tid < 128) red[tid] = 0.f;
  __syncthreads();
  for (int row = tid; row < L; row += NT) {
    const float4* u4 = reinterpret_cast<const float4*>(U) + row * gc4;
    float m = 0.f;
    for (int c = 0; c < gc4; ++c) { const float4 q = u4[c]; m = fmaxf(fmaxf(m, fmaxf(q.x, q.y)), fmaxf(q.z, q.w)); }
    Rm[row] = m;
  }
  __syncthreads();

  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  const float epsr = p.eps * (float)(N * N), epsi = p.kappa * epsr;      // the 1/N^2 of ifft2 is never applied

  // rectangle of the previous update (its max-cells are rebuilt in this update's phase A)
  int pr_r0 = 0, pr_cc0 = 0, pr_ncc = 0;
#ifdef FPM_STAGE_TIMING
  long long tacc_[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) tacc_[k] = 0;
  long long tprev_ = clock64();
#endif
  for (int u = 0; u < p.n_updates; ++u) {
    const int slot = (p.slot_begin + u) % p.n_leds;
    const int xs = cr_a.x, ys = cr_a.y;
    const int nslot = (slot + 1 == p.n_leds) ? 0 : slot + 1;
    const int nslot2 = (nslot + 1 == p.n_leds) ? 0 : nslot + 1;
    const short2 cr_c = p.crop[nslot2];
    float2* const Ocur = (u & 1) ? Ocb1 : Ocb0;
    float2* Oc = Ocur + ((xs + H + p.xlo) & 1);                    // window element (ir, jc) = Oc[ir*OCP + jc]
    float2* Ocn = (u & 1) ? Ocb0 : Ocb1;                           // next window's box, box-relative columns
    if (u == 0) { mbar_wait(wbar, wphase); wphase ^= 1; }
    const float* __restrict__ img = stack + (size_t)slot * N * N;
    if (tid == 0)   // pull the next LED's intensity tile towards L2 while this update runs
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(stack + (size_t)nslot * N * N), "r"((unsigned)(N * N * 4)) : "memory");
    // this update's rectangle in the centred spectrum and the max-cells it touches
    const int r0 = ys + H + p.ylo, c0 = xs + H + p.xlo, c1 = xs + H + p.xhi;
    const int cc0 = c0 >> 4, ncc = (c1 >> 4) - cc0 + 1;
    const int wc0 = cc0 << 4, wcols = ncc << 4;
    const int wsh = 32 - __clz(wcols - 1);                          // W rows are 2^wsh floats apart

    // max|objF| after the previous update: untouched cells (scan), its rectangle and edge pixels (phase C)
    float inv_objf_max;
    {
      const float m = warp_max(fmaxf(redC[lane], redU[lane]));      // (unused entries are zero)
      inv_objf_max = (u == 0) ? 0.f : rsqrt_fast(m);                // Q == 0 before the first update
    }
    FPM_TICK(9);
    // Maximum of |objF|^2 over the cells this update does not touch.  Rm[row] = max over the cells of row `row` is
    // current for every row outside the previous rectangle; a thread owns a row: rows of the previous or of this
    // rectangle are re-read from U (refreshing Rm), the others cost one load.  Threads t0, t0 + tn, ... of whole warps.
    auto untouched_max = [&](int t0, int tn) {
      float m = 0.f;
      for (int row = t0; row < L; row += tn) {
        const bool in_cur = (unsigned)(row - r0) < (unsigned)NR;
        const bool in_prev = (u > 0) && (unsigned)(row - pr_r0) < (unsigned)NR;
        if (in_cur || in_prev) {
          const float4* u4 = reinterpret_cast<const float4*>(U) + row * gc4;
          float full = 0.f, rest = 0.f;
          for (int c = 0; c < gc4; ++c) {
            const float4 q = u4[c];
            const int cell = 4 * c - cc0;                                 // component k is cell (cell + k) of the touched range
            full = fmaxf(fmaxf(full, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
            rest = fmaxf(rest, ((unsigned)(cell + 0) < (unsigned)ncc) ? 0.f : q.x);
            rest = fmaxf(rest, ((unsigned)(cell + 1) < (unsigned)ncc) ? 0.f : q.y);
            rest = fmaxf(rest, ((unsigned)(cell + 2) < (unsigned)ncc) ? 0.f : q.z);
            rest = fmaxf(rest, ((unsigned)(cell + 3) < (unsigned)ncc) ? 0.f : q.w);
          }
          Rm[row] = full;
          m = fmaxf(m, in_cur ? rest : full);
        } else m = fmaxf(m, Rm[row]);
      }
      m = warp_max(m);
      if (lane == 0) redU[warp] = m;
    };
    // ===== phase A =====
    float pm2 = 0.f;                                                // max|P|^2 for this update's object step
    if (tid < nA) {
      // pending pupil update P += Q / max|objF| (fpmMain.cpp:470-475) for the elements this thread reads, Phi = O*P,
      // inverse column stage A.  Every shared-memory load of the item is issued before the first dependent operation.
      const int i0 = tq, jc = tr;
      const int j = (p.xlo + jc) & (N - 1);
      const int ob = (i0 - p.ylo) * OCP + jc;
      float2 Ov[NIN], Qv[NIN], Pv[NIN];
      int pi[NIN];
      bool in[NIN];
      static_for<0, NIN>([&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int m = m_of(k);
        constexpr int off = (R2 * m < H) ? R2 * m : R2 * m - N;            // iw = i0 + off
        const int iw = i0 + off;
        in[k] = (iw >= p.ylo) && (iw <= p.yhi);
        pi[k] = in[k] ? ob + off * OCP : 0;
        Ov[k] = Oc[pi[k]]; Qv[k] = Qc[pi[k]]; Pv[k] = Pc[pi[k]];
      });
      float2 w[NIN], v[R1];
      static_for<0, NIN>([&](auto K) {
        constexpr int k = decltype(K)::value;
        const float2 Pn = cfma(inv_objf_max, Qv[k], Pv[k]);
        if (in[k]) Pc[pi[k]] = Pn;
        pm2 = fmaxf(pm2, in[k] ? fmaf(Pn.x, Pn.x, Pn.y * Pn.y) : 0.f);
        const float2 phi = cmul(Ov[k], Pn);
        w[k] = in[k] ? phi : make_float2(0.f, 0.f);
      });
      if constexpr (SIX) fft16_in6<true>(w, v);
      else {
#pragma unroll
        for (int k = 0; k < R1; ++k) v[k] = w[k];
        fftR<R1, true>(v);
      }
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1)
        fld[(i0 + R2 * k1) * PITCH + j] = twmul4(v[k1], twA[k1 * R2 + i0]);
    }
    const int rbt = side ? side_tid : rb_tid, rbn = side ? side_n : rb_n;
    if (rbt >= 0 && u > 0) {
      // the cells the previous rectangle touched take their rebuilt maxima (W holds every pixel of those cells)
      const int pwsh = 32 - __clz((pr_ncc << 4) - 1);
      const unsigned mul = (65536u + (unsigned)pr_ncc - 1u) / (unsigned)pr_ncc;
      for (int t = rbt; t < NR * pr_ncc; t += rbn) {
        const int a = (int)(((unsigned)t * mul) >> 16), b = t - a * pr_ncc;
        const float4* w4 = reinterpret_cast<const float4*>(W + (a << pwsh) + (b << 4));
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float4 v = w4[q]; m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
        U[(pr_r0 + a) * gc + pr_cc0 + b] = m;
      }
    }
    if (side && warp >= sw0) {                                      // whole warps: U is complete among them, then the scan
      asm volatile("bar.sync 8, %0;" ::"r"(side_n) : "memory");
      untouched_max(side_tid, side_n);
    }
    pm2 = warp_max(pm2);                                            // (the warp straddling nA has lanes of both kinds)
    if (lane == 0) redP[warp] = pm2;
    __syncthreads();
    FPM_TICK(1);
    // ===== phase B =====
    if (!side) untouched_max(tid, NT);
    FPM_TICK(10);
    // window of update u+1 -> the buffer update u-1 released.  Its TMA store was issued at the start of phase A and
    // must be complete (the windows overlap in the spectrum): requested after S4 the wait is free, and the load has
    // S5 and S6 to land.  (fence.proxy.async without a state space compiles to MEMBAR.ALL.GPU, ~1k cycles on the issuing
    // warp's row block.  None is needed: the buffer was last written by the generic proxy before the
    // fence.proxy.async.shared::cta that precedes the TMA store, and last read by that store.)
    auto next_window = [&]() {
      if (tid == FPM_TMA_TID && u > 0) {
        tma_store_wait_all();
        mbar_expect_tx(wbar, win_bytes);
        tma_load_window(Ocn, &p.tmap, 2 * ((cr_b.x + H + p.xlo) & ~1), cr_b.y + H + p.ylo, tile, wbar);
      }
    };
    // ---- S2: inverse column stage B ----
    col_items_B([&](int k1, int jc) {
      const int js = (p.xlo + jc) & (N - 1);
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = fld[(R2 * k1 + a) * PITCH + js];
      fftR<R2, true>(v);
#pragma unroll
      for (int a = 0; a < R2; ++a) fld[(R2 * k1 + a) * PITCH + js] = v[a];
    });
    row_block_sync();
    FPM_TICK(2);
    // ---- S3: inverse row stage A (lanes run over rows; columns outside the bbox are zero, not read) ----
    {
      constexpr int NI = R2 / WPB;
      auto load_in = [&](int j0, float2 (&w)[NIN]) {
        const float2* rp = fld + rb_row * PITCH + j0;
        static_for<0, NIN>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int m = m_of(k);
          const int jw = (R2 * m < H) ? j0 + R2 * m : j0 + R2 * m - N;
          w[k] = (jw >= p.xlo && jw <= p.xhi) ? rp[R2 * m] : make_float2(0.f, 0.f);       // (j0 is warp-uniform)
        });
      };
      float2 wn[NIN];
      load_in(rb_sub, wn);
      static_for<0, NI>([&](auto Q) {
        constexpr int qi = decltype(Q)::value;
        const int j0 = rb_sub + qi * WPB;
        float2 w[NIN], v[R1];
#pragma unroll
        for (int k = 0; k < NIN; ++k) w[k] = wn[k];
        if constexpr (qi + 1 < NI) load_in(j0 + WPB, wn);
        if constexpr (SIX) fft16_in6<true>(w, v);
        else {
#pragma unroll
          for (int k = 0; k < R1; ++k) v[k] = w[k];
          fftR<R1, true>(v);
        }
        float2* rp = fld + rb_row * PITCH + j0;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) rp[R2 * k1] = twmul4(v[k1], twA[k1 * R2 + j0]);
      });
    }
    row_block_sync();
    FPM_TICK(3);
    // ---- S4: inverse row stage B + amplitude replacement (fpmMain.cpp:378-393) + forward row stage B' ----
    {
      constexpr int S4R = R1 / WPB;
      float4 ivall[S4R][CH];
#pragma unroll
      for (int rq = 0; rq < S4R; ++rq) {
        const int g = (rb_sub + rq * WPB) * N + rb_row;
        const float4* ip = reinterpret_cast<const float4*>(img) + (size_t)g * CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) ivall[rq][c] = __ldg(ip + c);
      }
      float2 vnext[R2];
      {
        const float2* rp0 = fld + rb_row * PITCH + R2 * rb_sub;
#pragma unroll
        for (int a = 0; a < R2; ++a) vnext[a] = rp0[a];
      }
#pragma unroll
      for (int rq = 0; rq < S4R; ++rq) {
        const int k1 = rb_sub + rq * WPB;
        float2* rp = fld + rb_row * PITCH + R2 * k1;
        float2 v[R2];
#pragma unroll
        for (int a = 0; a < R2; ++a) v[a] = vnext[a];
        if (rq + 1 < S4R) {
#pragma unroll
          for (int a = 0; a < R2; ++a) vnext[a] = rp[R2 * WPB + a];
        }
        fftR<R2, true>(v);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          const float4 q4 = ivall[rq][k2 >> 2];
          const int e = k2 & 3;
          const float inv_i = (e == 0) ? q4.x : (e == 1) ? q4.y : (e == 2) ? q4.z : q4.w;
          const float2 tt = cadd(v[k2], make_float2(epsr, epsi));
          const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * inv_i);
          v[k2] = cscale(v[k2], sc);
        }
        fftR<R2, false>(v);
#pragma unroll
        for (int q = 0; q < R2; ++q) rp[q] = twmul4(v[q], twB[q * R1 + k1]);
      }
    }
    next_window();
    row_block_sync();
    FPM_TICK(4);
    // ---- S5: forward row stage A' (only bbox columns are stored: nothing else is read afterwards) ----
    {
      constexpr int NI = R2 / WPB;
      float2 vn[R1];
      {
        const float2* rp0 = fld + rb_row * PITCH + rb_sub;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) vn[k1] = rp0[R2 * k1];
      }
      static_for<0, NI>([&](auto Q) {
        constexpr int qi = decltype(Q)::value;
        const int q = rb_sub + qi * WPB;
        float2* rp = fld + rb_row * PITCH + q;
        float2 v[R1];
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = vn[k1];
        if constexpr (qi + 1 < NI) {
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) vn[k1] = rp[WPB + R2 * k1];
        }
        fftR<R1, false>(v);
        static_for<0, NIN>([&](auto K) {                             // (the other outputs are dead code)
          constexpr int k = decltype(K)::value;
          constexpr int r = m_of(k);
          const int jw = (R2 * r < H) ? q + R2 * r : q + R2 * r - N;
          if (jw >= p.xlo && jw <= p.xhi) rp[R2 * r] = v[r];
        });
      });
    }
    row_block_sync();
    FPM_TICK(5);
    // ---- S6: forward column stage B' ----
    col_items_B([&](int k1, int jc) {
      const int js = (p.xlo + jc) & (N - 1);
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = fld[(R2 * k1 + a) * PITCH + js];
      fftR<R2, false>(v);
#pragma unroll
      for (int q = 0; q < R2; ++q) fld[(R2 * k1 + q) * PITCH + js] = twmul4(v[q], twB[q * R1 + k1]);
    });
    __syncthreads();
    FPM_TICK(6);
    // ===== phase C: forward column stage A' of the bbox outputs + object update (fpmMain.cpp:406-447) =====
    const int r0n = cr_b.y + H + p.ylo, c0n = (cr_b.x + H + p.xlo) & ~1;   // origin of the next window's TMA box
    FPM_TICK(11);
    if (u > 0) { mbar_wait(wbar, wphase); wphase ^= 1; }           // the next window has landed (it is patched below)
    FPM_TICK(12);
    float mx = 0.f;                                                 // max |.|^2 over this thread's rectangle and edge pixels
    {
      // edge pixels of the touched cells outside the rectangle (unchanged by this update): per rectangle row the left
      // part of the first cell (lanes 0..15) and the right part of the last cell (lanes 16..31).  They read L2
      // (ld.cg): earlier windows were written back by TMA stores.  Issued here, consumed after the element loop.
      constexpr int EPRE = (SIX ? 48 : 64) / NW;                    // rows warp, warp + NW, ... (NR <= 47 or 64)
      const int ecw = (lane < 16) ? lane : wcols - 32 + lane;
      const bool evalid = (lane < 16) ? (wc0 + ecw < c0) : (wc0 + ecw > c1);
      float2 eraw[EPRE];
#pragma unroll
      for (int k = 0; k < EPRE; ++k) {
        const int it = warp + k * NW;
        eraw[k] = make_float2(0.f, 0.f);
        if (it < NR && evalid) eraw[k] = __ldcg(objFc + (size_t)(r0 + it) * L + wc0 + ecw);
      }
      // (1 / max|P| is known since the barrier after phase A; computed there it costs a register through phase B:
      //  neutral at N = 128, -3 % at N = 64)
      const float inv_pmax = rsqrt_fast(warp_max(redP[lane]));
      FPM_TICK(13);

      // one element: (ir, jc) of the bbox; on == false lanes compute on element 0 and store nothing
      auto element = [&](int ir, int jc, bool on) {
        const int iw = p.ylo + ir, i = iw & (N - 1), js = (p.xlo + jc) & (N - 1);
        const int q = i & (R2 - 1), r = i / R2;
        // Phi'(i, j) = sum_k1 fld[R2*k1 + q][j] * W_R1^(r*k1);  W_R1^(r*(k + R1/2)) = (-1)^r W_R1^(r*k)
        const float2* fp = fld + q * PITCH + js;
        const float4* tp = TW + r * TP;
        const float sg = (r & 1) ? -1.f : 1.f;
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < TK; k += 2) {
          const float2 y0 = cfma(sg, fp[(R2 * (k + TK)) * PITCH], fp[(R2 * k) * PITCH]);
          const float2 y1 = cfma(sg, fp[(R2 * (k + TK + 1)) * PITCH], fp[(R2 * (k + 1)) * PITCH]);
          acc0 = cfma4(acc0, y0, tp[k]);
          acc1 = cfma4(acc1, y1, tp[k + 1]);
        }
        const float2 phin = cadd(acc0, acc1);
        const int t = ir * OCP + jc;
        const float sup = Sc[t];
        const float2 O = Oc[t];
        const float2 Pv = Pc[t];
        const float2 d = csub(phin, cmul(O, Pv));                      // dPhi = Phi' - Phi
        // dO = d * |P| conj(P) / (max|P| * ((|P|^2 + delta2) + i*kappa*delta2))
        const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
        const float2 num = cmulc(d, Pv);
        const float A = pa2 + p.delta2;
        const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
        const float2 On = make_float2(O.x + (num.x * A + num.y * kd2) * sc, O.y + (num.y * A - num.x * kd2) * sc);
        const float a2n = fmaf(On.x, On.x, On.y * On.y);
        // Q = d * |O| conj(O) / ((|O|^2 + delta1) + i*kappa*delta1) * support   (fpmMain.cpp:459-472, O before the update)
        const float oa2 = fmaf(O.x, O.x, O.y * O.y);
        const float2 numq = cmulc(d, O);
        const float A1 = oa2 + p.delta1;
        const float sq = sqrt_fast(oa2) * sup * rcp_fast(fmaf(A1, A1, kd1 * kd1));
        // stores of switched-off lanes (and forwards outside the overlap) go to a sink: no branches, so that the two
        // elements a thread has in flight interleave
        const int rn = r0 + ir - r0n, cn = c0 + jc - c0n;
        const bool fw = on && (unsigned)rn < (unsigned)NR && (unsigned)cn < (unsigned)OCP;
        *(on ? Oc + t : sink2) = On;                                      // the box goes back to the spectrum with one TMA store
        *(fw ? Ocn + rn * OCP + cn : sink2) = On;                      // forward into the next window where the two overlap
        *(on ? W + (ir << wsh) + (c0 + jc - wc0) : sink1) = a2n;
        mx = fmaxf(mx, on ? a2n : 0.f);
        *(on ? Qc + t : sink2) = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
      };
      auto decode = [&](int it, int& ir, int& jc) -> bool {            // (branch-free: the two elements of a pair interleave)
        const bool full = it < nC_full;
        const int e = ((it - nC_full) << 5) + lane;
        const bool on = full || ((it < nC_items) && (e < NR * cb_nl));
        const int er = (int)(((unsigned)e * nl_mul) >> 16);
        ir = full ? (it >> fsh) : (on ? er : 0);
        jc = full ? ((it & fsh) << 5) + lane : (on ? (cb_nfull << 5) + e - er * cb_nl : 0);
        return on;
      };
      for (int it = warp; it < nC_items; it += 2 * NW) {
        int ira, jca, irb, jcb;
        const bool ona = decode(it, ira, jca);
        if (it + NW < nC_items) {                                   // two elements in flight
          const bool onb = decode(it + NW, irb, jcb);
          element(ira, jca, ona);
          element(irb, jcb, onb);
        } else element(ira, jca, ona);
      }
      FPM_TICK(15);
#pragma unroll
      for (int k = 0; k < EPRE; ++k) {
        const int it = warp + k * NW;
        asm volatile("" : "+f"(eraw[k].x), "+f"(eraw[k].y));       // consumed here, not where they were issued
        if (it < NR && evalid) {
          const float a2 = fmaf(eraw[k].x, eraw[k].x, eraw[k].y * eraw[k].y);
          W[(it << wsh) + ecw] = a2;
          mx = fmaxf(mx, a2);
        }
      }
    }
    mx = warp_max(mx);
    FPM_TICK(14);
    if (lane == 0) redC[warp] = mx;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // Oc writes -> visible to the TMA store
    __syncthreads();
    FPM_TICK(8);
    if (tid == FPM_TMA_TID) tma_store_window(Ocur, &p.tmap, 2 * (c0 & ~1), r0, tile);   // updated window -> objFc
    pr_r0 = r0; pr_cc0 = cc0; pr_ncc = ncc;
    cr_a = cr_b; cr_b = cr_c;
  }

#ifdef FPM_STAGE_TIMING
  if (tid == FPM_TICK_TID && blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p.stage_clk[k] += tacc_[k];
  }
#endif
  if (tid == FPM_TMA_TID) tma_store_wait_all();
  {                                                                  // the last update's pupil increment is still pending
    const float m = warp_max(fmaxf(redC[lane], redU[lane]));
    const float inv_objf_max = (p.n_updates > 0) ? rsqrt_fast(m) : 0.f;
    for (int t = tid; t < NR * OCP; t += NT) {
      const int ir = t / OCP, jc = t - ir * OCP;
      const float2 Pn = cfma(inv_objf_max, Qc[t], Pc[t]);
      if (jc < NC) Pg[((p.ylo + ir) & (N - 1)) * N + ((p.xlo + jc) & (N - 1))] = Pn;
    }
  }
  if (p.ucache) {
    // leave the cell maxima for the tile's next pass: the cells of the last rectangle are still to be rebuilt from W
    if (p.n_updates > 0) {
      const int pwsh = 32 - __clz((pr_ncc << 4) - 1);
      for (int t = tid; t < NR * pr_ncc; t += NT) {
        const int a = t / pr_ncc, b = t - a * pr_ncc;
        const float4* w4 = reinterpret_cast<const float4*>(W + (a << pwsh) + (b << 4));
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float4 v = w4[q]; m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
        U[(pr_r0 + a) * gc + pr_cc0 + b] = m;
      }
    }
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(p.ucache + (size_t)tile * L * gc);
    for (int t = tid; t < L * gc4; t += NT) dst[t] = reinterpret_cast<const float4*>(U)[t];
  }
}

}  // namespace fpm
