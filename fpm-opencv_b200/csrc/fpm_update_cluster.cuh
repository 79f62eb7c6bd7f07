// fpm_update_cluster.cuh -- one tile's sub-aperture update spread over a thread-block cluster (sm_100a).
//
// Same arithmetic and stage structure as fpm_update.cuh (fpmMain.cpp:350-475 per update); here C CTAs of one
// cluster share a tile, so a 256x256 field (512 KB) stays on chip and a single tile is no longer bound to one SM.
//
//   column passes (S1,S2,S6,S7), object / pupil update (C2): CTA k owns bbox columns [k*CPC, (k+1)*CPC) -- its
//       slice of P, Q = pending pupil increment, the support and a column slab  cslab[column][row position]
//   row passes (S3,S4,S5 incl. the amplitude replacement):   CTA k owns N/C scrambled row positions,
//       row slab  rslab[row][column]
//   The two transposes between them are written straight into the owner's shared memory by the producing butterfly
//   (st.async + the owner's mbarrier counts the bytes: no cluster-wide barrier, no fence on the data path).
//   max|objF| (fpmMain.cpp:460,467): the grid of cell maxima is distributed by cell row (cell row a lives in CTA
//   a mod C).  Per update ONE exchange: every rank sends max(its partial maxima of the touched cells = its slice's new
//   values + its share of the edge pixels, the maximum of its share of the untouched cells) to every rank.  The
//   untouched-cell maximum and the edge pixels do not depend on the update and are taken beside the column stages; the
//   merges of the partial maxima into the owners' grids (DSMEM atomics), the one release fence and the ubar arrivals
//   follow while the exchange is in flight, and are acquired by the next update's side work.
//   The spectrum window: narrow boxes on 128 x 128 tiles (SIX instances) keep each CTA's column slice of the window in
//   shared memory, double-buffered; C2 forwards every new value into the shared memory of the CTA that owns its column
//   in the NEXT window (st.async + that CTA's wbar), the part of the next slice outside this update's rectangle is read
//   from the spectrum by threads without a column item.  Other boxes read their slice from the spectrum (L2) in S1 and
//   C2.  The spectrum itself is written with plain coalesced stores in C2 (each CTA touches only its own columns within
//   an update; the release / acquire pair of ubar orders them between updates).
//   Threads without a column item in S1 / S2 ("helpers", the same count in every CTA) run beside the column stages.  The
//   first helper warp (duty warp) arms the mbarriers, prefetches the next 1/I rows and issues the release fence; the
//   others (side group, own named barrier) clear + scan the untouched cells, fetch the untouched part of the next window
//   slice and take the edge-pixel maxima (narrow boxes: beside S6 / S7).  The column threads meet at a named barrier
//   between S1 and S2 (narrow boxes: S6 and S7 too) instead of a block barrier.
//   Barrier rule (a hang was found the hard way): a barrier whose phase can complete the moment it is armed (wbar with
//   nothing to forward) is armed only after a block barrier that every waiter of the previous phase has passed.
//   [r2] measured, one tile: 128 x 128 with a 35 x 35 box on 4 CTAs 9.3 -> 6.5 us per update (S1 4.0k -> 1.7k cycles,
//   the work after C2 4.8k -> 2.2k of which 1.2k overlap the exchange), on 2 CTAs 11.6 -> 9.0; 85 x 85 box 11.3 -> 9.8;
//   256 x 256 on 8 CTAs 18.7 -> 17.5 (cfg5b), 19.2 -> 18.2 (cfg3).  The window slice on chip for the 85 x 85 box:
//   12.2 us (1870 forwards per CTA and update cost more than the loads they save), not used.
//   Also measured and dropped: the slice's partial cell maxima taken inside C2 (match.any + redux + one shared atomic per
//   distinct cell of the warp instead of the W pass after C2): C2 1.55k -> 2.9k cycles at N = 128, 5.7k -> 9.6k at 256.
#pragma once
#include <cooperative_groups.h>
#include "fpm_update.cuh"

#ifndef FPM_TICK_TID
#define FPM_TICK_TID 0        // the thread whose stage clocks the timing build reports
#endif

namespace fpm {
namespace cg = cooperative_groups;

// ---- DSMEM primitives: remote stores that signal the destination CTA's mbarrier (no cluster-wide fence) ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f2(uint32_t raddr, float2 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "f"(v.x), "f"(v.y), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(raddr), "f"(v), "r"(rbar) : "memory");
}
#ifdef FPM_CL_DEBUG
// developer build: a wait that gives up after 2^25 polls and says which barrier of which update never completed
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int id = -1, int u = -1) {
  __syncwarp();
  uint32_t bad;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .pred q;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "mov.u32 %0, 0;\n\t"
      "WAITD_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra DONED_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.gt.u32 q, n, 0x2000000;\n\t"
      "@!q bra WAITD_%=;\n\t"
      "mov.u32 %0, 1;\n\t"
      "DONED_%=:\n\t}" : "=r"(bad) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (bad) {
    if ((threadIdx.x & 31) == 0) printf("mbarrier %d never completed: update %d block %d warp %d parity %u\n", id, u, blockIdx.x, threadIdx.x >> 5, parity);
    __trap();
  }
}
#else
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int = -1, int = -1) {   // acquire at cluster scope
  // Whole warps wait.  The warp reconverges first: lanes that skipped the preceding column stage must not spin here
  // while their siblings still have the remote stores to issue that this (or a peer's) barrier is waiting for.
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
// relaxed: the caller issues ONE fence.acq_rel.cluster before a batch of these (a .release arrive costs a
// MEMBAR.ALL.GPU each)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t rbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}

// Shared-memory carve-up, identical in every CTA of the cluster (DSMEM addresses are rank-mapped offsets).
template <int N, int C> struct ClusterLayout {
  size_t rslab, cslab, twA, twB, Pc, Qc, Sc, Wb, U, Tm, red, pmx, omx, bars, total;
  int gro, tmr, tmc;
  // ws (narrow boxes on 128 x 128 tiles, the SIX instances of the kernel): this CTA's column slice of the window stays in
  // shared memory, double-buffered (see "window slice" below).  Wide boxes read the slice from the spectrum (L2) in S1
  // and C2: at N = 256 the slabs leave no room for it, and at N = 128 forwarding a wide rectangle element by element
  // costs more than the loads it saves (85 x 85 box on four CTAs: 12.2 us per update with the slice on chip, 11.3 without).
  __host__ __device__ ClusterLayout(int NR, int NC, int CPC, int L, int cs, bool ws) {
    using S = Shape<N>;
    size_t o = 0;
    rslab = o; o += (sizeof(float2) * (N / C) * S::PITCH + 15) / 16 * 16;
    cslab = o; o += (sizeof(float2) * CPC * (N + 1) + 15) / 16 * 16;
    twA = o; o += sizeof(float2) * N;
    twB = o; o += sizeof(float2) * N;
    Pc = o; o += (sizeof(float2) * NR * CPC + 15) / 16 * 16;
    Qc = o; o += (sizeof(float2) * NR * CPC + 15) / 16 * 16;
    Sc = o; o += (sizeof(float) * NR * CPC + 15) / 16 * 16;
    Wb = o; o += ws ? 2 * ((sizeof(float2) * NR * CPC + 15) / 16 * 16) : 0;
    gro = ((L >> cs) + C - 1) / C;                      // cell rows per CTA
    U = o; o += (sizeof(float) * gro * (L >> 4) + 15) / 16 * 16;
    tmr = (NR >> cs) + 2; tmc = (NC >> 4) + 2;
    Tm = o; o += (sizeof(unsigned) * tmr * tmc + 15) / 16 * 16;
    red = o; o += sizeof(float) * 64;
    pmx = o; o += sizeof(float) * 16;
    omx = o; o += sizeof(float) * 16;
    bars = o; o += sizeof(uint64_t) * 6;
    total = o;
  }
};

// SIX (N = 128 only): the box lies within +-(3 * R2 - 1), so the radix-16 stage-A butterflies see 6 of their 16 samples
// (and the stage-A' butterflies keep 6 outputs), as in fpm_update_phased_kernel.
template <int N, int C, int NT, bool SIX = false>
__global__ void __launch_bounds__(NT, 1) fpm_update_cluster_kernel(const __grid_constant__ UpdateParams p) {
  using S = Shape<N>;
  constexpr int R1 = S::R1, R2 = S::R2, PITCH = S::PITCH, CH = S::CH;
  constexpr int H = N / 2, NW = NT / 32;
  static_assert(!SIX || R1 == 16, "six-sample butterflies are radix 16");
  constexpr int NIN = SIX ? 6 : R1;                         // samples a stage-A butterfly reads
  auto m_of = [](int k) constexpr { return SIX ? ((k < 3) ? k : R1 - 6 + k) : k; };
  constexpr int RPC = N / C;                // scrambled row positions per CTA
  constexpr int PR = N + 1;                 // cslab pitch (float2 per column): odd, lanes = columns hit distinct banks
  static_assert(RPC % R2 == 0 && RPC % 32 == 0, "a stage-B work item must not straddle two row owners");
  static_assert(C <= 16, "per-rank slots");
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = p.tile0 + blockIdx.x / C;
  const int L = p.L;
  const int NR = p.yhi - p.ylo + 1, NC = p.xhi - p.xlo + 1;
  const int CPC = p.ocp;                                   // bbox columns per CTA (<= 32: one lane per column)
  const int cpc_inv = (65536 + CPC - 1) / CPC;
  const int jc0 = rank * CPC;
  const int ncl = max(0, min(CPC, NC - jc0));              // columns this CTA really has
  const int ncl_inv = ncl ? (65536 + ncl - 1) / ncl : 0;
  const int gc = L >> 4, gr = L >> p.cs;
  // column work items (group, local column) of this thread: R1 * ncl <= 16 * 32 = NT, so every column stage is a single
  // round and the item index is computed once (a runtime integer division per stage and update sits on the critical path)
  static_assert(Shape<N>::R1 * 32 <= NT, "one column work item per thread");
  const int tqc = ncl ? tid / ncl : 0, trc = ncl ? tid - tqc * ncl : 0;
  const int qNTc = ncl ? NT / ncl : 0, rNTc = ncl ? NT % ncl : 0;               // t += NT without dividing

  constexpr bool WS = SIX;
  const ClusterLayout<N, C> lay(NR, NC, CPC, L, p.cs, WS);
  float2* rslab = reinterpret_cast<float2*>(smem_raw + lay.rslab);
  float2* cslab = reinterpret_cast<float2*>(smem_raw + lay.cslab);
  float2* twA = reinterpret_cast<float2*>(smem_raw + lay.twA);
  float2* twB = reinterpret_cast<float2*>(smem_raw + lay.twB);
  float2* Pc = reinterpret_cast<float2*>(smem_raw + lay.Pc);
  float2* Qc = reinterpret_cast<float2*>(smem_raw + lay.Qc);
  float* Sc = reinterpret_cast<float*>(smem_raw + lay.Sc);
  float* U = reinterpret_cast<float*>(smem_raw + lay.U);          // cell rows a = rank, rank+C, ...: U[(a/C)*gc + b]
  unsigned* Tm = reinterpret_cast<unsigned*>(smem_raw + lay.Tm);  // partial maxima of the touched cells (bit patterns)
  float* red = reinterpret_cast<float*>(smem_raw + lay.red);
  float* pmx = reinterpret_cast<float*>(smem_raw + lay.pmx);      // [C] max|P|^2 of every rank's slice
  float* omx = reinterpret_cast<float*>(smem_raw + lay.omx);      // [C] max of every rank's share of U
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
  uint64_t* rbar = bars + 0;   // row slab (+ max|P|^2 slots) complete: counts the bytes of the remote st.async
  uint64_t* cbar = bars + 1;   // column slab complete
  uint64_t* ubar = bars + 2;   // every rank has merged its cell maxima (C arrivals); also orders the spectrum stores
  uint64_t* obar = bars + 3;   // every rank's grid maximum has arrived
  uint64_t* wbar = bars + 4;   // window slice of the next update complete (the forwarded part: bytes of remote st.async)
  const size_t wb_stride = (sizeof(float2) * NR * CPC + 15) / 16 * 16;
  float2* const Wb0 = reinterpret_cast<float2*>(smem_raw + lay.Wb);               // window slice [NR][CPC] of even updates
  float2* const Wb1 = reinterpret_cast<float2*>(smem_raw + lay.Wb + (WS ? wb_stride : 0));   // ... of odd updates
  float* W = reinterpret_cast<float*>(rslab);                     // |O_new|^2 on the slice; rslab is idle during C2/D
  const int tmc = lay.tmc;

  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* Pg = p.pupil + (size_t)tile * N * N;
  const float* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * N * N;

  // ---- prologue ----
  for (int t = tid; t < N; t += NT) {
    const int b = t / R2, a = t % R2;
    twA[t] = p.tw[a * b];
    const int a2 = t / R1, b2 = t % R1;
    twB[t] = p.tw[a2 * b2];
  }
  for (int t = tid; t < NR * CPC; t += NT) {
    const int ir = t / CPC, jcl = t - ir * CPC;
    float2 pv = make_float2(0.f, 0.f);
    float sv = 0.f;
    if (jcl < ncl) {
      const int iw = p.ylo + ir, jw = p.xlo + jc0 + jcl;
      const int gi = (iw & (N - 1)) * N + (jw & (N - 1));
      pv = Pg[gi]; sv = p.support[gi];
    }
    Pc[t] = pv; Sc[t] = sv; Qc[t] = make_float2(0.f, 0.f);
  }
  {
    const int nown = (gr - rank + C - 1) / C;              // cell rows owned by this CTA
    for (int it = warp; it < nown * (L >> 5); it += NW) {
      const int la = it / (L >> 5), seg = it % (L >> 5);
      const int cellrow = rank + la * C;
      const float2* src = objFc + (size_t)(cellrow << p.cs) * L + (seg << 5) + lane;
      float cm = 0.f;
      for (int rr = 0; rr < (1 << p.cs); ++rr) {
        const float2 o = __ldcg(src + (size_t)rr * L);
        cm = fmaxf(cm, fmaf(o.x, o.x, o.y * o.y));
      }
      cm = half_warp_max(cm, lane);
      if ((lane & 15) == 0) U[la * gc + 2 * seg + (lane >> 4)] = cm;
    }
  }
  for (int t = tid; t < lay.tmr * lay.tmc; t += NT) Tm[t] = 0u;
  if (tid < 64) red[tid] = 0.f;                            // (warps that only do side work never write their max|P|^2 slot)
  if (tid == 0) {
    mbar_init(rbar, 1); mbar_init(cbar, 1); mbar_init(ubar, C); mbar_init(obar, 1); mbar_init(wbar, 1);
  }
  const uint32_t rbar_a = smem_u32(rbar), cbar_a = smem_u32(cbar), ubar_a = smem_u32(ubar), obar_a = smem_u32(obar);
  const uint32_t wbar_a = smem_u32(wbar);
  const uint32_t rslab_a = smem_u32(rslab), cslab_a = smem_u32(cslab), pmx_a = smem_u32(pmx), omx_a = smem_u32(omx);
  const uint32_t rbar_bytes = (uint32_t)(sizeof(float2) * RPC * NC + sizeof(float) * C);
  const uint32_t cbar_bytes = (uint32_t)(sizeof(float2) * N * ncl);
  cluster.sync();                                          // every CTA is resident and initialised before any remote write

  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  const float epsr = p.eps * (float)(N * N), epsi = p.kappa * epsr;   // the 1/N^2 of ifft2 is never applied (see fpm_update.cuh)
  float inv_objf_max = 0.f;
#ifdef FPM_STAGE_TIMING
  long long tacc_[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) tacc_[k] = 0;
  long long tprev_ = clock64();
#endif

  // Window slice (WS): the columns of the window this CTA owns live in shared memory, Wb[u & 1] for update u.  The first
  // one is read from the spectrum here.  During update u the slice of update u + 1 is assembled in the other buffer: the
  // elements outside this update's rectangle are read from the spectrum by threads without column work (they were
  // written at least one update ago), the elements inside it are forwarded by C2 -- each new value goes straight into
  // the shared memory of the CTA that owns its column in the next window (st.async + that CTA's wbar).  S1 and C2 read
  // shared memory only; the spectrum stores of C2 stay (the spectrum is the result), off the critical path.
  short2 cr_next = p.crop[p.slot_begin % p.n_leds];
  short2 cr_next2 = p.crop[(p.slot_begin + 1) % p.n_leds];
  if constexpr (WS) {
    const float2* w0 = objFc + (size_t)(cr_next.y + H + p.ylo) * L + cr_next.x + H + p.xlo + jc0;
    for (int t = tid; t < NR * CPC; t += NT) {
      const int ir = t / CPC, jcl = t - ir * CPC;
      Wb0[t] = (jcl < ncl) ? __ldcg(w0 + (size_t)ir * L + jcl) : make_float2(0.f, 0.f);
    }
    __syncthreads();
  }
  int slot_it = p.slot_begin % p.n_leds;
  for (int u = 0; u < p.n_updates; ++u) {
    const int slot = slot_it;                               // (p.slot_begin + u) % p.n_leds without the division
    slot_it = (slot + 1 == p.n_leds) ? 0 : slot + 1;
    const short2 cr = cr_next, crn = cr_next2;
    cr_next = cr_next2;
    {
      const int nslot = (slot + 1 == p.n_leds) ? 0 : slot + 1;
      const int nslot2 = (nslot + 1 == p.n_leds) ? 0 : nslot + 1;
      cr_next2 = p.crop[nslot2];                            // in flight during this update
    }
    const int xs = cr.x, ys = cr.y;
    const float* __restrict__ img = stack + (size_t)slot * N * N;
    const int r0 = ys + H + p.ylo, r1 = ys + H + p.yhi, c0 = xs + H + p.xlo, c1 = xs + H + p.xhi;   // rectangle (inclusive)
    const int cr0 = r0 >> p.cs, ncr = (r1 >> p.cs) - cr0 + 1, cc0 = c0 >> 4, ncc = (c1 >> 4) - cc0 + 1;
    const float2* wrow = objFc + (size_t)r0 * L + c0 + jc0;     // (ir, jcl) -> wrow[ir*L + jcl]
#ifdef FPM_STAGE_TIMING
    asm volatile("" ::"r"(xs));
    FPM_TICK(11);
#endif

    const uint32_t ph = (uint32_t)(u & 1);
    // next update's rectangle, and the part of this CTA's next slice that lies inside this update's rectangle (forwarded)
    const bool has_next = WS && (u + 1 < p.n_updates);
    const int r0n = crn.y + H + p.ylo, c0n = crn.x + H + p.xlo;
    const int ovr = max(0, min(r1, r0n + NR - 1) - max(r0, r0n) + 1);
    const int ovc = (ncl > 0) ? max(0, min(c1, c0n + jc0 + ncl - 1) - max(c0, c0n + jc0) + 1) : 0;
    float2* const Wcur = (u & 1) ? Wb1 : Wb0;
    float2* const Wnxt = (u & 1) ? Wb0 : Wb1;
    const uint32_t wnxt_a = smem_u32(Wnxt);
    // Threads tid >= hlp0 have no column item in S1 / S2 (the same count in every CTA of the cluster): "helpers".  Their
    // first warp also arms the barriers and issues the L2 prefetch, so that the S1 warps start on the butterflies at once.
    const int hlp0 = min(NT, (R1 * CPC + 31) & ~31), hn = NT - hlp0;
    const bool helpers = hn >= 64;
    const int duty0 = helpers ? hlp0 : 0;
    // Narrow boxes: S1 + S2 are too short for all of the side work (the row stage S3 waited ~2k cycles for the helpers);
    // the edge-pixel maxima are taken while the forward column stages S6 / S7 run instead.  Wide boxes and 256 x 256
    // tiles: everything early (85 x 85 box on four CTAs: 9.8 us per update early, 10.6 late).
    constexpr bool EDGE_LATE = SIX;
    if constexpr (WS) {
      // this update's slice is complete once the forwarded values of the previous C2 have landed (phase u - 1 of wbar);
      // the warps of S1 wait (whole warps)
      if (u > 0 && warp * 32 < R2 * ncl) mbar_wait_cluster(wbar, (uint32_t)((u - 1) & 1), 4, u);
    }
    auto duties = [&]() {
      if (tid == duty0) {   // arm this update's transfers (bytes may already be arriving: the counts are signed)
        mbar_expect_tx(rbar, rbar_bytes);
        mbar_expect_tx(cbar, cbar_bytes);
        mbar_expect_tx(obar, (uint32_t)(sizeof(float) * C));
        // (wbar is armed after S3: with nothing to forward its phase completes at once, and a warp that has not polled the
        //  previous phase yet would then see that parity as the current, incomplete phase and wait for ever)
      }
      if ((unsigned)(tid - duty0) < (unsigned)R1) {   // next LED's 1/I rows of this CTA towards L2: R1 chunks of RPC*R2 floats
        const int nslot = (slot + 1 == p.n_leds) ? 0 : slot + 1;
        const float* nx = stack + (size_t)nslot * N * N + ((size_t)(tid - duty0) * N + rank * RPC) * R2;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)(RPC * R2 * 4)) : "memory");
      }
    };
    // The duty warp = the first helper warp.  It also issues the release fence at the end of the update (~1.3k cycles) and
    // comes out of it late, so it stays out of the other helpers' barrier: they are the "side group".
    const int sg0 = hlp0 + 32, sgn = hn - 32;
    if (!helpers || (tid >= hlp0 && tid < sg0)) duties();
    // (cell columns padded to a power of two: shifts instead of a runtime integer division on the per-update path)
    const int csh = 32 - __clz(ncc - 1);
    // max|objF| bookkeeping that does not depend on this update (side work, see below): once every rank's merges of the
    // previous update have landed (ubar), the owners clear the cells this rectangle touches -- they are rebuilt from
    // scratch, the partial maxima arrive after C2 -- and take the maximum of their share of the grid, i.e. of the cells
    // this update leaves alone.
    auto untouched_max = [&](int t0, int tn, bool own_barrier) {
      for (int t = t0; t < (ncr << csh); t += tn) {
        const int ta = t >> csh, tb = t & ((1 << csh) - 1);
        const int a = cr0 + ta;
        if (tb < ncc && a % C == rank) U[(a / C) * gc + cc0 + tb] = 0.f;
      }
      if (own_barrier) asm volatile("bar.sync 2, %0;" ::"r"(tn) : "memory");   // the helpers among themselves
      else __syncthreads();
      const int nown = (gr - rank + C - 1) / C;
      const float4* U4 = reinterpret_cast<const float4*>(U);
      const int n4 = (nown * gc) >> 2;
      float m = 0.f;
      for (int t = t0; t < n4; t += tn) { const float4 q = U4[t]; m = fmaxf(fmaxf(m, fmaxf(q.x, q.y)), fmaxf(q.z, q.w)); }
      m = warp_max(m);
      if (lane == 0) red[warp] = m;
    };

    // Partial maxima of the touched cells, first part: the pixels of those cells outside the rectangle, which this update
    // does not change (shared out over the cluster): strips above / below (full width of the touched cells) and left /
    // right (rectangle rows).  They do not depend on this update, so they are taken while the column stage B runs
    // (threads t0, t0 + tn, ... of every CTA).
    auto edge_maxima = [&](int t0, int tn) {
      const int wc0 = cc0 << 4, wcols = ncc << 4, rt0 = cr0 << p.cs, rt1 = ((cr0 + ncr) << p.cs) - 1;
      const int wl = c0 - wc0, wrt = wc0 + wcols - 1 - c1;
      // Strip items.  When they fit one item per thread of the cluster even with padded widths (2^wsh2 for the
      // full-width strips, 16 for the side strips) the index is split with shifts; a runtime integer division on the
      // per-update path costs more than the idle lanes.  Large rectangles keep the dense enumeration.
      const int wsh2 = 32 - __clz(wcols - 1), wmask = (1 << wsh2) - 1;
      const int P1 = (r0 - rt0) << wsh2, P2 = P1 + ((rt1 - r1) << wsh2), P3 = P2 + (NR << 4), n_pad = P3 + (NR << 4);
      const bool padded = n_pad <= C * NT;
      const int n_top = (r0 - rt0) * wcols, n_bot = (rt1 - r1) * wcols, n_left = NR * wl, n_right = NR * wrt;
      const int n_all = padded ? n_pad : n_top + n_bot + n_left + n_right;
      constexpr int DU = 4;
      for (int base = rank * tn + t0; base < n_all; base += DU * C * tn) {
        float2 o[DU];
        int cell[DU];
#pragma unroll
        for (int k = 0; k < DU; ++k) {
          const int t = base + k * C * tn;
          cell[k] = -1;
          if (t < n_all) {
            int r, c;
            if (padded) {
              bool ok;
              if (t < P1) { const int cc = t & wmask; r = rt0 + (t >> wsh2); c = wc0 + cc; ok = cc < wcols; }
              else if (t < P2) { const int s = t - P1, cc = s & wmask; r = r1 + 1 + (s >> wsh2); c = wc0 + cc; ok = cc < wcols; }
              else if (t < P3) { const int s = t - P2, cc = s & 15; r = r0 + (s >> 4); c = wc0 + cc; ok = cc < wl; }
              else { const int s = t - P3, cc = s & 15; r = r0 + (s >> 4); c = c1 + 1 + cc; ok = cc < wrt; }
              if (!ok) continue;
            } else {
              if (t < n_top) { r = rt0 + t / wcols; c = wc0 + t % wcols; }
              else if (t < n_top + n_bot) { const int s = t - n_top; r = r1 + 1 + s / wcols; c = wc0 + s % wcols; }
              else if (t < n_top + n_bot + n_left) { const int s = t - n_top - n_bot; r = r0 + s / wl; c = wc0 + s % wl; }
              else { const int s = t - n_top - n_bot - n_left; r = r0 + s / wrt; c = c1 + 1 + s % wrt; }
            }
            o[k] = __ldcg(objFc + (size_t)r * L + c);
            cell[k] = ((r >> p.cs) - cr0) * tmc + ((c >> 4) - cc0);
          }
        }
#pragma unroll
        for (int k = 0; k < DU; ++k)
          if (cell[k] >= 0) atomicMax(&Tm[cell[k]], __float_as_uint(fmaf(o[k].x, o[k].x, o[k].y * o[k].y)));
      }
    };
    // the part of the next window's slice that this update does not touch: spectrum -> Wnxt.  One slice element per
    // thread and round (two rounds in flight); t / ncl by a 16-bit reciprocal, exact for t < 2048 (narrow boxes: t < 47 * 32).
    auto next_slice = [&](int t0, int tn) {
      const float2* wn = objFc + (size_t)r0n * L + c0n + jc0;
      const int dr = r0n - r0, dc = c0n + jc0 - c0;             // next-slice (rn, lc) is (rn + dr, lc + dc) of this rectangle
      const int n = NR * ncl;
      for (int t = t0; t < n; t += 2 * tn) {
        const int tb = t + tn;
        const int rna = (t * ncl_inv) >> 16, lca = t - rna * ncl;
        const int rnb = (tb * ncl_inv) >> 16, lcb = tb - rnb * ncl;
        const bool fa = !((unsigned)(rna + dr) < (unsigned)NR && (unsigned)(lca + dc) < (unsigned)NC);
        const bool fb = tb < n && !((unsigned)(rnb + dr) < (unsigned)NR && (unsigned)(lcb + dc) < (unsigned)NC);
        float2 va = make_float2(0.f, 0.f), vb = va;
        if (fa) va = __ldcg(wn + (size_t)rna * L + lca);
        if (fb) vb = __ldcg(wn + (size_t)rnb * L + lcb);
        if (fa) Wnxt[rna * CPC + lca] = va;
        if (fb) Wnxt[rnb * CPC + lcb] = vb;
      }
    };
    // Threads without a column item in S1 / S2 take this side work while the column stages run; the column threads
    // meet at a named barrier between S1 and S2.
    // Every rank's release of the previous update (its merges into the owners' grids and its spectrum stores) is
    // acquired by whoever reads them: the side work, and the window loads of S1 / C2 when the slice is not on chip.
    if (u > 0 && (!WS || !helpers || tid >= hlp0)) mbar_wait_cluster(ubar, (uint32_t)((u - 1) & 1), 2, u);
#ifdef FPM_TICK_SIDE          // (timing build reporting a helper thread: slots 14 / 2 / 3 = ubar wait / untouched maximum / duties)
    FPM_TICK(14);
#endif
    if (!helpers || tid >= sg0) {                            // (without helper threads: everybody, ahead of S1)
      const int t0 = helpers ? tid - sg0 : tid, tn = helpers ? sgn : NT;
      untouched_max(t0, tn, helpers);
#ifdef FPM_TICK_SIDE
      FPM_TICK(2);
#endif
      if (has_next) next_slice(t0, tn);
      if (!EDGE_LATE || !helpers) edge_maxima(t0, tn);
    }
    FPM_TICK(0);
    // ===== S1: pending pupil update (fpmMain.cpp:470-475), Phi = O*P, cols stage A (inverse) on this CTA's columns =====
    if (!helpers || tid < hlp0) {
      float pm2 = 0.f;
      if (tid < R2 * ncl) {                              // work items packed densely over the lanes
        const int i0 = tqc, jcl = trc;
        {
          float2 w[NIN], v[R1];
          // Branch-free: rows outside the bbox read a clamped (valid) address and are zeroed afterwards, so the
          // window loads issue back to back and the pupil work below is straight-line code.
#pragma unroll
          for (int k = 0; k < NIN; ++k) {
            const int i = i0 + R2 * m_of(k);
            const int iw = (i < H) ? i : i - N;
            const int irc = min(max(iw - p.ylo, 0), NR - 1);
            if constexpr (WS) w[k] = Wcur[irc * CPC + jcl];
            else w[k] = wrow[irc * L + jcl];
          }
#pragma unroll
          for (int k = 0; k < NIN; ++k) {
            const int i = i0 + R2 * m_of(k);
            const int iw = (i < H) ? i : i - N;
            const bool in = (iw >= p.ylo && iw <= p.yhi);
            const int e = min(max(iw - p.ylo, 0), NR - 1) * CPC + jcl;
            const float2 Q = Qc[e];
            float2 Pv = Pc[e];
            Pv.x = fmaf(Q.x, inv_objf_max, Pv.x);
            Pv.y = fmaf(Q.y, inv_objf_max, Pv.y);
            if (in) Pc[e] = Pv;                              // (a clamped row belongs to another work item: read only)
            const float2 phi = cmul(w[k], Pv);
            pm2 = in ? fmaxf(pm2, fmaf(Pv.x, Pv.x, Pv.y * Pv.y)) : pm2;
            w[k] = in ? phi : make_float2(0.f, 0.f);
          }
          if constexpr (SIX) fft16_in6<true>(w, v);
          else {
#pragma unroll
            for (int k = 0; k < R1; ++k) v[k] = w[k];
            fftR<R1, true>(v);
          }
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) cslab[jcl * PR + i0 + R2 * k1] = twmul<true>(v[k1], twA[k1 * R2 + i0]);
        }
      }
      pm2 = warp_max(pm2);
      if (lane == 0) red[32 + warp] = pm2;
    }
    if (!helpers) __syncthreads();
    else if (tid < hlp0) asm volatile("bar.sync 1, %0;" ::"r"(hlp0) : "memory");
    FPM_TICK(1);
    // (measured: prefetch.global.L2 of the next LED's window slice from here changes nothing, 39.7 us per update either
    //  way at Nlarge 1536 -- the window loads of S1 / C2 are not waiting for DRAM)
    if (warp == 0) {                                        // this slice's max|P|^2 -> slot [rank] of every CTA
      float m = (lane < NW) ? red[32 + lane] : 0.f;
      m = warp_max(m);
      if (lane < C) st_async_f1(mapa_u32(pmx_a + 4u * rank, lane), m, mapa_u32(rbar_a, lane));
    }
    // ===== S2: cols stage B (inverse); the results go to the owners of the row positions =====
    if (tid < R1 * ncl) {
      const int k1 = tqc, jcl = trc;
      {
        const int js = (p.xlo + jc0 + jcl) & (N - 1);
        float2 v[R2];
#pragma unroll
        for (int a = 0; a < R2; ++a) v[a] = cslab[jcl * PR + R2 * k1 + a];
        fftR<R2, true>(v);
        const int pos0 = R2 * k1, dk = pos0 / RPC;
        const uint32_t dst = mapa_u32(rslab_a + (uint32_t)(sizeof(float2) * ((pos0 % RPC) * PITCH + js)), dk);
        const uint32_t dbar = mapa_u32(rbar_a, dk);
#pragma unroll
        for (int a = 0; a < R2; ++a) st_async_f2(dst + (uint32_t)(sizeof(float2) * a * PITCH), v[a], dbar);
      }
    }
#ifdef FPM_STAGE_TIMING
    FPM_TICK(14);
#endif
    mbar_wait_cluster(rbar, ph, 0, u);                            // row slab complete (every rank's S2 bytes have landed)
    FPM_TICK(2);
    // 1/I of this CTA's S4 work items: issued now, consumed after S3
    constexpr int S4R = (RPC * R1 + NT - 1) / NT;
    constexpr bool S4PRE = (S4R * CH <= 8);
    float4 ivall[S4PRE ? S4R : 1][CH];
    if constexpr (S4PRE) {
#pragma unroll
      for (int rq = 0; rq < S4R; ++rq) {
        const int g = tid + rq * NT;
        if (g < RPC * R1) {
          const int rl = g % RPC, k1 = g / RPC;
          const float4* ip = reinterpret_cast<const float4*>(img) + ((size_t)k1 * N + rank * RPC + rl) * CH;
#pragma unroll
          for (int c = 0; c < CH; ++c) ivall[rq][c] = __ldg(ip + c);
        }
      }
    }
    // ===== S3: rows stage A (inverse) on this CTA's rows; columns outside the bbox are zero, not read =====
    for (int g = tid; g < RPC * R2; g += NT) {
      const int rl = g % RPC, j0 = g / RPC;
      float2* rp = rslab + rl * PITCH + j0;
      float2 w[NIN], v[R1];
#pragma unroll
      for (int k = 0; k < NIN; ++k) {
        const int m = m_of(k);
        const int col = j0 + R2 * m;
        const int jw = (R2 * m < H) ? col : col - N;
        w[k] = (jw >= p.xlo && jw <= p.xhi) ? rp[R2 * m] : make_float2(0.f, 0.f);
      }
      if constexpr (SIX) fft16_in6<true>(w, v);
      else {
#pragma unroll
        for (int k = 0; k < R1; ++k) v[k] = w[k];
        fftR<R1, true>(v);
      }
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) rp[R2 * k1] = twmul<true>(v[k1], twA[k1 * R2 + j0]);
    }
    __syncthreads();
    if (tid == 0 && has_next) mbar_expect_tx(wbar, (uint32_t)(sizeof(float2) * ovr * ovc));   // every waiter of the previous phase is past it
    FPM_TICK(3);
    // ===== S4: rows stage B (inverse) + amplitude replacement (fpmMain.cpp:378-393) + rows stage B' (forward) =====
#pragma unroll
    for (int rq = 0; rq < S4R; ++rq) {
      const int g = tid + rq * NT;
      if (g >= RPC * R1) break;
      const int rl = g % RPC, k1 = g / RPC;
      float4 iv[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if constexpr (S4PRE) iv[c] = ivall[rq][c];
        else iv[c] = __ldg(reinterpret_cast<const float4*>(img) + ((size_t)k1 * N + rank * RPC + rl) * CH + c);
      }
      float2* rp = rslab + rl * PITCH + R2 * k1;
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = rp[a];
      fftR<R2, true>(v);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) {
        const float4 q4 = iv[k2 >> 2];
        const int e = k2 & 3;
        const float inv_i = (e == 0) ? q4.x : (e == 1) ? q4.y : (e == 2) ? q4.z : q4.w;
        const float2 tt = cadd(v[k2], make_float2(epsr, epsi));
        const float sc = rsqrt_fast(fmaf(tt.x, tt.x, tt.y * tt.y) * inv_i);
        v[k2] = cscale(v[k2], sc);
      }
      fftR<R2, false>(v);
#pragma unroll
      for (int q = 0; q < R2; ++q) rp[q] = twmul<false>(v[q], twB[q * R1 + k1]);
    }
    __syncthreads();
    FPM_TICK(4);
    // ===== S5: rows stage A' (forward); bbox columns go to the column owners =====
    for (int g = tid; g < RPC * R2; g += NT) {
      const int rl = g % RPC, q = g / RPC;
      const float2* rp = rslab + rl * PITCH + q;
      float2 v[R1];
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) v[k1] = rp[R2 * k1];
      fftR<R1, false>(v);
      const int pos = rank * RPC + rl;
#pragma unroll
      for (int k = 0; k < NIN; ++k) {                        // (the other outputs are dead code)
        const int r = m_of(k);
        const int col = R2 * r + q;
        const int jw = (R2 * r < H) ? col : col - N;
        if (jw >= p.xlo && jw <= p.xhi) {
          const int jc = jw - p.xlo;
          const int dk = (jc * cpc_inv) >> 16;              // jc / CPC (exact for jc < 256, CPC <= 32)
          st_async_f2(mapa_u32(cslab_a + (uint32_t)(sizeof(float2) * ((jc - dk * CPC) * PR + pos)), dk), v[r], mapa_u32(cbar_a, dk));
        }
      }
    }
#ifdef FPM_STAGE_TIMING
    FPM_TICK(15);
#endif
    mbar_wait_cluster(cbar, ph, 1, u);                            // column slab complete
    FPM_TICK(5);
    float pm2c = pmx[0];                                    // max|P|^2 over the whole pupil (all ranks' slices)
#pragma unroll
    for (int k = 1; k < C; ++k) pm2c = fmaxf(pm2c, pmx[k]);
    // ===== S6: cols stage B' (forward) =====
    if (tid < R1 * ncl) {
      const int k1 = tqc, jcl = trc;
      {
        float2* cp = cslab + jcl * PR + R2 * k1;
        float2 v[R2];
#pragma unroll
        for (int a = 0; a < R2; ++a) v[a] = cp[a];
        fftR<R2, false>(v);
#pragma unroll
        for (int q = 0; q < R2; ++q) cp[q] = twmul<false>(v[q], twB[q * R1 + k1]);
      }
    }
    if (EDGE_LATE && helpers) {                              // the column threads among themselves; helpers: edge pixels
      if (tid < hlp0) asm volatile("bar.sync 1, %0;" ::"r"(hlp0) : "memory");
      else edge_maxima(tid - hlp0, hn);
    } else __syncthreads();
    FPM_TICK(6);
    // ===== S7: cols stage A' (forward) -> Phi' in natural row order; only bbox rows are stored =====
    if (tid < R2 * ncl) {
      const int q = tqc, jcl = trc;
      {
        float2* cp = cslab + jcl * PR;
        float2 v[R1];
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = cp[R2 * k1 + q];
        fftR<R1, false>(v);
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
          const int r = m_of(k);
          const int i = R2 * r + q;
          const int iw = (i < H) ? i : i - N;
          if (iw >= p.ylo && iw <= p.yhi) cp[i] = v[r];
        }
      }
    }
    __syncthreads();
    FPM_TICK(7);
    // ===== C2: object update (fpmMain.cpp:406-447) and pupil-increment numerator (:459-472) on this CTA's columns =====
    {
      const float inv_pmax = rsqrt_fast(pm2c);
      float2* wr = objFc + (size_t)r0 * L + c0 + jc0;
      constexpr int UN = (N == 256) ? 9 : 6;                // window elements in flight per thread
      const int n = NR * ncl;
      const int qNT = qNTc, rNT = rNTc;
      int ir0 = tqc, jl0 = trc;
      for (int base = tid; base < n; base += UN * NT) {
        float2 Ov[UN];
        {
          int ir = ir0, jl = jl0;
#pragma unroll
          for (int k = 0; k < UN; ++k) {
            if constexpr (WS) Ov[k] = (base + k * NT < n) ? Wcur[ir * CPC + jl] : make_float2(0.f, 0.f);
            else Ov[k] = (base + k * NT < n) ? wr[(size_t)ir * L + jl] : make_float2(0.f, 0.f);
            ir += qNT; jl += rNT;
            if (jl >= ncl) { jl -= ncl; ++ir; }
          }
        }
#pragma unroll
        for (int k = 0; k < UN; ++k) {
          if (base + k * NT < n) {
            const int ir = ir0, lane_c = jl0;
            float a2n = 0.f;
            {
              const int iw = p.ylo + ir, i = iw & (N - 1);
              const int e = ir * CPC + lane_c;
              const float2 O = Ov[k];
              const float2 Pv = Pc[e];
              const float2 d = csub(cslab[lane_c * PR + i], cmul(O, Pv));         // dPhi = Phi' - Phi
              const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
              const float2 num = cmulc(d, Pv);
              const float A = pa2 + p.delta2;
              const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
              const float2 On = make_float2(O.x + (num.x * A + num.y * kd2) * sc, O.y + (num.y * A - num.x * kd2) * sc);
              wr[(size_t)ir * L + lane_c] = On;
              if constexpr (WS) {
                // forward into the next window's slice of the CTA that owns this column there
                const int rn = r0 + ir - r0n, cn = c0 + jc0 + lane_c - c0n;
                if (has_next && (unsigned)rn < (unsigned)NR && (unsigned)cn < (unsigned)NC) {
                  const int dk = (cn * cpc_inv) >> 16;                          // cn / CPC
                  st_async_f2(mapa_u32(wnxt_a + (uint32_t)(sizeof(float2) * (rn * CPC + cn - dk * CPC)), dk), On, mapa_u32(wbar_a, dk));
                }
              }
              a2n = fmaf(On.x, On.x, On.y * On.y);
              const float oa2 = fmaf(O.x, O.x, O.y * O.y);
              const float2 numq = cmulc(d, O);
              const float A1 = oa2 + p.delta1;
              const float sq = sqrt_fast(oa2) * Sc[e] * rcp_fast(fmaf(A1, A1, kd1 * kd1));
              Qc[e] = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
            }
            W[ir * CPC + lane_c] = a2n;
          }
          ir0 += qNT; jl0 += rNT;
          if (jl0 >= ncl) { jl0 -= ncl; ++ir0; }
        }
      }
    }
    __syncthreads();
    FPM_TICK(8);
    // ===== D: exact max|objF|.  Partial maxima of the touched cells: this slice's new values (from W) ... =====
    if (ncl > 0) {
      const int ca0 = c0 + jc0, ca1 = ca0 + ncl - 1;                    // absolute columns of the slice
      const int lcc0 = ca0 >> 4, nlc = (ca1 >> 4) - lcc0 + 1;
      const int lsh = 32 - __clz(nlc - 1);
      for (int t = tid; t < (NR << lsh); t += NT) {
        const int ir = t >> lsh, bl = t & ((1 << lsh) - 1);
        if (bl >= nlc) continue;
        const int lo = max(ca0, (lcc0 + bl) << 4), hi = min(ca1, ((lcc0 + bl) << 4) + 15);
        const float* wp = W + ir * CPC - ca0;
        float m = 0.f;
        for (int c = lo; c <= hi; ++c) m = fmaxf(m, wp[c]);
        atomicMax(&Tm[(((r0 + ir) >> p.cs) - cr0) * tmc + (lcc0 + bl - cc0)], __float_as_uint(m));
      }
    }
    __syncthreads();
#ifdef FPM_STAGE_TIMING
    FPM_TICK(12);
#endif
    // max|objF|^2 = max over the ranks of: the rank's partial maxima of the touched cells (its slice's new values, its share
    // of the edge pixels) and the maximum of its share of the untouched cells (side work above) -- one exchange.
    if (warp == 0) {
      float m = (lane < NW) ? red[lane] : 0.f;
      for (int t = lane; t < lay.tmr * lay.tmc; t += 32) m = fmaxf(m, __uint_as_float(Tm[t]));
      m = warp_max(m);
      if (lane < C) st_async_f1(mapa_u32(omx_a + 4u * rank, lane), m, mapa_u32(obar_a, lane));
    }
    __syncthreads();                                         // (warp 0 has read Tm)
    // Off the critical path (the exchange is in flight): the partial maxima go to the owners' grids, which the NEXT
    // update's side work scans after waiting for ubar.
    for (int t = tid; t < (ncr << csh); t += NT) {
      const int ta = t >> csh, tb = t & ((1 << csh) - 1);
      if (tb >= ncc) continue;
      const unsigned v = Tm[ta * tmc + tb];
      Tm[ta * tmc + tb] = 0u;
      if (v) {
        const int a = cr0 + ta;
        atomicMax(reinterpret_cast<unsigned*>(cluster.map_shared_rank(U, a % C)) + (a / C) * gc + cc0 + tb, v);
      }
    }
    __syncthreads();                                         // (Tm is clear before the next update's side work fills it)
#ifdef FPM_STAGE_TIMING
    FPM_TICK(13);
#endif
    if (tid == duty0) {
      // one cluster-scope release for the whole CTA (cumulative over the barrier above): the merged maxima and this
      // update's spectrum stores are visible to whoever observes the arrivals.  Issued by a helper thread: the ~1k cycles
      // of the fence overlap the exchange and the next update's S1.
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
#pragma unroll
      for (int k = 0; k < C; ++k) mbar_arrive_remote(mapa_u32(ubar_a, k));
    }
    FPM_TICK(9);
    mbar_wait_cluster(obar, ph, 3, u);
    {
      float om2 = omx[0];
#pragma unroll
      for (int k = 1; k < C; ++k) om2 = fmaxf(om2, omx[k]);
      inv_objf_max = rsqrt_fast(om2);
    }
    FPM_TICK(10);
  }

#ifdef FPM_STAGE_TIMING
  if (tid == FPM_TICK_TID && blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p.stage_clk[k] += tacc_[k];
  }
#endif
  // the last update's pupil increment is still pending
  for (int t = tid; t < NR * CPC; t += NT) {
    const int ir = t / CPC, jcl = t - ir * CPC;
    if (jcl < ncl) {
      const float2 Q = Qc[t];
      float2 v = Pc[t];
      v.x = fmaf(Q.x, inv_objf_max, v.x);
      v.y = fmaf(Q.y, inv_objf_max, v.y);
      const int iw = p.ylo + ir, jw = p.xlo + jc0 + jcl;
      Pg[(iw & (N - 1)) * N + (jw & (N - 1))] = v;
    }
  }
  cluster.sync();                                           // no CTA exits while its shared memory may still be addressed
}

}  // namespace fpm
