// fpm_update.cuh -- the fused sub-aperture update kernel (sm_100a).
//
// One CTA per tile, persistent over the whole (iteration x LED) sequence of that tile: the low-res
// exit-wave field never leaves shared memory between "crop * pupil" and the object / pupil update.
// Restates one body of the loop nest fpmMain.cpp:345-476 per "update" (SURVEY.md appendix A).
//
// Index conventions
//   window (i,j)      DC-at-corner, the reference's Objfcrop (fpmMain.cpp:361)
//   wrapped (iw,jw)   iw = i < N/2 ? i : i-N
//   absolute (r,c)    centred spectrum objFc: r = ys + N/2 + iw, c = xs + N/2 + jw  (the two fftShifts of
//                     fpmMain.cpp:358,361 folded into addressing)
//   support bbox      wrapped ranges [ylo,yhi] x [xlo,xhi] containing every non-zero pupilSupport pixel.
//                     P == 0 outside it for ever (P starts as the support, every increment is masked:
//                     fpmMain.cpp:313,472), so O*P, dO and dP vanish there and everything that touches
//                     P is restricted to the bbox with bit-identical results.
//
// N x N transform, separable, N = R1*R2 per dimension (64 = 8*8, 128 = 16*8, 256 = 16*16): four
// in-register radix stages per transform, the field exchanged through shared memory in between.
//   IFFT (DIF, natural in -> digit-scrambled out):  S1 cols-A, S2 cols-B, S3 rows-A, S4 rows-B
//   FFT  (DIT, scrambled in -> natural out):        S4 rows-B', S5 rows-A', S6 cols-B', S7 cols-A'
// S4 runs the last inverse stage, the amplitude replacement (fpmMain.cpp:378-393) and the first forward
// stage on the same registers; scrambled position p = R2*k1 + k2 holds index k1 + R1*k2, so no reorder
// pass exists.  Column stages only visit bbox columns; S3 reads bbox columns only (the rest is zero).
//   C2  object update on the bbox (fpmMain.cpp:406-447).  The window O was staged by TMA (3-D tensor map over
//       objFc, mbarrier completion) one update ahead; the new values are written into the window buffer, forwarded
//       into the next LED's window where the two overlap, and go back to objFc with one TMA store.
//       Q = pupil-update numerator (fpmMain.cpp:459-472) stays in shared memory.
//   D   exact max|objF| (fpmMain.cpp:460,467) from a grid U of per-cell maxima (2^cs rows x 16 columns): the cells
//       the rectangle touches are rebuilt from the on-chip new values + the unchanged outside pixels of the edge
//       cells; the maximum is a flat scan of U.
//   The pupil update P += Q / max|objF| (fpmMain.cpp:470-475) is applied by the next update's S1 (or the epilogue).
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_regs.cuh"

#ifdef FPM_STAGE_TIMING
#define FPM_TICK(k) do { long long t_ = clock64(); tacc_[k] += t_ - tprev_; tprev_ = t_; } while (0)   /* registers only */
#else
#define FPM_TICK(k) do {} while (0)
#endif

namespace fpm {

struct UpdateParams {
  CUtensorMap tmap;         // objFc as a 3-D float tensor (2*L, L, n_tiles); box = (2*ocp, NR, 1): one TMA = one bbox window
  float2* objFc;            // [n_tiles][L][L]   centred spectrum
  float2* pupil;            // [n_tiles][N][N]   DC-at-corner
  const float* stack;       // [n_tiles][n_leds][N*N]: 1/intensity (inf for 0) in the permuted device layout (stack_offset)
  const float* support;     // [N][N]
  const short2* crop;       // [n_leds] (x = cropXStart, y = cropYStart)
  const float2* tw;         // [N] exp(-2*pi*i*n/N)
  float2* field_gmem;       // [n_tiles][N][N+1] scratch when the field does not fit shared memory
  int L, n_leds;
  int tile0;                // first tile of this launch
  const int* tile_list;     // when set: CTA b works on tile tile_list[b] & 0x3fffffff (balanced passes, fpmb200_run); else tile0 + b
  float* ucache;            // fpm_update_phased_kernel: [n_tiles][L][L/16] cell maxima handed from pass to pass (or NULL)
  int slot_begin, n_updates;
  float delta1, delta2, eps, kappa;
  int ylo, yhi, xlo, xhi;   // support bbox (wrapped)
  int ocp;                  // row pitch (float2) of the on-chip window copies: even and >= NC+1 (TMA boxes start on 16 B)
  int cs;                   // log2 rows per max-cell (cells are (1<<cs) rows x 16 columns)
  long long* stage_clk;     // [16] per-stage cycle totals of CTA 0 (only with -DFPM_STAGE_TIMING)
  int noff[12];             // fpm_update_phased_kernel: byte offsets of its box-dependent shared-memory arrays (PhasedShape::layout)
};

template <int N> struct Shape {
  static constexpr int R1 = (N == 64) ? 8 : 16;
  static constexpr int R2 = N / R1;
  static constexpr int PITCH = N + 1;       // float2 per field row, odd: the 32 rows a warp touches in a row pass
                                            // (lanes = rows, same column) land on distinct bank pairs; column passes
                                            // (lanes = consecutive columns) are conflict-free for any pitch
  static constexpr int CH = R2 / 4;         // float4 chunks of 1/intensity per S4 work item
};

// Device layout of one N x N intensity image: S4 work item (scrambled row position p, k1) needs the R2 pixels
// x = k1 + R1*k2 of spatial row y = p/R2 + R1*(p%R2); they are stored contiguously at item index k1*N + p, so a
// warp (32 consecutive p, one k1) reads 32 adjacent chunks.  Returns the element offset of pixel (y, x).
// The device keeps 1/I as float (converted once at upload) so that the amplitude replacement
// sqrt(I) psi/|psi+eps| = psi * rsqrt(|psi+eps|^2 / I) costs one MUFU per pixel.
template <int N> __host__ __device__ __forceinline__ int stack_offset(int y, int x) {
  using S = Shape<N>;
  const int pos = S::R2 * (y % S::R1) + y / S::R1;
  return ((x % S::R1) * N + pos) * S::R2 + x / S::R1;
}

// maximum of non-negative floats over a warp: one REDUX on the bit patterns (order-preserving for v >= 0)
#ifndef FPM_EXP
#define FPM_EXP 0
#endif
__device__ __forceinline__ float warp_max(float v) {
#if FPM_EXP & 2
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
#else
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
#endif
}
// maximum over each aligned half-warp (xor shuffles stay inside a half; a per-half REDUX mask makes the
// compiler emit a divergence-handling loop that costs ~2k cycles per call site)
__device__ __forceinline__ float half_warp_max(float v, int lane) {
  (void)lane;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// 1/x (MUFU.RCP; the denominators of the update steps are (|.|^2 + delta)^2 + (kappa delta)^2 >= delta^2 > 0, far from the
// range where __fdividef has to rescale -- its range checks cost ~6 instructions per quotient)
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
#if FPM_EXP & 16
  bytes = 0;
#endif
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 3-D box (x = float index in the row, y = row, z = tile) global -> shared, completion on `bar`
__device__ __forceinline__ void tma_load_window(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
#if FPM_EXP & 16
  (void)dst; (void)map; (void)x; (void)y; (void)z; (void)bar; return;
#endif
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// one 3-D box shared -> global (bulk async group)
__device__ __forceinline__ void tma_store_window(const void* src, const CUtensorMap* map, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// NARROW: the support bbox lies inside [-(3*R2-1), 3*R2-1] in both directions (N = 128: a pupil radius <= 23, the shipped
// 128-pixel configurations).  Then only 6 of the 16 inputs of every first-stage radix-16 butterfly can be non-zero
// (S1, S3: fft16_in6) and only 6 of the 16 outputs of every last-stage butterfly are ever read (S5, S7: the others are
// not stored and their arithmetic is dead code); the set is known at compile time.
template <int N, int NT, int MINB, bool FIELD_SMEM, bool P_SMEM, bool Q_SMEM, bool NARROW = false>
__global__ void __launch_bounds__(NT, MINB) fpm_update_kernel(const __grid_constant__ UpdateParams p) {
  using S = Shape<N>;
  constexpr int R1 = S::R1, R2 = S::R2, PITCH = S::PITCH, CH = S::CH;
  static_assert(!NARROW || (R1 == 16 && P_SMEM && Q_SMEM), "narrow-pupil specialisation");
  constexpr int H = N / 2, NW = NT / 32;
  constexpr bool PIPE = !(FPM_EXP & 4);       // software-pipelined row passes (next item's loads ahead of this item's stores)
  extern __shared__ __align__(1024) unsigned char smem_raw[];   // TMA destinations need 128-byte alignment

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = p.tile_list ? (p.tile_list[blockIdx.x] & 0x3fffffff) : p.tile0 + blockIdx.x;
  const int L = p.L;
  const int NR = p.yhi - p.ylo + 1, NC = p.xhi - p.xlo + 1;
  const int gc = L >> 4, gr = L >> p.cs;                 // max-cells are (1<<cs) rows x 16 columns
  const int qNT = NT / NC, rNT = NT % NC;                 // element index stepping: t += NT without dividing
  const int tq = tid / NC, tr = tid - tq * NC;            // (row, column) of work item tid in every NC-wide decomposition
  auto step_nt = [&](int& q, int& r) { q += qNT; r += rNT; if (r >= NC) { r -= NC; ++q; } };
  const int tmr = (NR >> p.cs) + 2, tmc = (NC >> 4) + 2;   // cells a bbox rectangle can touch
  const int wshmax = 32 - __clz((tmc << 4) - 1);           // log2 of the padded width of the touched-cell window

  // ---- shared memory carve-up ----
  unsigned char* sp = smem_raw;
  float2* fld;
  if constexpr (FIELD_SMEM) { fld = reinterpret_cast<float2*>(sp); sp += (sizeof(float2) * N * PITCH + 15) / 16 * 16; }
  else fld = p.field_gmem + (size_t)tile * N * PITCH;
  // twiddles as twmul4 operands: twA (inverse stages) [b*R2 + a] = conj(W^(a*b)), twB (forward stages) [a*R1 + b] = W^(a*b)
  float4* twA = reinterpret_cast<float4*>(sp); sp += sizeof(float4) * N;
  float4* twB = reinterpret_cast<float4*>(sp); sp += sizeof(float4) * N;
  float* red = reinterpret_cast<float*>(sp);   sp += sizeof(float) * 64;   // [0..31] objF, [32..63] pupil partial maxima
  float2* Pc = nullptr;
  if constexpr (P_SMEM) { Pc = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * NR * NC; }
  float2* Qc = nullptr;                                                    // Q = pupil increment * max|objF| of the previous update
  float2* Ocb[2] = {nullptr, nullptr};                                     // windows O of this and the next update (TMA destinations)
  const int OCP = p.ocp;
  if constexpr (Q_SMEM) {
    Qc = reinterpret_cast<float2*>(sp); sp += sizeof(float2) * NR * NC;
    sp += (128u - (smem_u32(sp) & 127u)) & 127u;          // absolute 128-byte alignment (the dynamic base is only 16-aligned)
    Ocb[0] = reinterpret_cast<float2*>(sp); sp += ((sizeof(float2) * NR * OCP + 127) / 128) * 128;
    Ocb[1] = reinterpret_cast<float2*>(sp); sp += ((sizeof(float2) * NR * OCP + 127) / 128) * 128;
  }
  __shared__ __align__(8) uint64_t wbar;                                   // completion barrier of the window TMA
  const uint32_t win_bytes = (uint32_t)(sizeof(float2) * NR * OCP);
  float* Sc = nullptr;                                                     // support on the bbox
  if constexpr (Q_SMEM) { Sc = reinterpret_cast<float*>(sp); sp += sizeof(float) * NR * NC; }
  unsigned* Tm = reinterpret_cast<unsigned*>(sp); sp += sizeof(unsigned) * tmr * tmc;   // new maxima of the touched cells (bit patterns)
  float* W = nullptr;                                                      // |.|^2 of every pixel of the touched cells
  if constexpr (Q_SMEM) { sp = smem_raw + ((size_t)(sp - smem_raw) + 15) / 16 * 16; W = reinterpret_cast<float*>(sp); sp += sizeof(float) * ((tmr << p.cs) << wshmax); }
  sp = smem_raw + ((size_t)(sp - smem_raw) + 15) / 16 * 16;
  float* U = reinterpret_cast<float*>(sp);                                 // [gr][gc] exact cell maxima of |objFc|^2

  float2* objFc = p.objFc + (size_t)tile * L * L;
  float2* Pg = p.pupil + (size_t)tile * N * N;
  const float* __restrict__ stack = p.stack + (size_t)tile * p.n_leds * N * N;

  auto Pref = [&](int iw, int jw) -> float2& {
    if constexpr (P_SMEM) return Pc[(iw - p.ylo) * NC + (jw - p.xlo)];
    else return Pg[(iw & (N - 1)) * N + (jw & (N - 1))];
  };
  auto Qref = [&](int iw, int jw) -> float2& {
    if constexpr (Q_SMEM) return Qc[(iw - p.ylo) * NC + (jw - p.xlo)];
    // no room for a separate buffer: Q takes the place of Phi' in the field (C2 reads Phi'(i,j) and then writes
    // Q(i,j) from the same thread; pass E consumes it before the next S1 overwrites the field)
    else return fld[(iw & (N - 1)) * PITCH + (jw & (N - 1))];
  };
  // Exact maxima of the two cells (cellrow, 2*seg) and (cellrow, 2*seg+1) from memory: one warp, 32 columns.
  // Every lane of a half-warp returns its cell's maximum.
  auto cell_pair_max = [&](int cellrow, int seg) -> float {
    const float2* src = objFc + (size_t)(cellrow << p.cs) * L + (seg << 5) + lane;
    float cm = 0.f;
    for (int rr = 0; rr < (1 << p.cs); ++rr) {
      const float2 o = src[(size_t)rr * L];
      cm = fmaxf(cm, fmaf(o.x, o.x, o.y * o.y));
    }
    return half_warp_max(cm, lane);
  };
  // Row passes: the WPB = NT/N consecutive warps {WPB*b .. WPB*b + WPB-1} own the 32 rows of block b in S3, S4 and
  // S5 alike (lane = row inside the block, the warp's rank inside the group picks the sub-transforms), so a named
  // barrier over those warps replaces the block-wide barrier between the row stages.  Consecutive warps sit on
  // different SM sub-partitions: every scheduler holds one warp of each row block, and the blocks drift apart
  // (one block's barrier wait is filled with another block's butterflies).
  static_assert(NT % N == 0 && N % 32 == 0 && N / 32 <= 15, "row-block barriers");
  constexpr int WPB = NT / N;
  const int rb_row = 32 * (warp / WPB) + lane, rb_sub = warp % WPB;
  auto row_block_sync = [&]() {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / WPB), "r"(WPB * 32) : "memory");
  };
  // Column stage B work items (k1, bbox column jc), R1 * NC of them.  A warp takes 32 consecutive columns of ONE k1
  // (64 consecutive shared-memory words per access: no bank conflicts, k1 and its twiddles warp-uniform); the NC % 32
  // left-over columns of every k1 are packed into a short second round, one item per thread (its index is computed
  // once: an integer division on the per-update path of the critical warps costs more than the conflicts did).
  const int cb_nfull = NC >> 5, cb_nl = NC & 31;
  static_assert(R1 * 31 <= NT, "one left-over item per thread");
  const int cb_k1 = tid / max(cb_nl, 1), cb_jc = (cb_nfull << 5) + tid % max(cb_nl, 1);
  auto col_items_B = [&](auto&& body) {
    for (int wi = warp; wi < R1 * cb_nfull; wi += NW) body(wi % R1, ((wi / R1) << 5) + lane);
    if (cb_k1 < R1 && cb_nl) body(cb_k1, cb_jc);
  };
  // ---- prologue: tables, pupil -> shared memory, max|P|^2, max-cell grid, first window ----
  for (int t = tid; t < N; t += NT) {
    const int b = t / R2, a = t % R2;
    const float2 wa = p.tw[a * b];
    twA[t] = make_float4(wa.x, -wa.y, wa.y, wa.x);
    const int a2 = t / R1, b2 = t % R1;
    const float2 wb = p.tw[a2 * b2];
    twB[t] = make_float4(wb.x, wb.y, -wb.y, wb.x);
  }
  for (int t = tid; t < NR * NC; t += NT) {
    const int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
    const int gi = (iw & (N - 1)) * N + (jw & (N - 1));
    if constexpr (P_SMEM) Pc[t] = Pg[gi];
    if constexpr (Q_SMEM) { Sc[t] = p.support[gi]; Qc[t] = make_float2(0.f, 0.f); }
  }
  for (int it = warp; it < gr * (L >> 5); it += NW) {
    const int cellrow = it / (L >> 5), seg = it % (L >> 5);
    const float cm = cell_pair_max(cellrow, seg);
    if ((lane & 15) == 0) U[cellrow * gc + 2 * seg + (lane >> 4)] = cm;
  }
  for (int t = tid; t < tmr * tmc; t += NT) Tm[t] = 0u;
  if constexpr (!Q_SMEM) {                           // max|P|^2 of the incoming pupil (later maintained by pass E)
    float pm2 = 0.f;
    for (int t = tid; t < NR * NC; t += NT) {
      const int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
      const float2 v = Pg[(iw & (N - 1)) * N + (jw & (N - 1))];
      pm2 = fmaxf(pm2, fmaf(v.x, v.x, v.y * v.y));
    }
    pm2 = warp_max(pm2);
    if (lane == 0) red[32 + warp] = pm2;
  }
  // crop origins of this update, the next and the one after (windows are fetched one update ahead)
  short2 cr_a = p.crop[p.slot_begin % p.n_leds], cr_b = p.crop[(p.slot_begin + 1) % p.n_leds];
  uint32_t wphase = 0;
  if constexpr (Q_SMEM) {
    if (tid == 0) {
      mbar_init(&wbar, 1);
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_expect_tx(&wbar, 2 * win_bytes);
      // TMA box starts must be 16-byte aligned: fetch from the even column at or left of the window (ocp has room)
      tma_load_window(Ocb[0], &p.tmap, 2 * ((cr_a.x + H + p.xlo) & ~1), cr_a.y + H + p.ylo, tile, &wbar);
      tma_load_window(Ocb[1], &p.tmap, 2 * ((cr_b.x + H + p.xlo) & ~1), cr_b.y + H + p.ylo, tile, &wbar);
    }
  }
  __syncthreads();

  const float kd1 = p.kappa * p.delta1, kd2 = p.kappa * p.delta2;
  // psi' = a psi/|psi + eps| with psi = raw/N^2  ==  a raw/|raw + N^2 eps|: the 1/N^2 of ifft2 is never applied
  const float epsr = p.eps * (float)(N * N), epsi = p.kappa * epsr;

  float inv_objf_max = 0.f;          // 1 / max|objF| after the previous update (its Q is still pending in Qc)
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
    float2* const Ocur = (u & 1) ? Ocb[1] : Ocb[0];                // (selects, not a runtime-indexed local array)
    float2* Oc = Ocur + ((xs + H + p.xlo) & 1);                    // window element (ir, jc) = Oc[ir*OCP + jc]
    float2* Ocn = (u & 1) ? Ocb[0] : Ocb[1];                       // next window's box, box-relative columns
    if constexpr (Q_SMEM) { if (u == 0) { mbar_wait(&wbar, wphase); wphase ^= 1; } }
    const float* __restrict__ img = stack + (size_t)slot * N * N;
    if (tid == 0)   // pull the next LED's intensity tile towards L2 while this update runs
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(stack + (size_t)nslot * N * N), "r"((unsigned)(N * N * 4)) : "memory");
    float2* wbase = objFc + (size_t)(ys + H) * L + (xs + H);   // absolute address of (iw=0, jw=0)
    // (Measured on the large-pupil variant -- window in global memory, cfg5: 85 x 85 box, 6.4 M updates/s at 148 tiles:
    //  prefetch.global.L2 of the next window from here: 6.41 M against 6.47 M without; C2 / E restructured row-wise
    //  (32 consecutive columns of a row per warp, uniform row addresses, one shared atomic per max-cell, 4 or 8 items in
    //  flight): 5.85 M.  Neither DRAM latency nor the per-element index arithmetic is what binds C2 there.)

    // ===== S1: pending pupil update P += Q / max|objF| (fpmMain.cpp:470-475) for the rows this thread owns, then
    //       Phi = O*P and cols stage A (inverse).  max|P|^2 for this update's object step is reduced on the way. =====
    {
      float pm2 = 0.f;
      for (int g = tid, i0 = tq, jc = tr; g < R2 * NC; g += NT) {
        const int jw = p.xlo + jc, j = jw & (N - 1);
        float2 v[R1];
        if constexpr (NARROW) {
          // inputs i = i0 + R2*m inside the bbox: m = 0, 1, 2 (iw = i) and m = R1-3 .. R1-1 (iw = i - N) at most
          float2 w6[6];
          const int ob = (i0 - p.ylo) * OCP + jc, pb = (i0 - p.ylo) * NC + jc;
          static_for<0, 6>([&](auto K) {
            constexpr int k = decltype(K)::value;
            constexpr int off = (k < 3) ? R2 * k : R2 * (R1 - 6 + k) - N;      // iw = i0 + off
            const int iw = i0 + off;
            const bool in = (iw >= p.ylo) && (iw <= p.yhi);
            const int oi = in ? ob + off * OCP : 0, pi = in ? pb + off * NC : 0;
            const float2 O = Oc[oi], Q = Qc[pi];
            float2 Pv = Pc[pi];
            Pv.x = fmaf(Q.x, inv_objf_max, Pv.x);
            Pv.y = fmaf(Q.y, inv_objf_max, Pv.y);
            if (in) Pc[pi] = Pv;
            pm2 = fmaxf(pm2, in ? fmaf(Pv.x, Pv.x, Pv.y * Pv.y) : 0.f);
            const float2 phi = cmul(O, Pv);
            w6[k] = in ? phi : make_float2(0.f, 0.f);
          });
          FPM_TICK(11);
          fft16_in6<true>(w6, v);
        } else if constexpr (Q_SMEM) {
#pragma unroll
          for (int m = 0; m < R1; ++m) {
            const int i = i0 + R2 * m;
            const int iw = (i < H) ? i : i - N;
            if (iw >= p.ylo && iw <= p.yhi) {
              const float2 O = Oc[(iw - p.ylo) * OCP + jc];
              const float2 Q = Qref(iw, jw);
              float2& pr = Pref(iw, jw);
              float2 Pv = pr;
              Pv.x = fmaf(Q.x, inv_objf_max, Pv.x);
              Pv.y = fmaf(Q.y, inv_objf_max, Pv.y);
              pr = Pv;
              pm2 = fmaxf(pm2, fmaf(Pv.x, Pv.x, Pv.y * Pv.y));
              v[m] = cmul(O, Pv);
            } else v[m] = make_float2(0.f, 0.f);
          }
        } else {
          // window in global memory (the pupil update ran as its own pass E): every load of the work item is issued
          // before the first use -- one exposed L2 latency per item instead of one per element
#pragma unroll
          for (int m = 0; m < R1; ++m) {
            const int i = i0 + R2 * m;
            const int iw = (i < H) ? i : i - N;
            const int iwc = min(max(iw, p.ylo), p.yhi);             // (clamped, unconditional: predicated loads of the in-box
            v[m] = wbase[iwc * L + jw];                              //  rows only were slower, 10.2 k against 6.9 k cycles)
          }
#pragma unroll
          for (int m = 0; m < R1; ++m) {
            const int i = i0 + R2 * m;
            const int iw = (i < H) ? i : i - N;
            const bool in = (iw >= p.ylo && iw <= p.yhi);
            const float2 phi = cmul(v[m], Pref(min(max(iw, p.ylo), p.yhi), jw));
            v[m] = in ? phi : make_float2(0.f, 0.f);
          }
        }
        if constexpr (!NARROW) { FPM_TICK(11); fftR<R1, true>(v); }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1)
          fld[(i0 + R2 * k1) * PITCH + j] = twmul4(v[k1], twA[k1 * R2 + i0]);
        step_nt(i0, jc);
      }
      if constexpr (Q_SMEM) {
        pm2 = warp_max(pm2);
        if (lane == 0) red[32 + warp] = pm2;
      }
    }
    __syncthreads();
    FPM_TICK(1);
    // ================= S2: cols stage B (inverse) =================
    col_items_B([&](int k1, int jc) {
      const int js = (p.xlo + jc) & (N - 1);
      float2 v[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a] = fld[(R2 * k1 + a) * PITCH + js];
      fftR<R2, true>(v);
#pragma unroll
      for (int a = 0; a < R2; ++a) fld[(R2 * k1 + a) * PITCH + js] = v[a];
    });
    __syncthreads();
    FPM_TICK(2);
    if constexpr (Q_SMEM) {
      // Window of update u+1 -> the buffer update u-1 released.  Its TMA store was issued a third of an update ago, so
      // the wait is free; the load has until C2 (which forwards this update's values into it) to land.
      if (tid == 0 && u > 0) {
        tma_store_wait_all();
        // (a fence.proxy.async without a state space compiles to MEMBAR.ALL.GPU: ~1k cycles on this warp per update)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&wbar, win_bytes);
        tma_load_window(Ocn, &p.tmap, 2 * ((cr_b.x + H + p.xlo) & ~1), cr_b.y + H + p.ylo, tile, &wbar);
      }
    }
    // ===== S3: rows stage A (inverse); columns outside the bbox are zero, not read.  Lanes run over rows. =====
    // (a thread's work items touch disjoint elements of its row, so the inputs of item q+1 are loaded before the
    //  outputs of item q are stored: the shared-memory latency of the next item hides behind this item's butterfly)
    if constexpr (NARROW && PIPE) {
      constexpr int NI = R2 / WPB;
      auto load6 = [&](int j0, float2 (&w)[6]) {
        const float2* rp = fld + rb_row * PITCH + j0;
        static_for<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int m = (k < 3) ? k : R1 - 6 + k;
          const int jw = (k < 3) ? j0 + R2 * m : j0 + R2 * m - N;
          w[k] = (jw >= p.xlo && jw <= p.xhi) ? rp[R2 * m] : make_float2(0.f, 0.f);       // (j0 is warp-uniform)
        });
      };
      float2 wn[6];
      load6(rb_sub, wn);
      static_for<0, NI>([&](auto Q) {
        constexpr int qi = decltype(Q)::value;
        const int j0 = rb_sub + qi * WPB;
        float2 w6[6], v[R1];
#pragma unroll
        for (int k = 0; k < 6; ++k) w6[k] = wn[k];
        if constexpr (qi + 1 < NI) load6(j0 + WPB, wn);
        fft16_in6<true>(w6, v);
        float2* rp = fld + rb_row * PITCH + j0;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) rp[R2 * k1] = twmul4(v[k1], twA[k1 * R2 + j0]);
      });
    } else
    for (int j0 = rb_sub; j0 < R2; j0 += WPB) {
      float2* rp = fld + rb_row * PITCH + j0;
      float2 v[R1];
      if constexpr (NARROW) {
        float2 w6[6];
        static_for<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int m = (k < 3) ? k : R1 - 6 + k;
          const int jw = (k < 3) ? j0 + R2 * m : j0 + R2 * m - N;
          w6[k] = (jw >= p.xlo && jw <= p.xhi) ? rp[R2 * m] : make_float2(0.f, 0.f);     // (j0 is warp-uniform)
        });
        fft16_in6<true>(w6, v);
      } else {
#pragma unroll
        for (int m = 0; m < R1; ++m) {
          const int col = j0 + R2 * m;
          const int jw = (R2 * m < H) ? col : col - N;                 // R2 | H: the group does not straddle H
          v[m] = (jw >= p.xlo && jw <= p.xhi) ? rp[R2 * m] : make_float2(0.f, 0.f);
        }
        fftR<R1, true>(v);
      }
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) rp[R2 * k1] = twmul4(v[k1], twA[k1 * R2 + j0]);
    }
    row_block_sync();          // S3 -> S4 -> S5 exchange data only within a block of 32 rows (= NT/N warps)
    FPM_TICK(3);
    // ===== S4: rows stage B (inverse) + amplitude replacement + rows stage B' (forward) =====
    constexpr int S4R = R1 / WPB;                               // work items per thread
    static_assert(R1 % WPB == 0 && R2 % WPB == 0, "row work items per warp");
    constexpr bool S4PRE = (S4R * CH <= 8);                      // all 1/I up front when they fit 32 registers
    float4 ivall[S4PRE ? S4R : 1][CH];
    if constexpr (S4PRE) {
#pragma unroll
      for (int rq = 0; rq < S4R; ++rq) {
        const int g = (rb_sub + rq * WPB) * N + rb_row;                // item (k1, row): index k1*N + row
        const float4* ip = reinterpret_cast<const float4*>(img) + (size_t)g * CH;   // permuted layout: item g owns R2 pixels
#pragma unroll
        for (int c = 0; c < CH; ++c) ivall[rq][c] = __ldg(ip + c);
      }
    }
    constexpr bool PIPE4 = PIPE && (R2 <= 8);
    float2 vnext[PIPE4 ? R2 : 1];
    if constexpr (PIPE4) {
      const float2* rp0 = fld + rb_row * PITCH + R2 * rb_sub;
#pragma unroll
      for (int a = 0; a < R2; ++a) vnext[a] = rp0[a];
    }
#pragma unroll
    for (int rq = 0; rq < S4R; ++rq) {
      const int k1 = rb_sub + rq * WPB, row = rb_row;
      const int g = k1 * N + row;
      float4 iv[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if constexpr (S4PRE) iv[c] = ivall[rq][c];
        else iv[c] = __ldg(reinterpret_cast<const float4*>(img) + (size_t)g * CH + c);
      }
      float2* rp = fld + row * PITCH + R2 * k1;
      float2 v[R2];
      if constexpr (PIPE4) {
#pragma unroll
        for (int a = 0; a < R2; ++a) v[a] = vnext[a];
        if (rq + 1 < S4R) {
#pragma unroll
          for (int a = 0; a < R2; ++a) vnext[a] = rp[R2 * WPB + a];      // item rq+1: k1 + WPB
        }
      } else {
#pragma unroll
        for (int a = 0; a < R2; ++a) v[a] = rp[a];
      }
      fftR<R2, true>(v);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) {
        // psi' = sqrt(I) * psi / |psi + eps| = psi * rsqrt(|psi + eps|^2 / I)   (fpmMain.cpp:378-393), pixel x = k1 + R1*k2
        const float4 q4 = iv[k2 >> 2];
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
    row_block_sync();
    FPM_TICK(4);
    // ================= S5: rows stage A' (forward) =================
    if constexpr (NARROW && PIPE) {
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
        // S6, S7 and C2 read bbox columns only: six candidate outputs, the rest of the butterfly is dead code
        static_for<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int r = (k < 3) ? k : R1 - 6 + k;
          const int jw = (k < 3) ? q + R2 * r : q + R2 * r - N;
          if (jw >= p.xlo && jw <= p.xhi) rp[R2 * r] = v[r];                            // (q is warp-uniform)
        });
      });
    } else
    for (int q = rb_sub; q < R2; q += WPB) {
      float2* rp = fld + rb_row * PITCH + q;
      float2 v[R1];
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) v[k1] = rp[R2 * k1];
      fftR<R1, false>(v);
      if constexpr (NARROW) {
        // S6, S7 and C2 read bbox columns only: six candidate outputs, the rest of the butterfly is dead code
        static_for<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int r = (k < 3) ? k : R1 - 6 + k;
          const int jw = (k < 3) ? q + R2 * r : q + R2 * r - N;
          if (jw >= p.xlo && jw <= p.xhi) rp[R2 * r] = v[r];                            // (q is warp-uniform)
        });
      } else {
#pragma unroll
        for (int r = 0; r < R1; ++r) rp[R2 * r] = v[r];
      }
    }
    __syncthreads();
    FPM_TICK(5);
    // ================= S6: cols stage B' (forward) =================
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
    // ===== S7: cols stage A' (forward) -> Phi' in natural order; only bbox rows are stored (C2 reads nothing else) =====
    for (int g = tid, q = tq, jc = tr; g < R2 * NC; g += NT) {
      const int js = (p.xlo + jc) & (N - 1);
      float2 v[R1];
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) v[k1] = fld[(R2 * k1 + q) * PITCH + js];
      fftR<R1, false>(v);
      if constexpr (NARROW) {
        static_for<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          constexpr int r = (k < 3) ? k : R1 - 6 + k;
          const int i = R2 * r + q;
          const int iw = (k < 3) ? i : i - N;
          if (iw >= p.ylo && iw <= p.yhi) fld[i * PITCH + js] = v[r];
        });
      } else {
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          const int i = R2 * r + q;
          const int iw = (i < H) ? i : i - N;
          if (iw >= p.ylo && iw <= p.yhi) fld[i * PITCH + js] = v[r];
        }
      }
      step_nt(q, jc);
    }
    __syncthreads();
    FPM_TICK(7);
    // ===== C2: object update on the bbox (fpmMain.cpp:406-447), one element per thread and pass =====
    const int r0 = ys + H + p.ylo, r1 = ys + H + p.yhi, c0 = xs + H + p.xlo, c1 = xs + H + p.xhi;   // rectangle (inclusive)
    const int cr0 = r0 >> p.cs, ncr = (r1 >> p.cs) - cr0 + 1, cc0 = c0 >> 4, ncc = (c1 >> 4) - cc0 + 1;
    // the next LED's window was fetched by TMA before this update's writes: wait for it, then patch the overlap
    const int r0n = cr_b.y + H + p.ylo, c0n = (cr_b.x + H + p.xlo) & ~1;   // origin of the next window's TMA box
    FPM_TICK(12);
    if constexpr (Q_SMEM) { if (u > 0) { mbar_wait(&wbar, wphase); wphase ^= 1; } }
    FPM_TICK(13);
    {
      // Exact max|objF| bookkeeping (fpmMain.cpp:460,467): a grid U of per-cell maxima of |objFc|^2 (2^cs rows x 16
      // columns).  The cells this rectangle touches are rebuilt in Tm by atomicMax: new values of the rectangle's
      // pixels below, plus the pixels of those cells OUTSIDE the rectangle, which this update does not change --
      // their loads are issued first and consumed after the element loop (one L2 latency hidden behind it).  They read
      // L2 (ld.cg): earlier windows were written back by TMA stores, which do not update this SM's L1.
      const int wc0 = cc0 << 4, wcols = ncc << 4;                            // touched cells span these columns
      const int rt0 = cr0 << p.cs, nrt = ncr << p.cs;                        // ... and these rows
      const int wsh = 32 - __clz(wcols - 1);                                  // W rows are 2^wsh floats apart
      const int n_out = nrt << wsh;
      // fast path (one-row cells, everything on chip): per rectangle row the outside pixels are the left part of the
      // first cell (lanes 0..15) and the right part of the last cell (lanes 16..31): one warp-wide load per row
      constexpr int EPRE = 4;
      const bool fast_edges = Q_SMEM && (p.cs == 0);
      const int ecw = (lane < 16) ? lane : wcols - 32 + lane;
      const bool evalid = (lane < 16) ? (wc0 + ecw < c0) : (wc0 + ecw > c1);
      float2 eraw[EPRE];
      if (fast_edges) {
#pragma unroll
        for (int k = 0; k < EPRE; ++k) {
          const int it = warp + k * NW;
          eraw[k] = make_float2(0.f, 0.f);
          if (it < NR && evalid) eraw[k] = __ldcg(objFc + (size_t)(r0 + it) * L + wc0 + ecw);
        }
      }
      const float pm2 = warp_max(red[32 + (lane % NW)]);                       // (one load + REDUX instead of NW dependent loads)
      const float inv_pmax = rsqrt_fast(pm2);                                  // 1 / max|P|
      FPM_TICK(14);
      const int n = NR * NC;
      int ir = tq, jc = tr;
      constexpr int CU = Q_SMEM ? 2 : 6;                                       // elements in flight per thread
      for (int base = 0; base < n; base += CU * NT) {
        float2 Og[Q_SMEM ? 1 : CU];
        if constexpr (!Q_SMEM) {                                               // window in global memory: loads first
          int ir2 = ir, jc2 = jc;
#pragma unroll
          for (int k = 0; k < CU; ++k) {
            Og[k] = (base + k * NT + tid < n) ? wbase[(p.ylo + ir2) * L + p.xlo + jc2] : make_float2(0.f, 0.f);
            ir2 += qNT; jc2 += rNT;
            if (jc2 >= NC) { jc2 -= NC; ++ir2; }
          }
        }
#pragma unroll
        for (int k = 0; k < CU; ++k) {
          const int t = base + k * NT + tid;
          if (t < n) {
            const int iw = p.ylo + ir, jw = p.xlo + jc;
            const int i = iw & (N - 1), j = jw & (N - 1);
            float sup;
            if constexpr (Q_SMEM) sup = Sc[t]; else sup = __ldg(p.support + i * N + j);
            float2* gp = wbase + iw * L + jw;
            float2 O;
            if constexpr (Q_SMEM) O = Oc[ir * OCP + jc]; else O = Og[k];
            const float2 Pv = Pref(iw, jw);
            const float2 d = csub(fld[i * PITCH + j], cmul(O, Pv));           // dPhi = Phi' - Phi
            // dO = d * |P| conj(P) / (max|P| * ((|P|^2 + delta2) + i*kappa*delta2))
            const float pa2 = fmaf(Pv.x, Pv.x, Pv.y * Pv.y);
            const float2 num = cmulc(d, Pv);
            const float A = pa2 + p.delta2;
            const float sc = sqrt_fast(pa2) * inv_pmax * rcp_fast(fmaf(A, A, kd2 * kd2));
            const float2 On = make_float2(O.x + (num.x * A + num.y * kd2) * sc, O.y + (num.y * A - num.x * kd2) * sc);
            if constexpr (Q_SMEM) {
              Oc[ir * OCP + jc] = On;        // the whole box goes back to the spectrum with one TMA store after the barrier
              // forward the new value into the next window's box where the two overlap (the box was fetched earlier)
              const int rn = r0 + ir - r0n, cn = c0 + jc - c0n;
              if ((unsigned)rn < (unsigned)NR && (unsigned)cn < (unsigned)OCP) Ocn[rn * OCP + cn] = On;
            } else {
              *gp = On;
            }
            const float a2n = fmaf(On.x, On.x, On.y * On.y);
            if constexpr (Q_SMEM) W[((r0 + ir - rt0) << wsh) + (c0 + jc - wc0)] = a2n;
            // (measured without these atomics: 10.5 k instead of 11.4 k cycles for the 85 x 85 box -- they are not what binds)
            else atomicMax(&Tm[(((r0 + ir) >> p.cs) - cr0) * tmc + (((c0 + jc) >> 4) - cc0)], __float_as_uint(a2n));
            // Q = d * |O| conj(O) / ((|O|^2 + delta1) + i*kappa*delta1) * support   (fpmMain.cpp:459-472, O before the update)
            const float oa2 = fmaf(O.x, O.x, O.y * O.y);
            const float2 numq = cmulc(d, O);
            const float A1 = oa2 + p.delta1;
            const float sq = sqrt_fast(oa2) * sup * rcp_fast(fmaf(A1, A1, kd1 * kd1));
            Qref(iw, jw) = make_float2((numq.x * A1 + numq.y * kd1) * sq, (numq.y * A1 - numq.x * kd1) * sq);
          }
          ir += qNT; jc += rNT;
          if (jc >= NC) { jc -= NC; ++ir; }
        }
      }
      FPM_TICK(15);
      if (fast_edges) {
#pragma unroll
        for (int k = 0; k < EPRE; ++k) {
          const int it = warp + k * NW;
          asm volatile("" : "+f"(eraw[k].x), "+f"(eraw[k].y));     // the loads are consumed here, not where they were issued
          if (it < NR && evalid) W[(it << wsh) + ecw] = fmaf(eraw[k].x, eraw[k].x, eraw[k].y * eraw[k].y);
        }
        for (int it = warp + EPRE * NW; it < NR; it += NW)
          if (evalid) { const float2 o = __ldcg(objFc + (size_t)(r0 + it) * L + wc0 + ecw); W[(it << wsh) + ecw] = fmaf(o.x, o.x, o.y * o.y); }
      } else {
        // large rectangles / multi-row cells: the frame of the touched cells around the rectangle, as four strips
        // (above, below: full width; left, right: rectangle rows), four loads in flight per thread
        (void)n_out;
        const int rt1 = rt0 + nrt - 1;
        const int n_top = (r0 - rt0) * wcols, n_bot = (rt1 - r1) * wcols;
        const int wl = c0 - wc0, wrt = wc0 + wcols - 1 - c1;
        const int n_all = n_top + n_bot + NR * (wl + wrt);
        constexpr int DU = 4;
        for (int base = tid; base < n_all; base += DU * NT) {
          float2 o[DU];
          int rr[DU], cw[DU];
#pragma unroll
          for (int k = 0; k < DU; ++k) {
            const int t = base + k * NT;
            rr[k] = -1;
            if (t < n_all) {
              int r, c;
              if (t < n_top) { r = rt0 + t / wcols; c = wc0 + t % wcols; }
              else if (t < n_top + n_bot) { const int q = t - n_top; r = r1 + 1 + q / wcols; c = wc0 + q % wcols; }
              else if (t < n_top + n_bot + NR * wl) { const int q = t - n_top - n_bot; r = r0 + q / wl; c = wc0 + q % wl; }
              else { const int q = t - n_top - n_bot - NR * wl; r = r0 + q / wrt; c = c1 + 1 + q % wrt; }
              o[k] = __ldcg(objFc + (size_t)r * L + c);
              rr[k] = r - rt0; cw[k] = c - wc0;
            }
          }
#pragma unroll
          for (int k = 0; k < DU; ++k)
            if (rr[k] >= 0) {
              const float a2o = fmaf(o[k].x, o[k].x, o[k].y * o[k].y);
              if constexpr (Q_SMEM) W[(rr[k] << wsh) + cw[k]] = a2o;
              else atomicMax(&Tm[(rr[k] >> p.cs) * tmc + (cw[k] >> 4)], __float_as_uint(a2o));
            }
        }
      }
    }
    if constexpr (Q_SMEM) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // Oc writes -> visible to the TMA store
    __syncthreads();
    FPM_TICK(8);
    if constexpr (Q_SMEM) {
      if (tid == 0) tma_store_window(Ocur, &p.tmap, 2 * (c0 & ~1), r0, tile);          // updated window -> objFc
    }
    // ===== D: the touched cells take their rebuilt maxima; max|objF|^2 = max over the whole grid =====
    for (int t = tid; t < ncr * ncc; t += NT) {          // one thread per touched cell
      const int a = t / ncc, b = t - a * ncc;
      float m = 0.f;
      if constexpr (Q_SMEM) {
        const int wsh = 32 - __clz((ncc << 4) - 1);
        for (int rr = a << p.cs; rr < ((a + 1) << p.cs); ++rr) {
          const float4* w4 = reinterpret_cast<const float4*>(W + (rr << wsh) + (b << 4));
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float4 v = w4[q]; m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w)); }
        }
      } else {
        m = __uint_as_float(Tm[a * tmc + b]);
        Tm[a * tmc + b] = 0u;
      }
      U[(cr0 + a) * gc + cc0 + b] = m;
    }
    __syncthreads();
    {
      const float4* U4 = reinterpret_cast<const float4*>(U);
      const int n4 = (gr * gc) >> 2;                       // gc is a multiple of 4 (Nlarge multiple of 64)
      float m = 0.f;
      for (int base = 0; base < n4; base += 4 * NT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int t = base + k * NT + tid;
          if (t < n4) { const float4 q = U4[t]; m = fmaxf(fmaxf(m, fmaxf(q.x, q.y)), fmaxf(q.z, q.w)); }
        }
      }
      m = warp_max(m);
      if (lane == 0) red[warp] = m;
    }
    __syncthreads();
    FPM_TICK(9);
    {
      const float om2 = warp_max(red[lane % NW]);
      inv_objf_max = rsqrt_fast(om2);                // applied to P by the next S1 (or by the epilogue below)
    }
    if constexpr (!Q_SMEM) {
      // ===== E: pupil update P += Q / max|objF| (fpmMain.cpp:470-475) as its own coalesced pass; max|P|^2 =====
      float pm2 = 0.f;
      constexpr int EU = 5;                          // Q loads in flight per thread
      int ir = tq, jc = tr;
      for (int base = tid; base < NR * NC; base += EU * NT) {
        float2 Qv[EU];
        {
          int ir2 = ir, jc2 = jc;
#pragma unroll
          for (int k = 0; k < EU; ++k) {
            Qv[k] = (base + k * NT < NR * NC) ? Qref(p.ylo + ir2, p.xlo + jc2) : make_float2(0.f, 0.f);
            ir2 += qNT; jc2 += rNT;
            if (jc2 >= NC) { jc2 -= NC; ++ir2; }
          }
        }
#pragma unroll
        for (int k = 0; k < EU; ++k) {
          if (base + k * NT < NR * NC) {
            float2& pr = Pref(p.ylo + ir, p.xlo + jc);
            float2 v = pr;
            v.x = fmaf(Qv[k].x, inv_objf_max, v.x);
            v.y = fmaf(Qv[k].y, inv_objf_max, v.y);
            pr = v;
            pm2 = fmaxf(pm2, fmaf(v.x, v.x, v.y * v.y));
          }
          ir += qNT; jc += rNT;
          if (jc >= NC) { jc -= NC; ++ir; }
        }
      }
      pm2 = warp_max(pm2);
      if (lane == 0) red[32 + warp] = pm2;          // read by the next update's C2, several barriers from here
      __syncthreads();                               // P is complete before the next S1 reads it
    }
    cr_a = cr_b; cr_b = cr_c;
    FPM_TICK(10);
  }

#ifdef FPM_STAGE_TIMING
  if (tid == 0 && blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p.stage_clk[k] += tacc_[k];
  }
#endif
  if constexpr (Q_SMEM) { if (tid == 0) tma_store_wait_all(); }   // the last window stores have read their buffers
  if constexpr (Q_SMEM) {                           // the last update's pupil increment is still pending
    for (int t = tid; t < NR * NC; t += NT) {
      const int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
      const float2 Q = Qref(iw, jw);
      float2& pr = Pref(iw, jw);
      float2 v = pr;
      v.x = fmaf(Q.x, inv_objf_max, v.x);
      v.y = fmaf(Q.y, inv_objf_max, v.y);
      pr = v;
    }
  }
  __syncthreads();
  if constexpr (P_SMEM) {
    for (int t = tid; t < NR * NC; t += NT) {
      const int iw = p.ylo + t / NC, jw = p.xlo + t % NC;
      Pg[(iw & (N - 1)) * N + (jw & (N - 1))] = Pc[t];
    }
  }
}

// Uploaded uint16 images -> 1/I as float in the device layout (runs once per upload).  One CTA converts a band of
// RB rows of one image: 16-byte loads of the uint16 rows, the permutation of stack_offset<N> goes through shared
// memory (padded against bank conflicts), and the band leaves as contiguous runs of RUN floats (>= 128 bytes): both
// sides of the copy are coalesced.  grid = (images, N / RB).
template <int N> struct ConvertShape {
  static constexpr int RB = (N * N <= 8192) ? N : 8192 / N;                 // rows per band (32 KB of floats)
  static constexpr int NQ = RB / Shape<N>::R1;                              // values of y / R1 inside a band
  static constexpr int RUN = NQ * Shape<N>::R2;                             // contiguous output floats per (x % R1, y % R1)
  static_assert(RB % Shape<N>::R1 == 0 && (N % 8) == 0, "band shape");
};
template <int N>
__global__ void __launch_bounds__(256) stack_convert_kernel(float* stack, const uint16_t* raw, long long first_image) {
  using S = Shape<N>;
  using CS = ConvertShape<N>;
  constexpr int R1 = S::R1, R2 = S::R2, RB = CS::RB, RUN = CS::RUN, NB = RB * N;
  __shared__ float buf[NB + NB / 32 + 32];
  auto pad = [](int lo) { return lo + (lo >> 5) + ((lo >= NB / 2) ? 16 : 0); };
  const int band = blockIdx.y;                                              // rows [band*RB, +RB): y / R1 = band*NQ + yq
  const uint16_t* src = raw + ((size_t)(first_image + blockIdx.x) * N + (size_t)band * RB) * N;
  float* dst = stack + (size_t)(first_image + blockIdx.x) * N * N;
  constexpr int NLD = NB / 8 / 256;                                         // 16-byte loads per thread: all in flight together
  static_assert(NB % (8 * 256) == 0, "whole rounds of loads");
  uint4 qq[NLD];
#pragma unroll
  for (int r = 0; r < NLD; ++r) {
    const int t = threadIdx.x + r * 256;
    const int yy = t / (N / 8), c = t - yy * (N / 8);
    qq[r] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)yy * N) + c);   // (streaming: read once)
  }
#pragma unroll
  for (int r = 0; r < NLD; ++r) {                                           // 8 pixels of one row per thread and round
    const int t = threadIdx.x + r * 256;
    const int yy = t / (N / 8), c = t - yy * (N / 8);
    const uint4 q = qq[r];
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    const int ym = yy % R1, yq = yy / R1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int x = 8 * c + k;
      const unsigned v = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
      const int lo = ((x % R1) * R1 + ym) * RUN + yq * R2 + x / R1;
      buf[pad(lo)] = 1.0f / (float)v;                                       // 0 -> +inf: rsqrt(inf) = 0 = sqrt(0)
    }
  }
  __syncthreads();
  static_assert(RUN % 4 == 0 && NB % (4 * 256) == 0, "16-byte stores");
  for (int lo = 4 * threadIdx.x; lo < NB; lo += 4 * 256) {                  // four consecutive floats of a run per store
    const int r = lo / RUN, within = lo - r * RUN;
    const int A = r / R1, ym = r - A * R1;                                  // A = x % R1
    // stack_offset<N>(y, x) = ((x % R1) * N + R2 * (y % R1) + y / R1) * R2 + x / R1, with y / R1 = band*NQ + within / R2
    const int pl = pad(lo);                                                 // (a group of four never crosses a pad)
    const float4 v = make_float4(buf[pl], buf[pl + 1], buf[pl + 2], buf[pl + 3]);
    __stcs(reinterpret_cast<float4*>(dst + ((size_t)A * N + R2 * ym + band * CS::NQ) * R2 + within), v);
  }
}

}  // namespace fpm
