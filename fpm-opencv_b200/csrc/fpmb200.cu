// fpmb200.cu -- C-ABI layer (include/fpmb200.h) over the sm_100a kernels of fpm_kernels.cuh.
// Replaces the cv::UMat / cvComplex op sequence of runFPM() (fpmMain.cpp:274-498).
// No CPU fallback: every entry point fails when no CUDA device can run the kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/fpmb200.h"
#include "fpm_kernels.cuh"

using namespace fpm;

// threads per CTA of fpm_update_pruned_kernel: 20 warps (5 per scheduler, 96 registers).  The stages are a few rounds
// of long dependent butterfly chains, i.e. latency-bound: measured at Np = 200 / 148 tiles, 512 threads 4.44 M
// updates/s, 544: 4.29 M, 640: 4.62 M, 704 / 768 / 896 (80 / 80 / 72 registers, spills): 4.16 / 4.25 / 4.12 M.
#ifndef FPM_PRUNED_NT
#define FPM_PRUNED_NT 640
#endif
static constexpr int PRUNED_NT = FPM_PRUNED_NT;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(FPMB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct fpmb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_tiles = 0, N = 0, L = 0, n_leds = 0;
  float delta1 = 5.f, delta2 = 10.f, eps = 1e-10f, kappa = 1.f;   // fpmMain.cpp:567-568, fpmMain.h:99
  float2* objFc = nullptr;     // [n_tiles][L][L] centred
  float2* objCrop = nullptr;   // [n_tiles][L][L]
  float2* pupil = nullptr;     // [n_tiles][N][N]
  float* stack = nullptr;      // [n_tiles][n_leds][N*N] 1/I in the kernel's layout
  uint16_t* raw = nullptr;     // [n_tiles][n_leds][N][N] upload staging (as given by the host)
  float* support = nullptr;    // [N][N]
  short2* crop = nullptr;      // [n_leds]
  float2* twN = nullptr;       // [N]
  float2* twL = nullptr;       // [L]
  float2* field_gmem = nullptr;
  float2* scratch = nullptr;   // staging: max(L*L, init batch * N*N)
  size_t scratch_elems = 0;
  int ylo = 0, yhi = -1, xlo = 0, xhi = -1;
  bool have_leds = false, have_support = false, have_stack = false;
  // kernel variant
  bool field_smem = false, p_smem = false, q_smem = false;
  bool narrow = false;         // bbox within +-(3*R2-1): the pruned-butterfly instantiation of fpm_update_kernel (N = 128)
  bool phased = false;         // run by fpm_update_phased_kernel (three phases per update; fpm_update_phased.cuh)
  int cs = 0;                  // log2 rows per max-cell
  size_t smem_bytes = 0;
  int max_smem_optin = 0, sm_count = 0;
  long long launches = 0;
  long long* stage_clk = nullptr;
  CUtensorMap tmap;            // objFc as (2*L, L, n_tiles) floats with a (2*ocp, NR, 1) box
  bool have_tmap = false;
  int ocp = 0;
  // general (unfused) path for tile sizes other than 64/128/256
  bool general = false;
  bool gfused = false;         // general sizes whose field fits shared memory twice: fpm_update_general_kernel
  bool gpruned = false;        // general sizes with a compiled R1 x R2 plan whose pupil box fits: fpm_update_pruned_kernel
  int stack_r1 = 0;            // stack layout of the general path (stack_pos_offset): 0 natural, else R1 of the pruned plan
  int cb = 0;                  // columns per batch of fpm_update_pruned_kernel
  float2* gfield = nullptr;    // [n_tiles][N][N]
  float2* gq = nullptr;        // [n_tiles][N][N]
  float* gcells = nullptr;     // [n_tiles][cgr][cgc]
  float* gscal = nullptr;      // [n_tiles][4]
  int cgr = 0, cgc = 0;
  // one pass over all LEDs of the general path (11 launches per LED) captured as a CUDA graph, per tile range
  struct IterGraph { int first, n; cudaGraphExec_t exec; long long nodes; };
  std::vector<IterGraph> iter_graphs;
  // full-FOV helpers (csrc/fpm_fov.cuh)
  int2* origins = nullptr;     // [n_tiles] ROI origin of every tile in the camera frame
  bool have_origins = false;
  uint16_t* frame_dev = nullptr;
  size_t frame_elems = 0;
  int* bg_dev = nullptr;       // [n_leds] background value subtracted from each LED frame
  int origin_max_x = 0, origin_max_y = 0, origin_min_y = 0;
  float* mosaic_dev = nullptr;
  size_t mosaic_elems = 0;
  cudaEvent_t events[64] = {};  // stream markers of fpmb200_event_record
  int* sched_dev = nullptr;    // tile lists of the balanced passes of fpmb200_run ...
  size_t sched_cap = 0;
  int sched_first = -1, sched_n = 0, sched_iters = 0;   // ... cached for this (first, n, iters)
  std::vector<int> sched_counts;
  float* ucache = nullptr;     // [n_tiles][L][L/16]: cell maxima of fpm_update_phased_kernel between balanced passes
  int cluster_req = 0;         // CTAs per tile asked for (0 = choose)
  int cluster = 1;             // CTAs per tile in use (1 = fpm_update_kernel, >1 = fpm_update_cluster_kernel)
  int cpc = 0;                 // bbox columns per CTA of the cluster kernel
  char variant[200] = "unallocated";
};

static int select_variant(fpmb200_ctx* c);
static void factorize(int n, fpm::LineFFTParams& p);
static int general_fused_plan(int N);
static int pruned_plan(int N);

extern "C" const char* fpmb200_last_error(void) { return g_err.c_str(); }
extern "C" int fpmb200_abi_version(void) { return 1; }

extern "C" int fpmb200_create(int device, fpmb200_ctx** out) {
  if (!out) return fail(FPMB200_ERR_ARG, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(FPMB200_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(FPMB200_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(FPMB200_ERR_CUDA, "device %d is sm_%d%d; libfpmb200 is built for sm_100a only", device, prop.major, prop.minor);
  fpmb200_ctx* c = new fpmb200_ctx();
  c->device = device;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->sm_count = prop.multiProcessorCount;
  if (cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); e2 != cudaSuccess) {
    delete c;
    return fail(FPMB200_ERR_CUDA, "cudaStreamCreateWithFlags failed: %s", cudaGetErrorString(e2));
  }
  *out = c;
  return FPMB200_OK;
}

// Every copy / memset goes through the context's own (non-blocking) stream: a plain cudaMemcpy from
// pageable memory may return before its DMA has landed, and the legacy stream is not ordered with
// a cudaStreamNonBlocking stream -- kernels launched next could read stale data.
static cudaError_t copy_sync(fpmb200_ctx* c, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, c->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(c->stream);
}

static void drop_graphs(fpmb200_ctx* c);

static void free_tiles(fpmb200_ctx* c) {
  drop_graphs(c);
  cudaFree(c->objFc); cudaFree(c->objCrop); cudaFree(c->pupil); cudaFree(c->stack); cudaFree(c->raw); cudaFree(c->support);
  cudaFree(c->sched_dev); c->sched_dev = nullptr; c->sched_cap = 0; c->sched_first = -1;
  cudaFree(c->ucache); c->ucache = nullptr;
  cudaFree(c->crop); cudaFree(c->twN); cudaFree(c->twL); cudaFree(c->field_gmem); cudaFree(c->scratch);
  cudaFree(c->gfield); cudaFree(c->gq); cudaFree(c->gcells); cudaFree(c->gscal);
  cudaFree(c->origins); cudaFree(c->frame_dev); cudaFree(c->bg_dev); cudaFree(c->mosaic_dev);
  c->origins = nullptr; c->frame_dev = nullptr; c->bg_dev = nullptr; c->mosaic_dev = nullptr;
  c->have_origins = false; c->frame_elems = c->mosaic_elems = 0;
  c->gfield = c->gq = nullptr; c->gcells = c->gscal = nullptr; c->general = false;
  c->objFc = c->objCrop = c->pupil = c->twN = c->twL = c->field_gmem = c->scratch = nullptr;
  c->stack = nullptr; c->raw = nullptr; c->support = nullptr; c->crop = nullptr;
  c->have_leds = c->have_support = c->have_stack = false;
  c->n_tiles = 0;
}

extern "C" void fpmb200_destroy(fpmb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_tiles(c);
  for (cudaEvent_t e : c->events) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static int upload_twiddles(fpmb200_ctx* c, float2* dst, int n) {
  std::vector<float2> h(n);
  for (int k = 0; k < n; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)n;
    double cr = cos(a), ci = sin(a);
    if (fabs(cr) < 1e-12) cr = 0;      // exact zeros at the quarter turns
    if (fabs(ci) < 1e-12) ci = 0;
    h[k] = make_float2((float)cr, (float)ci);
  }
  CK(copy_sync(c, dst, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
  return FPMB200_OK;
}

static int tiles_alloc_impl(fpmb200_ctx* c, int n_tiles, int Np, int Nlarge, int n_leds) {
  const size_t LL = (size_t)Nlarge * Nlarge, NN = (size_t)Np * Np;
  CK(cudaMalloc(&c->objFc, sizeof(float2) * LL * n_tiles));
  CK(cudaMalloc(&c->objCrop, sizeof(float2) * LL * n_tiles));
  CK(cudaMalloc(&c->pupil, sizeof(float2) * NN * n_tiles));
  CK(cudaMalloc(&c->stack, sizeof(float) * NN * n_leds * n_tiles));
  CK(cudaMalloc(&c->raw, sizeof(uint16_t) * NN * n_leds * n_tiles));
  CK(cudaMalloc(&c->support, sizeof(float) * NN));
  CK(cudaMalloc(&c->crop, sizeof(short2) * n_leds));
  CK(cudaMalloc(&c->twN, sizeof(float2) * Np));
  CK(cudaMalloc(&c->twL, sizeof(float2) * Nlarge));
  {
    // staging for the spectrum seeds of as many tiles per batch as 256 MB hold (at least 64), and one spectrum
    size_t nb = std::min<size_t>((size_t)n_tiles, ((size_t)256 << 20) / (sizeof(float2) * NN));
    nb = std::max<size_t>(nb, 64);
    c->scratch_elems = std::max(LL, NN * nb);
  }
  CK(cudaMalloc(&c->scratch, sizeof(float2) * c->scratch_elems));
  CK(cudaMemsetAsync(c->objFc, 0, sizeof(float2) * LL * n_tiles, c->stream));
  CK(cudaMemsetAsync(c->objCrop, 0, sizeof(float2) * LL * n_tiles, c->stream));
  CK(cudaMemsetAsync(c->pupil, 0, sizeof(float2) * NN * n_tiles, c->stream));
  CK(cudaMemsetAsync(c->scratch, 0, sizeof(float2) * c->scratch_elems, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int rc;
  if ((rc = upload_twiddles(c, c->twN, Np)) != FPMB200_OK) return rc;
  if ((rc = upload_twiddles(c, c->twL, Nlarge)) != FPMB200_OK) return rc;
  return FPMB200_OK;
}

extern "C" int fpmb200_tiles_alloc(fpmb200_ctx* c, int n_tiles, int Np, int Nlarge, int n_leds) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  if (n_tiles <= 0 || n_leds <= 0) return fail(FPMB200_ERR_ARG, "n_tiles and n_leds must be positive");
  // tiles index gridDim.y of the batched helper kernels (initialisation, line FFTs, general path)
  if (n_tiles > 65535) return fail(FPMB200_ERR_ARG, "n_tiles=%d exceeds 65535 tiles per context (use several contexts)", n_tiles);
  auto smooth235 = [](int v) { for (int f : {2, 3, 5}) while (v % f == 0) v /= f; return v == 1; };
  if (Np < 8 || Np > 1024 || (Np & 1) || !smooth235(Np))
    return fail(FPMB200_ERR_ARG, "Np=%d unsupported: even, 8..1024, prime factors 2,3,5 only (fused kernels: 64, 128, 256)", Np);
  if (Nlarge < Np || Nlarge > 3584 || (Nlarge & 1) || !smooth235(Nlarge))
    return fail(FPMB200_ERR_ARG, "Nlarge=%d must be even, in [Np, 3584], with prime factors 2,3,5 only", Nlarge);
  // the fused kernels need power-of-two tiles and 64-aligned spectra; everything else takes the general path
  // (FPMB200_FORCE_GENERAL=1: developer switch, power-of-two tiles through the general-path kernels)
  const char* fg = getenv("FPMB200_FORCE_GENERAL");
  const bool general = !((Np == 64 || Np == 128 || Np == 256) && Nlarge % 64 == 0) || (fg && fg[0] == '1');
  CK(cudaSetDevice(c->device));
  free_tiles(c);
  c->N = Np; c->L = Nlarge; c->n_leds = n_leds;
  c->general = general;
  c->stack_r1 = general ? pruned_plan(Np) / 100 : 0;       // the stack layout depends on Np only, never on the support
  // n_tiles is published only when every buffer exists: a failed allocation leaves an unallocated context
  const int rc = tiles_alloc_impl(c, n_tiles, Np, Nlarge, n_leds);
  if (rc != FPMB200_OK) {
    const std::string msg = g_err;
    cudaGetLastError();
    free_tiles(c);
    g_err = msg;
    return rc;
  }
  c->n_tiles = n_tiles;
  snprintf(c->variant, sizeof c->variant, "allocated (support not uploaded yet)");
  return FPMB200_OK;
}

extern "C" int fpmb200_set_params(fpmb200_ctx* c, float delta1, float delta2, float eps, int literal_scalar) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  c->delta1 = delta1; c->delta2 = delta2; c->eps = eps; c->kappa = literal_scalar ? 1.f : 0.f;
  drop_graphs(c);
  return FPMB200_OK;
}

extern "C" int fpmb200_set_cluster(fpmb200_ctx* c, int ctas_per_tile) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  if (ctas_per_tile < 0 || ctas_per_tile > 8) return fail(FPMB200_ERR_ARG, "ctas_per_tile must be 0 (choose), 1, 2, 4 or 8");
  c->cluster_req = ctas_per_tile;
  return c->have_support ? select_variant(c) : FPMB200_OK;
}

extern "C" int fpmb200_upload_leds(fpmb200_ctx* c, const int16_t* cx, const int16_t* cy, int n_leds) {
  if (!c || !cx || !cy) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->n_tiles) return fail(FPMB200_ERR_STATE, "fpmb200_tiles_alloc first");
  if (n_leds != c->n_leds) return fail(FPMB200_ERR_ARG, "n_leds=%d differs from the allocation (%d)", n_leds, c->n_leds);
  std::vector<short2> h(n_leds);
  for (int k = 0; k < n_leds; ++k) {
    if (cx[k] < 0 || cy[k] < 0 || cx[k] + c->N > c->L || cy[k] + c->N > c->L)
      return fail(FPMB200_ERR_ARG, "slot %d: crop origin (%d,%d) leaves the %dx%d spectrum", k, cx[k], cy[k], c->L, c->L);
    h[k] = make_short2(cx[k], cy[k]);
  }
  CK(cudaSetDevice(c->device));
  CK(copy_sync(c, c->crop, h.data(), sizeof(short2) * n_leds, cudaMemcpyHostToDevice));
  c->have_leds = true;
  return FPMB200_OK;
}

static size_t update_smem_bytes(const fpmb200_ctx* c, bool field_smem, bool p_smem, bool q_smem, int cs) {
  const int N = c->N, PITCH = N + 1;
  const size_t bb = sizeof(float2) * (size_t)(c->yhi - c->ylo + 1) * (c->xhi - c->xlo + 1);
  size_t b = 0;
  if (field_smem) b += (sizeof(float2) * N * PITCH + 15) / 16 * 16;
  b += sizeof(float4) * N * 2 + sizeof(float) * 64;                                   // twA, twB (twmul4 operands), red
  if (p_smem) b += bb;
  if (q_smem) {                                                                        // Qc + two TMA window buffers
    const int NRb = c->yhi - c->ylo + 1, ocp = ((c->xhi - c->xlo + 1) + 2) & ~1;
    b += bb + 128 + 2 * ((sizeof(float2) * (size_t)NRb * ocp + 127) / 128 * 128);
  }
  b += sizeof(float) * (size_t)(c->L >> cs) * (c->L >> 4);                            // U
  if (q_smem) b += bb / 2;                                                             // Sc (support on the bbox)
  b += sizeof(unsigned) * (size_t)(((c->yhi - c->ylo + 1) >> cs) + 2) * (((c->xhi - c->xlo + 1) >> 4) + 2);   // Tm
  if (q_smem) {                                                                        // W: every pixel of the touched cells
    const int tmr = ((c->yhi - c->ylo + 1) >> cs) + 2, tmc = ((c->xhi - c->xlo + 1) >> 4) + 2;
    int wsh = 0; while ((1 << wsh) < (tmc << 4)) ++wsh;
    b += 16 + sizeof(float) * (size_t)((tmr << cs) << wsh);
  }
  b += 16;                                                                             // alignment of U
  return b;
}

extern "C" int fpmb200_upload_pupil_support(fpmb200_ctx* c, const float* mask) {
  if (!c || !mask) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->n_tiles) return fail(FPMB200_ERR_STATE, "fpmb200_tiles_alloc first");
  const int N = c->N, H = N / 2;
  int ylo = H, yhi = -H - 1, xlo = H, xhi = -H - 1;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j)
      if (mask[i * N + j] != 0.f) {
        int iw = i < H ? i : i - N, jw = j < H ? j : j - N;
        if (iw < ylo) ylo = iw; if (iw > yhi) yhi = iw;
        if (jw < xlo) xlo = jw; if (jw > xhi) xhi = jw;
      }
  if (yhi < ylo) return fail(FPMB200_ERR_ARG, "pupil support is empty");
  c->ylo = ylo; c->yhi = yhi; c->xlo = xlo; c->xhi = xhi;
  CK(cudaSetDevice(c->device));
  CK(copy_sync(c, c->support, mask, sizeof(float) * N * N, cudaMemcpyHostToDevice));
  c->have_support = true;
  return select_variant(c);
}

// the cluster kernel instance for this box: six-sample stage-A butterflies when a 128 x 128 tile's box lies within +-23
typedef void (*cluster_kernel_t)(const UpdateParams);
template <int N>
static bool cluster_six(const fpmb200_ctx* c) {
  const int lim = 3 * Shape<N>::R2 - 1;
  return N == 128 && c->ylo >= -lim && c->yhi <= lim && c->xlo >= -lim && c->xhi <= lim;
}
template <int N, int C>
static cluster_kernel_t cluster_kernel_for(const fpmb200_ctx* c) {
  if constexpr (N == 128) {
    if (cluster_six<N>(c)) return fpm_update_cluster_kernel<N, C, 512, true>;
  }
  return fpm_update_cluster_kernel<N, C, 512, false>;
}

template <int N, int C>
static bool cluster_fits(fpmb200_ctx* c, int* cpc_out, int* cs_out, size_t* bytes_out, int* max_clusters = nullptr) {
  const int NR = c->yhi - c->ylo + 1, NC = c->xhi - c->xlo + 1;
  const int cpc = (NC + C - 1) / C;
  if (cpc > 32) return false;                                  // one lane per column in the column passes
  for (int cs = 0; cs <= 4; ++cs) {
    const ClusterLayout<N, C> lay(NR, NC, cpc, c->L, cs, cluster_six<N>(c));
    if (sizeof(float) * (size_t)lay.gro * (c->L >> 4) > 8 * 1024) continue;      // scanned once per update
    if (sizeof(float) * (size_t)NR * cpc > sizeof(float2) * (size_t)(N / C) * (N + 1)) continue;   // W aliases the row slab
    if (lay.total > (size_t)c->max_smem_optin) continue;
    auto k = cluster_kernel_for<N, C>(c);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total) != cudaSuccess) { cudaGetLastError(); continue; }
    if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(C, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = lay.total;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, k, &cfg) != cudaSuccess || n_clusters < 1) { cudaGetLastError(); continue; }
    *cpc_out = cpc; *cs_out = cs; *bytes_out = lay.total;
    if (max_clusters) *max_clusters = n_clusters;
    return true;
  }
  return false;
}

// ---- choose the kernel variant for this (N, L, bbox): CTAs per tile and what lives in shared memory ----
static int select_variant(fpmb200_ctx* c) {
  drop_graphs(c);              // captured passes bake in the bounding box of the support and the kernel variant
  const int N = c->N;
  const int ylo = c->ylo, yhi = c->yhi, xlo = c->xlo, xhi = c->xhi;
  CK(cudaSetDevice(c->device));
  if (c->general) {
    if (c->cluster_req > 1) return fail(FPMB200_ERR_ARG, "cluster kernels exist for Np = 128 and 256 only (Np=%d)", N);
    c->cluster = 1;
    c->cgr = c->cgc = (c->L + 15) / 16;
    // one CTA per tile with the field in shared memory when two copies of it fit (Np 90, 100 of the shipped JSONs);
    // FPMB200_GENERAL_UNFUSED=1 keeps the per-step kernels, FPMB200_GENERAL_PLAN=0 the run-time radices (developer A/B)
    LineFFTParams fp;
    factorize(N, fp);
    {
      const char* e = getenv("FPMB200_GENERAL_UNFUSED");
      c->gfused = fp.nrad <= 8 && general_fused_smem_bytes(N, c->cgr, c->cgc) <= (size_t)c->max_smem_optin && !(e && e[0] == '1') &&
                  !c->stack_r1;          // (a position-major stack belongs to fpm_update_pruned_kernel)
    }
    if (c->gfused) {
      c->smem_bytes = general_fused_smem_bytes(N, c->cgr, c->cgc);
      const int plan = general_fused_plan(N);
      // (measured: the pupil kept in shared memory, compact over the box, changes nothing -- 9.63 against 9.53 M updates/s
      //  at Np = 90, 7.41 against 7.43 M at Np = 100: the batched L2 loads of P were already hidden)
      char radices[64];
      if (plan) snprintf(radices, sizeof radices, "radix %d x %d in registers, pruned to the pupil box", plan / 100, plan % 100);
      else snprintf(radices, sizeof radices, "run-time radices, %d stages", fp.nrad);
      snprintf(c->variant, sizeof c->variant,
               "general path, fused: fpm_update_general_kernel (one CTA per tile, field in shared memory, Stockham %s) "
               "Np=%d Nlarge=%d maxcell=16x16 smem=%zuB", radices, N, c->L, c->smem_bytes);
      return FPMB200_OK;
    }
    // Larger tiles with a compiled R1 x R2 plan (the shipped 200 = 20 x 10): the field is never materialised -- box rows
    // in shared memory, columns in batches (fpm_pruned_fused.cuh); FPMB200_GENERAL_UNFUSED=1 keeps the per-step kernels
    c->gpruned = false;
    if (const int plan = pruned_plan(N)) {
      const char* e = getenv("FPMB200_GENERAL_UNFUSED");
      const int R1 = plan / 100, R2 = plan % 100, NT = PRUNED_NT;
      const int nrb = c->yhi - c->ylo + 1;
      const size_t base = pruned_fused_smem_bytes(N, nrb, 0, c->cgr, c->cgc);
      const long long room = (long long)c->max_smem_optin - (long long)base;
      const int cbmax = room > 0 ? (int)std::min<long long>(room / ((long long)sizeof(float2) * N), N) : 0;
      if (cbmax >= 8 && !(e && e[0] == '1')) {
        // columns per batch: whole rounds of NT work items in the three stages of a batch, few batches
        auto rounds = [&](int items) { return (items + NT - 1) / NT; };
        const double c1 = R1 * log2((double)R1), c2 = R2 * log2((double)R2);
        double best = 1e300;
        for (int cb = 8; cb <= cbmax; ++cb) {
          double cost = 0;
          for (int c0 = 0; c0 < N; c0 += cb) {
            const int n = std::min(cb, N - c0);
            cost += 2 * rounds(n * R2) * c1 + 2 * rounds(n * R1) * c2 + 0.5 * (c1 + c2);     // + barrier / ramp cost per batch
          }
          if (cost <= best) { best = cost; c->cb = cb; }
        }
        c->gpruned = true;
        {
          const int lim = 3 * R2;
          const char* en = getenv("FPMB200_PRUNED_NARROW");
          c->narrow = R1 == 20 && ylo >= -lim && yhi <= lim - 1 && xlo >= -lim && xhi <= lim - 1 && !(en && en[0] == '0');
        }
        c->smem_bytes = pruned_fused_smem_bytes(N, nrb, c->cb, c->cgr, c->cgc);
        snprintf(c->variant, sizeof c->variant,
                 "general path, fused: fpm_update_pruned_kernel (one CTA per tile, box rows + %d-column batches in shared "
                 "memory, radix %d x %d in place%s) Np=%d Nlarge=%d bbox=[%d..%d]x[%d..%d] maxcell=16x16 smem=%zuB",
                 c->cb, R1, R2, c->narrow ? ", six-sample stage-A butterflies" : "", N, c->L, ylo, yhi, xlo, xhi, c->smem_bytes);
        return FPMB200_OK;
      }
    }
    if (!c->gfield) {
      const size_t NN = (size_t)N * N;
      CK(cudaMalloc(&c->gfield, sizeof(float2) * NN * c->n_tiles));
      CK(cudaMalloc(&c->gq, sizeof(float2) * NN * c->n_tiles));
      CK(cudaMalloc(&c->gcells, sizeof(float) * (size_t)c->cgr * c->cgc * c->n_tiles));
      CK(cudaMalloc(&c->gscal, sizeof(float) * 4 * c->n_tiles));
      CK(cudaMemsetAsync(c->gq, 0, sizeof(float2) * NN * c->n_tiles, c->stream));
      CK(cudaMemsetAsync(c->gscal, 0, sizeof(float) * 4 * c->n_tiles, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    snprintf(c->variant, sizeof c->variant,
             "general path (unfused: line_fft_kernel + 6 elementwise kernels per update) Np=%d Nlarge=%d maxcell=16x16", N, c->L);
    return FPMB200_OK;
  }
  // A 256x256 field does not fit one SM: spread the tile over a cluster of 8 CTAs (field, pupil and pupil increment on
  // chip, transposes through DSMEM).  128x128 tiles use a cluster only on request (fpmb200_set_cluster).
  c->cluster = 1;
  {
    // default: 256x256 tiles always (the field does not fit one SM otherwise); 128x128 tiles when there are so few
    // tiles that SMs would idle anyway (lower latency per update): four CTAs per tile up to a quarter of the SMs
    // (6.5 us per update against 10.1 us on one CTA), two CTAs per tile up to half of them when the box is narrow (window
    // slice on chip: 9.0 us) -- e.g. the 40 tiles a GPU holds of a 2560 x 2160 frame spread over eight GPUs
    const int want = c->cluster_req ? c->cluster_req
                   : (N == 256 ? 8 : (N == 128 && c->n_tiles * 4 <= c->sm_count) ? 4
                   : (N == 128 && c->n_tiles * 2 <= c->sm_count && cluster_six<128>(c)) ? 2 : 1);
    bool ok = false;
    int max_clusters = 0;
#if !defined(FPM_DEV_FAST) || defined(FPM_DEV_CLUSTER)
    if (N == 256 && want == 8) ok = cluster_fits<256, 8>(c, &c->cpc, &c->cs, &c->smem_bytes);
    else if (N == 128 && want == 4) {
      ok = cluster_fits<128, 4>(c, &c->cpc, &c->cs, &c->smem_bytes, &max_clusters);
      // the library's own choice only pays when every tile's cluster is resident at once (GPC boundaries strand SMs)
      if (ok && !c->cluster_req && c->n_tiles > max_clusters) ok = false;
    }
    else if (N == 128 && want == 2) {
      ok = cluster_fits<128, 2>(c, &c->cpc, &c->cs, &c->smem_bytes, &max_clusters);
      if (ok && !c->cluster_req && c->n_tiles > max_clusters) ok = false;
    }
    else
#endif
    if (want != 1 && c->cluster_req)
      return fail(FPMB200_ERR_ARG, "%d CTAs per tile is not available for Np=%d (256: 8; 128: 2 or 4; any: 1)", want, N);
    if (ok) {
      c->cluster = want;
      snprintf(c->variant, sizeof c->variant,
               "fpm_update_cluster_kernel<N=%d,cluster=%d%s%s> bbox=[%d..%d]x[%d..%d] cols/CTA=%d maxcell=%dx16 smem=%zuB", N, want,
               (N == 128 && cluster_six<128>(c)) ? ",window=smem,pruned radix-16" : "", "",
               ylo, yhi, xlo, xhi, c->cpc, 1 << c->cs, c->smem_bytes);
      return FPMB200_OK;
    }
    if (c->cluster_req > 1) return fail(FPMB200_ERR_ARG, "the %d-CTA cluster kernel does not fit this geometry (Np=%d, Nlarge=%d)", want, N, c->L);
  }
  const size_t cap = (size_t)c->max_smem_optin;
  c->field_smem = (N <= 128);
  // what lives in shared memory: prefer pupil + pupil-increment on chip with the finest max-cells that fit
  bool found = false;
  for (int pq = 0; pq < 3 && !found; ++pq) {
    const bool ps = pq < 2, qs = pq < 1 && (2 * (((xhi - xlo + 1) + 2) & ~1) <= 256);   // TMA box <= 256 elements per dim
    for (int cs = 0; cs <= 4 && !found; ++cs) {
      if (sizeof(float) * (size_t)(c->L >> cs) * (c->L >> 4) > 48 * 1024) continue;
      if (update_smem_bytes(c, c->field_smem, ps, qs, cs) <= cap) {
        c->p_smem = ps; c->q_smem = qs; c->cs = cs; found = true;
      }
    }
  }
  if (!found) return fail(FPMB200_ERR_ARG, "update kernel does not fit %zu B of shared memory (Np=%d, Nlarge=%d)", cap, N, c->L);
  const int cs = c->cs;
  {
    const int lim = 3 * (N / 16) - 1;        // N = 128: R1 = 16, R2 = 8
    c->narrow = (N == 128) && c->p_smem && c->q_smem && ylo >= -lim && yhi <= lim && xlo >= -lim && xhi <= lim;
  }
  c->smem_bytes = update_smem_bytes(c, c->field_smem, c->p_smem, c->q_smem, cs);
  {
    // everything on chip with one-row max-cells, 64 x 64 tiles or narrow pupils on 128 x 128 tiles: the three-phase
    // kernel (FPMB200_UPDATE_V1=1 keeps fpm_update_kernel, developer A/B and the tests of that kernel)
    const char* e = getenv("FPMB200_UPDATE_V1");
    const int NRb = yhi - ylo + 1, NCb = xhi - xlo + 1, ocp = (NCb + 2) & ~1;
    int noff[PhasedShape<128>::NOFF];
    bool ok = c->field_smem && c->p_smem && c->q_smem && cs == 0 && (c->L % 64) == 0 && !(e && e[0] == '1');
    if (ok && N == 128) { ok = c->narrow && PhasedShape<128>::box_ok(ylo, yhi, xlo, xhi); PhasedShape<128>::layout(NRb, NCb, ocp, c->L, noff); }
    else if (ok && N == 64) { ok = PhasedShape<64>::box_ok(ylo, yhi, xlo, xhi); PhasedShape<64>::layout(NRb, NCb, ocp, c->L, noff); }
    else ok = false;
    c->phased = ok && (size_t)noff[PhasedShape<128>::TOTAL] <= cap;
    if (c->phased) c->smem_bytes = (size_t)noff[PhasedShape<128>::TOTAL];
  }
  if (!c->field_smem && !c->field_gmem)
    CK(cudaMalloc(&c->field_gmem, sizeof(float2) * (size_t)N * (N + 1) * c->n_tiles));
  c->ocp = ((xhi - xlo + 1) + 2) & ~1;       // even, with room for the 16-byte alignment of TMA box starts
  c->have_tmap = false;
  if (c->q_smem) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(FPMB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t gdim[3] = {(cuuint64_t)2 * c->L, (cuuint64_t)c->L, (cuuint64_t)c->n_tiles};
    const cuuint64_t gstr[2] = {(cuuint64_t)2 * c->L * sizeof(float), (cuuint64_t)2 * c->L * c->L * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * c->ocp), (cuuint32_t)(yhi - ylo + 1), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->objFc, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FPMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    c->have_tmap = true;
  }
  snprintf(c->variant, sizeof c->variant,
           "%s<N=%d,field=%s,pupil=%s,dP=%s%s> bbox=[%d..%d]x[%d..%d] maxcell=%dx16 smem=%zuB",
           c->phased ? "fpm_update_phased_kernel" : "fpm_update_kernel", N,
           c->field_smem ? "smem" : "gmem", c->p_smem ? "smem" : "gmem", c->q_smem ? "smem" : "field",
           c->narrow ? ",pruned radix-16" : "", ylo, yhi, xlo, xhi,
           1 << cs, c->smem_bytes);
  return FPMB200_OK;
}

static int check_range(fpmb200_ctx* c, int first, int n) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  if (!c->n_tiles) return fail(FPMB200_ERR_STATE, "fpmb200_tiles_alloc first");
  if (first < 0 || n <= 0 || first + n > c->n_tiles)
    return fail(FPMB200_ERR_ARG, "tile range [%d,%d) outside [0,%d)", first, first + n, c->n_tiles);
  return FPMB200_OK;
}

extern "C" int fpmb200_upload_stack(fpmb200_ctx* c, int first, int n, const uint16_t* stack, void* stream) {
  int rc = check_range(c, first, n);
  if (rc) return rc;
  if (!stack) return fail(FPMB200_ERR_ARG, "stack is NULL");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t per = (size_t)c->N * c->N * c->n_leds;
  CK(cudaMemcpyAsync(c->raw + per * first, stack, sizeof(uint16_t) * per * n, cudaMemcpyHostToDevice, st));
  // once per upload: uint16 -> 1/I (float) in the layout the update kernel streams (stack_offset<N>)
  const long long first_img = (long long)first * c->n_leds;
  const int n_img = n * c->n_leds;
  if (c->general) {
    const long long n_el = (long long)n_img * c->N * c->N;
    stack_convert_general<<<(int)((n_el + 256 * 8 - 1) / (256 * 8)), 256, 0, st>>>(c->stack, c->raw, first_img * c->N * c->N, n_el,
                                                                                    c->N, c->stack_r1);
  } else switch (c->N) {
    case 64: stack_convert_kernel<64><<<dim3(n_img, 64 / ConvertShape<64>::RB), 256, 0, st>>>(c->stack, c->raw, first_img); break;
    case 128: stack_convert_kernel<128><<<dim3(n_img, 128 / ConvertShape<128>::RB), 256, 0, st>>>(c->stack, c->raw, first_img); break;
    case 256: stack_convert_kernel<256><<<dim3(n_img, 256 / ConvertShape<256>::RB), 256, 0, st>>>(c->stack, c->raw, first_img); break;
  }
  c->launches++;
  CK(cudaGetLastError());
  c->have_stack = true;
  return FPMB200_OK;
}

// ---- generic 2-D FFT driver (rows then columns) ------------------------------------------
static void factorize(int n, LineFFTParams& p) {
  p.nrad = 0;
  while (n % 4 == 0) { p.rad[p.nrad++] = 4; n /= 4; }
  while (n % 2 == 0) { p.rad[p.nrad++] = 2; n /= 2; }
  while (n % 3 == 0) { p.rad[p.nrad++] = 3; n /= 3; }
  while (n % 5 == 0) { p.rad[p.nrad++] = 5; n /= 5; }
}

// 2-D transform of `batch` n x n images in place: rows, then columns.  row0 / nrows (col0 / ncols) restrict the row
// (column) pass to a wrapped range of lines: rows known to be zero need no transform, columns whose result is not
// used need none either (the update's O * P is zero outside the bounding box of the pupil support, and only the box
// of Phi' is consumed).  nrows = ncols = n: the full transform.
template <bool INV>
static int fft2d(fpmb200_ctx* c, float2* data, int n, const float2* tw, int batch, long long batch_stride, float scale,
                 cudaStream_t st, int row0 = 0, int nrows = -1, int col0 = 0, int ncols = -1) {
  if (nrows < 0) nrows = n;
  if (ncols < 0) ncols = n;
  LineFFTParams p;
  memset(&p, 0, sizeof p);
  p.data = data; p.tw = tw; p.batch_stride = batch_stride; p.n = n; p.n_lines = n;
  factorize(n, p);
  // 4 lines per CTA (16 adjacent lines per CTA for the column pass -- 128 contiguous bytes per sample index instead
  // of 32 -- measured the same at 148 tiles and slower for a single tile: fewer CTAs)
  constexpr int LINES = 4;
  const size_t smem = sizeof(float2) * 2 * LINES * (n + 1);
  if (smem > (size_t)c->max_smem_optin) return fail(FPMB200_ERR_ARG, "line FFT of length %d needs %zu B shared memory", n, smem);
  auto kern = line_fft_kernel<INV, LINES>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.elem_stride = 1; p.line_stride = n; p.scale = 1.f;            // rows
  p.line_first = row0; p.n_sel = nrows;
  kern<<<dim3((nrows + LINES - 1) / LINES, batch), 256, smem, st>>>(p);
  p.elem_stride = n; p.line_stride = 1; p.scale = scale;          // columns
  p.line_first = col0; p.n_sel = ncols;
  kern<<<dim3((ncols + LINES - 1) / LINES, batch), 256, smem, st>>>(p);
  c->launches += 2;
  CK(cudaGetLastError());
  return FPMB200_OK;
}

// n images of L x L through the planned kernels of fpm_fft2d.cuh (rows src -> dst, optionally through the fftShift;
// columns in place); FPMB200_ERR_STATE (without an error message) when L has no compiled plan
template <int R0, int R1, int R2, bool INV>
static int planned_fft2d_launch(fpmb200_ctx* c, const float2* src, float2* dst, const float2* tw, int n, int shift, float scale,
                                cudaStream_t st) {
  using PS = PlanShape<R0, R1, R2>;
  auto k = plan_fft_kernel<R0, R1, R2, INV>;
  if (PS::smem > (size_t)c->max_smem_optin) return FPMB200_ERR_STATE;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS::smem));
  PlanFFTParams p;
  memset(&p, 0, sizeof p);
  const long long LL = (long long)PS::L * PS::L;
  p.tw = tw; p.src_stride = LL; p.dst_stride = LL;
  const dim3 grid((PS::L + PS::LINES - 1) / PS::LINES, n);
  p.src = src; p.dst = dst; p.shift = shift; p.cols = 0; p.scale = 1.f;       // rows
  k<<<grid, 256, PS::smem, st>>>(p);
  p.src = dst; p.shift = 0; p.cols = 1; p.scale = scale;                      // columns, in place
  k<<<grid, 256, PS::smem, st>>>(p);
  c->launches += 2;
  CK(cudaGetLastError());
  return FPMB200_OK;
}
// objCrop = IDFT(fftShift(objFc)), fpmMain.cpp:481
static int planned_ifft2d(fpmb200_ctx* c, const float2* src, float2* dst, int n, cudaStream_t st) {
  const float sc = 1.0f / ((float)c->L * (float)c->L);
  switch (c->L) {
    case 256: return planned_fft2d_launch<16, 16, 1, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 384: return planned_fft2d_launch<8, 8, 6, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 512: return planned_fft2d_launch<8, 8, 8, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 1024: return planned_fft2d_launch<16, 8, 8, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 1536: return planned_fft2d_launch<16, 16, 6, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 600: return planned_fft2d_launch<10, 10, 6, true>(c, src, dst, c->twL, n, 1, sc, st);
    case 360: return planned_fft2d_launch<10, 6, 6, true>(c, src, dst, c->twL, n, 1, sc, st);
    default: return FPMB200_ERR_STATE;
  }
}
// forward Np x Np transform of the spectrum seed, in place (fpmMain.cpp:325)
static int planned_seed_fft2d(fpmb200_ctx* c, float2* data, int n, cudaStream_t st) {
  switch (c->N) {
    case 64: return planned_fft2d_launch<8, 8, 1, false>(c, data, data, c->twN, n, 0, 1.f, st);
    case 128: return planned_fft2d_launch<16, 8, 1, false>(c, data, data, c->twN, n, 0, 1.f, st);
    case 256: return planned_fft2d_launch<16, 16, 1, false>(c, data, data, c->twN, n, 0, 1.f, st);
    default: return FPMB200_ERR_STATE;
  }
}

extern "C" int fpmb200_init_tiles(fpmb200_ctx* c, int first, int n, int init_slot, void* stream) {
  int rc = check_range(c, first, n);
  if (rc) return rc;
  if (!c->have_support || !c->have_stack) return fail(FPMB200_ERR_STATE, "upload the pupil support and the stack first");
  if (init_slot < 0 || init_slot >= c->n_leds) return fail(FPMB200_ERR_ARG, "init_led_slot %d outside [0,%d)", init_slot, c->n_leds);
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const int N = c->N, L = c->L;
  const int batch_max = (int)(c->scratch_elems / ((size_t)N * N));
  for (int t0 = first; t0 < first + n; t0 += batch_max) {
    const int b = (first + n - t0) < batch_max ? (first + n - t0) : batch_max;
    if (c->general) gen_init_amp<<<dim3(16, b), 256, 0, st>>>(c->scratch, c->stack, c->n_leds, init_slot, t0, N, c->stack_r1);
    else switch (N) {
      case 64: init_amp_kernel<64><<<dim3(16, b), 256, 0, st>>>(c->scratch, c->stack, c->n_leds, init_slot, t0); break;
      case 128: init_amp_kernel<128><<<dim3(16, b), 256, 0, st>>>(c->scratch, c->stack, c->n_leds, init_slot, t0); break;
      case 256: init_amp_kernel<256><<<dim3(16, b), 256, 0, st>>>(c->scratch, c->stack, c->n_leds, init_slot, t0); break;
    }
    c->launches++;
    {
      const char* e = getenv("FPMB200_FINALIZE_GENERIC");
      rc = (e && e[0] == '1') ? FPMB200_ERR_STATE : planned_seed_fft2d(c, c->scratch, b, st);
      if (rc == FPMB200_ERR_STATE) rc = fft2d<false>(c, c->scratch, N, c->twN, b, (long long)N * N, 1.f, st);
      if (rc) return rc;
    }
    init_place_kernel<<<dim3(64, b), 256, 0, st>>>(c->objFc, c->pupil, c->scratch, c->support, N, L, t0);
    c->launches++;
  }
  CK(cudaGetLastError());
  return FPMB200_OK;
}

template <int N, int NT, int MINB>
static int launch_update(fpmb200_ctx* c, const UpdateParams& p, int n_blocks, cudaStream_t st) {
  void (*k)(const UpdateParams) = nullptr;   // (declared __grid_constant__ in the kernel)
  constexpr bool FS = (N <= 128);
  if constexpr (N == 128 || N == 64) {
    if (c->phased) {
      UpdateParams pn = p;
      PhasedShape<N>::layout(c->yhi - c->ylo + 1, c->xhi - c->xlo + 1, c->ocp, c->L, pn.noff);
      k = fpm_update_phased_kernel<N, 512>;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
      k<<<n_blocks, 512, c->smem_bytes, st>>>(pn);
      c->launches++;
      CK(cudaGetLastError());
      return FPMB200_OK;
    }
  }
  if constexpr (N == 128) {
    if (false) {}
#ifndef FPM_DEV_FAST
    else if (c->p_smem && c->q_smem && c->narrow) k = fpm_update_kernel<N, NT, MINB, FS, true, true, true>;
#endif
  }
  if (k) {}
#ifndef FPM_DEV_FAST
  else if (c->p_smem && c->q_smem) k = fpm_update_kernel<N, NT, MINB, FS, true, true>;
  else if (c->p_smem) k = fpm_update_kernel<N, NT, MINB, FS, true, false>;
  else k = fpm_update_kernel<N, NT, MINB, FS, false, false>;
#else
  else return fail(FPMB200_ERR_STATE, "FPM_DEV_FAST build: only the narrow 128 x 128 kernel is compiled");
#endif
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  k<<<n_blocks, NT, c->smem_bytes, st>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  return FPMB200_OK;
}

template <int N, int C>
static int launch_cluster(fpmb200_ctx* c, const UpdateParams& p, int n_tiles, cudaStream_t st) {
  auto k = cluster_kernel_for<N, C>(c);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(n_tiles * C, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = c->smem_bytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k, p));
  c->launches++;
  return FPMB200_OK;
}

// The unfused path (csrc/fpm_general.cuh): 11 launches per update, every launch covers tiles [first, first+n).
static void drop_graphs(fpmb200_ctx* c) {
  for (auto& g : c->iter_graphs) cudaGraphExecDestroy(g.exec);
  c->iter_graphs.clear();
}

// threads per CTA of fpm_update_general_kernel
#ifndef FPM_GEN_NT
#define FPM_GEN_NT 512
#endif
static constexpr int GEN_NT = FPM_GEN_NT;


// R1 * 100 + R2 of the compiled two-stage plan for Np, 0 if there is none
static int general_fused_plan(int N) {
  const char* e = getenv("FPMB200_GENERAL_PLAN");
  if (e && e[0] == '0') return 0;
  switch (N) {
    case 90: return 1009;
    case 100: return 1010;
    case 80: return 1008;
    case 72: return 908;
    case 96: return 1606;
    case 60: return 1006;
    default: return 0;
  }
}

static int run_updates_general_fused(fpmb200_ctx* c, int first, int n, int slot_begin, int n_updates, cudaStream_t st) {
  GeneralFusedParams p;
  memset(&p, 0, sizeof p);
  p.objFc = c->objFc; p.pupil = c->pupil; p.stack = c->stack; p.support = c->support; p.crop = c->crop; p.tw = c->twN;
  p.N = c->N; p.L = c->L; p.n_leds = c->n_leds; p.tile0 = first; p.slot_begin = slot_begin; p.n_updates = n_updates;
  p.cgr = c->cgr; p.cgc = c->cgc;
  p.ylo = c->ylo; p.yhi = c->yhi; p.xlo = c->xlo; p.xhi = c->xhi;
  p.delta1 = c->delta1; p.delta2 = c->delta2; p.eps = c->eps; p.kappa = c->kappa;
  LineFFTParams fp;
  factorize(c->N, fp);
  p.nrad = fp.nrad;
  for (int s = 0; s < fp.nrad; ++s) p.rad[s] = fp.rad[s];
#ifdef FPM_STAGE_TIMING
  if (!c->stage_clk) { CK(cudaMalloc(&c->stage_clk, 16 * sizeof(long long))); CK(cudaMemset(c->stage_clk, 0, 16 * sizeof(long long))); }
  p.stage_clk = c->stage_clk;
#endif
  // two-stage plans with compile-time radices for the sizes of the shipped JSONs (and their neighbours); other sizes
  // take the radices at run time
#ifdef FPM_DEV_FAST
  void (*k)(const GeneralFusedParams) = nullptr;
  return fail(FPMB200_ERR_STATE, "FPM_DEV_FAST build");
#else
  void (*k)(const GeneralFusedParams) = fpm_update_general_kernel<GEN_NT, 0, 0>;
  switch (general_fused_plan(c->N)) {
    case 1009: k = fpm_update_general_kernel<GEN_NT, 10, 9>; break;
    case 1010: k = fpm_update_general_kernel<GEN_NT, 10, 10>; break;
    case 1008: k = fpm_update_general_kernel<GEN_NT, 10, 8>; break;
    case 908: k = fpm_update_general_kernel<GEN_NT, 9, 8>; break;
    case 1606: k = fpm_update_general_kernel<GEN_NT, 16, 6>; break;
    case 1006: k = fpm_update_general_kernel<GEN_NT, 10, 6>; break;
    default: break;
  }
#endif
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  k<<<n, GEN_NT, c->smem_bytes, st>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  return FPMB200_OK;
}

// R1 * 100 + R2 of the compiled in-place plan of fpm_update_pruned_kernel for Np, 0 if there is none
static int pruned_plan(int N) {
  // (measured: the shipped small tiles through this kernel at two CTAs of 320 threads per SM -- plans 10 x 9 and 10 x 10,
  //  296 tiles -- give 10.8 M updates/s at Np = 90 against 9.65 M of fpm_update_general_kernel and the same 7.49 M at
  //  Np = 100; not worth a second stack layout for those sizes)
  switch (N) {
    case 200: return 2010;      // dataset_dogStomach.json as shipped
    case 160: return 1610;
    case 240: return 2012;
    case 300: return 2015;
    case 128: return 1608;      // only reachable with FPMB200_FORCE_GENERAL=1
    default: return 0;
  }
}

static int run_updates_pruned(fpmb200_ctx* c, int first, int n, int slot_begin, int n_updates, cudaStream_t st) {
  PrunedParams p;
  memset(&p, 0, sizeof p);
  p.objFc = c->objFc; p.pupil = c->pupil; p.stack = c->stack; p.support = c->support; p.crop = c->crop; p.tw = c->twN;
  p.L = c->L; p.n_leds = c->n_leds; p.tile0 = first; p.slot_begin = slot_begin; p.n_updates = n_updates;
  p.cgr = c->cgr; p.cgc = c->cgc; p.cb = c->cb;
  p.ylo = c->ylo; p.yhi = c->yhi; p.xlo = c->xlo; p.xhi = c->xhi;
  p.delta1 = c->delta1; p.delta2 = c->delta2; p.eps = c->eps; p.kappa = c->kappa;
#ifdef FPM_STAGE_TIMING
  if (!c->stage_clk) { CK(cudaMalloc(&c->stage_clk, 16 * sizeof(long long))); CK(cudaMemset(c->stage_clk, 0, 16 * sizeof(long long))); }
  p.stage_clk = c->stage_clk;
#endif
  void (*k)(const PrunedParams) = nullptr;
  const bool nw = c->narrow;      // the box lies within +-3*R2: six-sample butterflies in the stage-A passes (R1 = 20 plans)
#ifdef FPM_DEV_FAST
  (void)nw;
  return fail(FPMB200_ERR_STATE, "FPM_DEV_FAST build");
#else
  switch (pruned_plan(c->N)) {
    case 2010: k = nw ? fpm_update_pruned_kernel<PRUNED_NT, 20, 10, true> : fpm_update_pruned_kernel<PRUNED_NT, 20, 10, false>; break;
    case 1610: k = fpm_update_pruned_kernel<PRUNED_NT, 16, 10, false>; break;
    case 2012: k = nw ? fpm_update_pruned_kernel<PRUNED_NT, 20, 12, true> : fpm_update_pruned_kernel<PRUNED_NT, 20, 12, false>; break;
    case 2015: k = nw ? fpm_update_pruned_kernel<PRUNED_NT, 20, 15, true> : fpm_update_pruned_kernel<PRUNED_NT, 20, 15, false>; break;
    case 1608: k = fpm_update_pruned_kernel<PRUNED_NT, 16, 8, false>; break;
    default: return fail(FPMB200_ERR_STATE, "no pruned plan for Np=%d", c->N);
  }
#endif
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  k<<<n, PRUNED_NT, c->smem_bytes, st>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  return FPMB200_OK;
}

static int run_updates_general(fpmb200_ctx* c, int first, int n, int slot_begin, int n_updates, cudaStream_t st) {
  if (c->gfused) return run_updates_general_fused(c, first, n, slot_begin, n_updates, st);
  if (c->gpruned) return run_updates_pruned(c, first, n, slot_begin, n_updates, st);
  GeneralParams p;
  memset(&p, 0, sizeof p);
  const int N = c->N;
  p.objFc = c->objFc; p.pupil = c->pupil; p.stack = c->stack; p.support = c->support; p.crop = c->crop;
  p.field = c->gfield; p.q = c->gq; p.cells = c->gcells; p.scal = c->gscal;
  p.N = N; p.L = c->L; p.n_leds = c->n_leds; p.tile0 = first; p.cgr = c->cgr; p.cgc = c->cgc;
  p.delta1 = c->delta1; p.delta2 = c->delta2; p.eps = c->eps; p.kappa = c->kappa;
  p.stack_r1 = c->stack_r1;
  // bounding box of the pupil support: P, Q and the object increment vanish outside it
  p.ylo = c->ylo; p.xlo = c->xlo; p.nrb = c->yhi - c->ylo + 1; p.ncb = c->xhi - c->xlo + 1;
  const int row0 = c->ylo < 0 ? c->ylo + N : c->ylo, col0 = c->xlo < 0 ? c->xlo + N : c->xlo;
  const int bx = (N * N + 255) / 256 < 64 ? (N * N + 255) / 256 : 64;
  const dim3 ge(bx, n);
  const int nbx = (p.nrb * p.ncb + 255) / 256 < 64 ? (p.nrb * p.ncb + 255) / 256 : 64;
  const dim3 gb(nbx, n);                                          // kernels over the box
  const int tc = (p.nrb > p.ncb ? p.nrb : p.ncb) / 16 + 2;       // cells the box of a window can touch per dimension
  float2* fld = c->gfield + (size_t)first * N * N;
  // state the loop carries in scal: max|P|^2 of the current pupil; the cell grid of the current spectrum
  p.apply = 1; p.slot = 0;
  gen_cells_update<<<dim3(c->cgr * c->cgc, n), 256, 0, st>>>(p, 1);
  gen_cells_max<<<n, 256, 0, st>>>(p);
  p.apply = 0;
  gen_pupil_update<<<gb, 256, 0, st>>>(p);
  c->launches += 3;
  p.apply = 1;
  auto enqueue = [&](int u0, int u1) -> int {
    int rc;
    for (int u = u0; u < u1; ++u) {
      p.slot = (slot_begin + u) % c->n_leds;
      gen_window_mul<<<ge, 256, 0, st>>>(p);
      // inverse: only the box's rows are non-zero; forward: only the box's columns are consumed
      if ((rc = fft2d<true>(c, fld, N, c->twN, n, (long long)N * N, 1.f / ((float)N * (float)N), st, row0, p.nrb, 0, N))) return rc;
      gen_amplitude<<<ge, 256, 0, st>>>(p);
      if ((rc = fft2d<false>(c, fld, N, c->twN, n, (long long)N * N, 1.f, st, 0, N, col0, p.ncb))) return rc;
      gen_object_update<<<gb, 256, 0, st>>>(p);
      gen_cells_update<<<dim3(tc * tc, n), 256, 0, st>>>(p, 0);
      gen_cells_max<<<n, 256, 0, st>>>(p);
      gen_pupil_update<<<gb, 256, 0, st>>>(p);
      c->launches += 10;
    }
    return FPMB200_OK;
  };
  int rc;
  // whole passes over the LED list replay a captured graph (the loop is launch-bound for a handful of tiles)
  const int nl = c->n_leds;
  if (slot_begin == 0 && n_updates >= nl && n_updates % nl == 0) {
    fpmb200_ctx::IterGraph* g = nullptr;
    for (auto& e : c->iter_graphs) if (e.first == first && e.n == n) g = &e;
    if (!g) {
      const long long l0 = c->launches;
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      rc = enqueue(0, nl);
      cudaError_t e2 = cudaStreamEndCapture(st, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      CK(e2);
      cudaGraphExec_t exec = nullptr;
      CK(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      c->iter_graphs.push_back({first, n, exec, c->launches - l0});
      c->launches = l0;
      g = &c->iter_graphs.back();
    }
    for (int it = 0; it < n_updates / nl; ++it) {
      CK(cudaGraphLaunch(g->exec, st));
      c->launches += g->nodes;
    }
    return FPMB200_OK;
  }
  if ((rc = enqueue(0, n_updates))) return rc;
  CK(cudaGetLastError());
  return FPMB200_OK;
}

static int run_updates(fpmb200_ctx* c, int first, int n, int slot_begin, int n_updates, cudaStream_t st, const int* tile_list = nullptr,
                       float* ucache = nullptr) {
  if (!c->have_support || !c->have_leds) return fail(FPMB200_ERR_STATE, "upload LED tables and the pupil support first");
  if (c->general) { CK(cudaSetDevice(c->device)); return run_updates_general(c, first, n, slot_begin, n_updates, st); }
  UpdateParams p;
  memset(&p, 0, sizeof p);
  p.objFc = c->objFc; p.pupil = c->pupil; p.stack = c->stack; p.support = c->support; p.crop = c->crop;
  p.tw = c->twN; p.field_gmem = c->field_gmem; p.L = c->L; p.n_leds = c->n_leds; p.tile0 = first;
  p.slot_begin = slot_begin; p.n_updates = n_updates; p.tile_list = tile_list; p.ucache = ucache;
  p.delta1 = c->delta1; p.delta2 = c->delta2; p.eps = c->eps; p.kappa = c->kappa;
  p.ylo = c->ylo; p.yhi = c->yhi; p.xlo = c->xlo; p.xhi = c->xhi; p.cs = c->cs; p.ocp = c->ocp;
  if (c->have_tmap) p.tmap = c->tmap;
#ifdef FPM_STAGE_TIMING
  if (!c->stage_clk) { CK(cudaMalloc(&c->stage_clk, 16 * sizeof(long long))); CK(cudaMemset(c->stage_clk, 0, 16 * sizeof(long long))); }
  p.stage_clk = c->stage_clk;
#endif
  CK(cudaSetDevice(c->device));
  if (c->cluster > 1) {
    p.ocp = c->cpc;
#if !defined(FPM_DEV_FAST) || defined(FPM_DEV_CLUSTER)
    if (c->N == 256 && c->cluster == 8) return launch_cluster<256, 8>(c, p, n, st);
    if (c->N == 128 && c->cluster == 4) return launch_cluster<128, 4>(c, p, n, st);
    if (c->N == 128 && c->cluster == 2) return launch_cluster<128, 2>(c, p, n, st);
#endif
    return fail(FPMB200_ERR_STATE, "no cluster kernel for Np=%d x %d CTAs", c->N, c->cluster);
  }
  switch (c->N) {
#ifndef FPM_DEV_FAST
    case 256: return launch_update<256, 512, 1>(c, p, n, st);
#endif
    case 64: return launch_update<64, 512, 1>(c, p, n, st);
    case 128: return launch_update<128, 512, 1>(c, p, n, st);
  }
  return fail(FPMB200_ERR_ARG, "unsupported Np");
}

extern "C" int fpmb200_run(fpmb200_ctx* c, int first, int n, int iters, void* stream) {
  int rc = check_range(c, first, n);
  if (rc) return rc;
  if (iters < 0) return fail(FPMB200_ERR_ARG, "iters < 0");
  if (iters == 0) return FPMB200_OK;
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  // One persistent CTA per tile runs all iterations; with more tiles than SMs the last wave is partly empty (320 tiles =
  // 148 + 148 + 24: the third wave leaves 124 SMs idle for a third of the run).  Tiles are independent and an iteration
  // boundary is a clean cut (the launch-to-launch state is objFc / pupil in global memory: "steps == run", bit for bit),
  // so the run is re-cut into passes of ONE iteration over at most sm_count tiles, tiles with the most iterations left
  // first: max(iters, ceil(n * iters / sm_count)) passes instead of ceil(n / sm_count) * iters.  Each pass pays the
  // kernel's prologue again (cell maxima from the spectrum, pupil, first windows: a few per cent of an iteration).
  // FPMB200_RUN_BALANCED=0 keeps the single launch.
  const int S = c->sm_count;
  const char* e = getenv("FPMB200_RUN_BALANCED");
  if (!c->general && c->cluster == 1 && n > S && iters >= 2 && !(e && e[0] == '0')) {
    const long long waves = (long long)((n + S - 1) / S) * iters;
    const long long passes = std::max<long long>(iters, ((long long)n * iters + S - 1) / S);
    if (passes * 21 < waves * 20) {                          // worth at least the 5 % the extra prologues may cost
      CK(cudaSetDevice(c->device));
      if (c->sched_first != first || c->sched_n != n || c->sched_iters != iters) {       // (re)build and upload the schedule
        std::vector<int> left(n, iters), lists, order(n);
        c->sched_counts.clear();
        for (long long done = 0; done < (long long)n * iters;) {
          // tiles with the most iterations left first (stable: lower index first)
          for (int t = 0; t < n; ++t) order[t] = t;
          std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return left[a] > left[b]; });
          int cnt = 0;
          for (int k = 0; k < n && cnt < S; ++k)
            if (left[order[k]] > 0) {
              // bit 30: not the tile's first pass of this run -- the phased kernel finds its cell maxima in c->ucache
              lists.push_back((first + order[k]) | (left[order[k]] < iters ? 0x40000000 : 0));
              --left[order[k]]; ++cnt;
            }
          c->sched_counts.push_back(cnt);
          done += cnt;
        }
        c->sched_first = -1;
        CK(cudaDeviceSynchronize());                         // earlier passes (on any stream) may still read the old lists
        if (lists.size() > c->sched_cap) {
          cudaFree(c->sched_dev); c->sched_dev = nullptr; c->sched_cap = 0;
          CK(cudaMalloc(&c->sched_dev, sizeof(int) * lists.size()));
          c->sched_cap = lists.size();
        }
        CK(cudaMemcpyAsync(c->sched_dev, lists.data(), sizeof(int) * lists.size(), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        c->sched_first = first; c->sched_n = n; c->sched_iters = iters;
      }
      if (c->phased && !c->ucache)
        CK(cudaMalloc(&c->ucache, sizeof(float) * (size_t)c->n_tiles * c->L * (c->L >> 4)));
      const std::vector<int>& counts = c->sched_counts;
      size_t off = 0;
      for (int cnt : counts) {
        if ((rc = run_updates(c, first, cnt, 0, c->n_leds, st, c->sched_dev + off, c->phased ? c->ucache : nullptr))) return rc;
        off += cnt;
      }
      return FPMB200_OK;
    }
  }
  return run_updates(c, first, n, 0, iters * c->n_leds, st);
}

extern "C" int fpmb200_step(fpmb200_ctx* c, int tile, int led_slot) {
  int rc = check_range(c, tile, 1);
  if (rc) return rc;
  if (led_slot < 0 || led_slot >= c->n_leds) return fail(FPMB200_ERR_ARG, "led_slot %d outside [0,%d)", led_slot, c->n_leds);
  if ((rc = run_updates(c, tile, 1, led_slot, 1, c->stream))) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return FPMB200_OK;
}

extern "C" int fpmb200_finalize(fpmb200_ctx* c, int first, int n, void* stream) {
  int rc = check_range(c, first, n);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t LL = (size_t)c->L * c->L;
  // Nlarge with a compiled plan: two launches, the fftShift folded into the row pass's loads (fpm_fft2d.cuh);
  // FPMB200_FINALIZE_GENERIC=1 keeps the run-time-radix path (developer A/B)
  {
    const char* e = getenv("FPMB200_FINALIZE_GENERIC");
    if (!(e && e[0] == '1')) {
      rc = planned_ifft2d(c, c->objFc + LL * first, c->objCrop + LL * first, n, st);
      if (rc != FPMB200_ERR_STATE) return rc;
    }
  }
  shift_copy_kernel<<<dim3(64, n), 256, 0, st>>>(c->objCrop + LL * first, c->objFc + LL * first, c->L, (long long)LL, (long long)LL);
  c->launches++;
  return fft2d<true>(c, c->objCrop + LL * first, c->L, c->twL, n, (long long)LL, 1.0f / ((float)c->L * (float)c->L), st);
}

extern "C" int fpmb200_upload_state(fpmb200_ctx* c, int tile, const float* objF, const float* pupil) {
  int rc = check_range(c, tile, 1);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const size_t LL = (size_t)c->L * c->L, NN = (size_t)c->N * c->N;
  CK(cudaDeviceSynchronize());
  if (objF) {
    CK(copy_sync(c, c->scratch, objF, sizeof(float2) * LL, cudaMemcpyHostToDevice));
    shift_copy_kernel<<<dim3(64, 1), 256, 0, c->stream>>>(c->objFc + LL * tile, c->scratch, c->L, 0, 0);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
  }
  if (pupil) CK(copy_sync(c, c->pupil + NN * tile, pupil, sizeof(float2) * NN, cudaMemcpyHostToDevice));
  return FPMB200_OK;
}

extern "C" int fpmb200_download(fpmb200_ctx* c, int tile, float* objF, float* objCrop, float* pupil) {
  int rc = check_range(c, tile, 1);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const size_t LL = (size_t)c->L * c->L, NN = (size_t)c->N * c->N;
  CK(cudaDeviceSynchronize());
  if (objF) {
    shift_copy_kernel<<<dim3(64, 1), 256, 0, c->stream>>>(c->scratch, c->objFc + LL * tile, c->L, 0, 0);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
    CK(copy_sync(c, objF, c->scratch, sizeof(float2) * LL, cudaMemcpyDeviceToHost));
  }
  if (objCrop) CK(copy_sync(c, objCrop, c->objCrop + LL * tile, sizeof(float2) * LL, cudaMemcpyDeviceToHost));
  if (pupil) CK(copy_sync(c, pupil, c->pupil + NN * tile, sizeof(float2) * NN, cudaMemcpyDeviceToHost));
  return FPMB200_OK;
}

extern "C" int fpmb200_download_objcrop(fpmb200_ctx* c, int first, int n, float* objCrop, void* stream) {
  int rc = check_range(c, first, n);
  if (rc) return rc;
  if (!objCrop) return fail(FPMB200_ERR_ARG, "objCrop is NULL");
  CK(cudaSetDevice(c->device));
  const size_t LL = (size_t)c->L * c->L;
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  CK(cudaMemcpyAsync(objCrop, c->objCrop + LL * first, sizeof(float2) * LL * n, cudaMemcpyDeviceToHost, st));
  return FPMB200_OK;
}

extern "C" int fpmb200_device_buffer(fpmb200_ctx* c, int which, int tile, void** ptr, unsigned long long* bytes) {
  int rc = check_range(c, tile, 1);
  if (rc) return rc;
  if (!ptr) return fail(FPMB200_ERR_ARG, "ptr is NULL");
  const size_t LL = (size_t)c->L * c->L, NN = (size_t)c->N * c->N;
  size_t b = 0;
  switch (which) {
    case 0: b = sizeof(float2) * LL; *ptr = c->objFc + LL * tile; break;
    case 1: b = sizeof(float2) * LL; *ptr = c->objCrop + LL * tile; break;
    case 2: b = sizeof(float2) * NN; *ptr = c->pupil + NN * tile; break;
    case 3: b = sizeof(float) * NN * c->n_leds; *ptr = c->stack + NN * c->n_leds * tile; break;
    case 4: b = sizeof(uint16_t) * NN * c->n_leds; *ptr = c->raw + NN * c->n_leds * tile; break;
    default: return fail(FPMB200_ERR_ARG, "which=%d not in 0..4", which);
  }
  if (bytes) *bytes = b;
  return FPMB200_OK;
}

// ---- full field of view: frame ingest and mosaic ---------------------------------------
extern "C" int fpmb200_set_tile_origins(fpmb200_ctx* c, const int32_t* x, const int32_t* y, int n_tiles) {
  if (!c || !x || !y) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->n_tiles) return fail(FPMB200_ERR_STATE, "fpmb200_tiles_alloc first");
  if (n_tiles != c->n_tiles) return fail(FPMB200_ERR_ARG, "n_tiles=%d differs from the allocation (%d)", n_tiles, c->n_tiles);
  std::vector<int2> h(n_tiles);
  for (int t = 0; t < n_tiles; ++t) {
    if (x[t] < 0 || y[t] < 0) return fail(FPMB200_ERR_ARG, "tile %d: negative ROI origin (%d,%d)", t, x[t], y[t]);
    h[t] = make_int2(x[t], y[t]);
  }
  CK(cudaSetDevice(c->device));
  if (!c->origins) CK(cudaMalloc(&c->origins, sizeof(int2) * n_tiles));
  if (!c->bg_dev) { CK(cudaMalloc(&c->bg_dev, sizeof(int) * c->n_leds)); CK(cudaMemsetAsync(c->bg_dev, 0, sizeof(int) * c->n_leds, c->stream)); }
  CK(copy_sync(c, c->origins, h.data(), sizeof(int2) * n_tiles, cudaMemcpyHostToDevice));
  c->have_origins = true;
  // the largest ROI corner, checked against every frame
  c->origin_max_x = 0; c->origin_max_y = 0; c->origin_min_y = y[0];
  for (int t = 0; t < n_tiles; ++t) {
    if (x[t] > c->origin_max_x) c->origin_max_x = x[t];
    if (y[t] > c->origin_max_y) c->origin_max_y = y[t];
    if (y[t] < c->origin_min_y) c->origin_min_y = y[t];
  }
  return FPMB200_OK;
}

extern "C" int fpmb200_ingest_frame(fpmb200_ctx* c, int led_slot, const uint16_t* frame, int width, int height, int divisor,
                                    int bk1x, int bk1y, int bk2x, int bk2y, int bg_threshold, void* stream) {
  if (!c || !frame) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->have_origins) return fail(FPMB200_ERR_STATE, "fpmb200_set_tile_origins first");
  if (led_slot < 0 || led_slot >= c->n_leds) return fail(FPMB200_ERR_ARG, "led_slot %d outside [0,%d)", led_slot, c->n_leds);
  const int Np = c->N;
  auto inside = [&](int x, int y) { return x >= 0 && y >= 0 && x + Np <= width && y + Np <= height; };
  if (width <= 0 || height <= 0 || !inside(c->origin_max_x, c->origin_max_y) || !inside(bk1x, bk1y) || !inside(bk2x, bk2y))
    return fail(FPMB200_ERR_ARG, "a tile or background ROI leaves the %dx%d frame", width, height);
  if (divisor < 0) return fail(FPMB200_ERR_ARG, "divisor < 0");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t elems = (size_t)width * height;
  if (elems > c->frame_elems) {
    CK(cudaStreamSynchronize(st));
    cudaFree(c->frame_dev);
    c->frame_dev = nullptr; c->frame_elems = 0;
    CK(cudaMalloc(&c->frame_dev, sizeof(uint16_t) * elems));
    c->frame_elems = elems;
  }
  CK(cudaMemcpyAsync(c->frame_dev, frame, sizeof(uint16_t) * elems, cudaMemcpyHostToDevice, st));
  ingest_bg_kernel<<<1, 1024, 0, st>>>(c->frame_dev, width, Np, bk1x, bk1y, bk2x, bk2y, bg_threshold, c->bg_dev + led_slot);
  const int R1 = c->general ? (c->stack_r1 ? c->stack_r1 : 1) : (Np == 64 ? 8 : 16);
  const int bx = (Np * Np + 255) / 256 < 16 ? (Np * Np + 255) / 256 : 16;
  ingest_tiles_kernel<<<dim3(bx, c->n_tiles), 256, 0, st>>>(c->frame_dev, width, c->origins, 0, c->raw, c->stack, c->n_leds,
                                                             led_slot, Np, R1, c->general ? (c->stack_r1 ? 2 : 0) : 1, divisor, c->bg_dev + led_slot, 0);
  c->launches += 2;
  CK(cudaGetLastError());
  c->have_stack = true;
  return FPMB200_OK;
}

extern "C" int fpmb200_ingest_rows(fpmb200_ctx* c, int led_slot, const uint16_t* rows, int width, int row0, int n_rows, int divisor,
                                   int bg_val, void* stream) {
  if (!c || !rows) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->have_origins) return fail(FPMB200_ERR_STATE, "fpmb200_set_tile_origins first");
  if (led_slot < 0 || led_slot >= c->n_leds) return fail(FPMB200_ERR_ARG, "led_slot %d outside [0,%d)", led_slot, c->n_leds);
  const int Np = c->N;
  if (width <= 0 || n_rows <= 0 || row0 < 0 || c->origin_max_x + Np > width || c->origin_min_y < row0 ||
      c->origin_max_y + Np > row0 + n_rows)
    return fail(FPMB200_ERR_ARG, "rows [%d,%d) x %d columns do not cover this context's tiles (ROI rows %d..%d, columns ..%d)", row0,
                row0 + n_rows, width, c->origin_min_y, c->origin_max_y + Np, c->origin_max_x + Np);
  if (divisor < 0) return fail(FPMB200_ERR_ARG, "divisor < 0");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  const size_t elems = (size_t)width * n_rows;
  if (elems > c->frame_elems) {
    CK(cudaStreamSynchronize(st));
    cudaFree(c->frame_dev);
    c->frame_dev = nullptr; c->frame_elems = 0;
    CK(cudaMalloc(&c->frame_dev, sizeof(uint16_t) * elems));
    c->frame_elems = elems;
  }
  CK(cudaMemcpyAsync(c->frame_dev, rows, sizeof(uint16_t) * elems, cudaMemcpyHostToDevice, st));
  const int R1 = c->general ? (c->stack_r1 ? c->stack_r1 : 1) : (Np == 64 ? 8 : 16);
  const int bx = (Np * Np + 255) / 256 < 16 ? (Np * Np + 255) / 256 : 16;
  // the kernel addresses the frame by absolute row: hand it the (virtual) address of row 0
  ingest_tiles_kernel<<<dim3(bx, c->n_tiles), 256, 0, st>>>(c->frame_dev - (ptrdiff_t)row0 * width, width, c->origins, 0, c->raw, c->stack,
                                                             c->n_leds, led_slot, Np, R1, c->general ? (c->stack_r1 ? 2 : 0) : 1, divisor,
                                                             nullptr, bg_val);
  set_bg_kernel<<<1, 1, 0, st>>>(c->bg_dev + led_slot, bg_val);
  c->launches += 2;
  CK(cudaGetLastError());
  c->have_stack = true;
  return FPMB200_OK;
}

extern "C" int fpmb200_ingest_bg(fpmb200_ctx* c, int32_t* bg_val) {
  if (!c || !bg_val) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->bg_dev) return fail(FPMB200_ERR_STATE, "no frame ingested yet");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  CK(copy_sync(c, bg_val, c->bg_dev, sizeof(int) * c->n_leds, cudaMemcpyDeviceToHost));
  return FPMB200_OK;
}

extern "C" int fpmb200_mosaic(fpmb200_ctx* c, const void* tiles_device, int nx, int ny, int step, float* out, int out_on_device,
                              void* stream) {
  if (!c || !out) return fail(FPMB200_ERR_ARG, "NULL argument");
  if (!c->n_tiles) return fail(FPMB200_ERR_STATE, "fpmb200_tiles_alloc first");
  if (nx <= 0 || ny <= 0 || step <= 0 || step > c->N) return fail(FPMB200_ERR_ARG, "need nx, ny > 0 and 0 < step <= Np");
  if (!tiles_device && nx * ny > c->n_tiles) return fail(FPMB200_ERR_ARG, "%dx%d tiles exceed the allocation (%d)", nx, ny, c->n_tiles);
  if (c->L % c->N) return fail(FPMB200_ERR_ARG, "Nlarge is not a multiple of Np");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  MosaicParams p;
  p.tiles = tiles_device ? (const float2*)tiles_device : c->objCrop;
  p.L = c->L; p.Np = c->N; p.step = step; p.nx = nx; p.ny = ny;
  const int f = c->L / c->N;
  p.Wm = ((nx - 1) * step + c->N) * f; p.Hm = ((ny - 1) * step + c->N) * f;
  const size_t elems = (size_t)p.Wm * p.Hm;
  if (out_on_device) p.out = out;
  else {
    if (elems > c->mosaic_elems) {
      CK(cudaStreamSynchronize(st));
      cudaFree(c->mosaic_dev);
      c->mosaic_dev = nullptr; c->mosaic_elems = 0;
      CK(cudaMalloc(&c->mosaic_dev, sizeof(float) * elems));
      c->mosaic_elems = elems;
    }
    p.out = c->mosaic_dev;
  }
  mosaic_kernel<<<c->sm_count * 8, 256, 0, st>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  if (!out_on_device) CK(cudaMemcpyAsync(out, c->mosaic_dev, sizeof(float) * elems, cudaMemcpyDeviceToHost, st));
  return FPMB200_OK;
}

// ---- single-process multi-GPU hand-off: a plain device buffer on this context's GPU and a (peer) copy of finished
//      objCrop tiles into it -- the "final gather" of a full-FOV run driven by one host thread ----
extern "C" int fpmb200_device_alloc(fpmb200_ctx* c, unsigned long long bytes, void** ptr) {
  if (!c || !ptr) return fail(FPMB200_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMalloc(ptr, bytes));
  return FPMB200_OK;
}
extern "C" int fpmb200_device_free(fpmb200_ctx* c, void* ptr) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaFree(ptr));
  return FPMB200_OK;
}
extern "C" int fpmb200_copy_objcrop_to(fpmb200_ctx* src, int tile_first, int n, fpmb200_ctx* dst, void* dst_ptr, void* stream) {
  int rc = check_range(src, tile_first, n);
  if (rc) return rc;
  if (!dst || !dst_ptr) return fail(FPMB200_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(src->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : src->stream;
  const size_t LL = (size_t)src->L * src->L;
  CK(cudaMemcpyPeerAsync(dst_ptr, dst->device, src->objCrop + LL * tile_first, src->device, sizeof(float2) * LL * n, st));
  return FPMB200_OK;
}

extern "C" int fpmb200_host_alloc(unsigned long long bytes, int write_combined, void** ptr) {
  if (!ptr || !bytes) return fail(FPMB200_ERR_ARG, "NULL argument or zero size");
  *ptr = nullptr;
  CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return FPMB200_OK;
}
extern "C" int fpmb200_host_free(void* ptr) {
  if (ptr) CK(cudaFreeHost(ptr));
  return FPMB200_OK;
}

extern "C" int fpmb200_event_record(fpmb200_ctx* c, int slot, void* stream) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  if (slot < 0 || slot >= 64) return fail(FPMB200_ERR_ARG, "marker slot %d outside 0..63", slot);
  CK(cudaSetDevice(c->device));
  if (!c->events[slot]) CK(cudaEventCreateWithFlags(&c->events[slot], cudaEventDisableTiming));
  CK(cudaEventRecord(c->events[slot], stream ? (cudaStream_t)stream : c->stream));
  return FPMB200_OK;
}
extern "C" int fpmb200_event_sync(fpmb200_ctx* c, int slot) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  if (slot < 0 || slot >= 64) return fail(FPMB200_ERR_ARG, "marker slot %d outside 0..63", slot);
  if (!c->events[slot]) return FPMB200_OK;
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->events[slot]));
  return FPMB200_OK;
}

extern "C" int fpmb200_sync(fpmb200_ctx* c) {
  if (!c) return fail(FPMB200_ERR_ARG, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  return FPMB200_OK;
}

#ifdef FPM_STAGE_TIMING
extern "C" int fpmb200_stage_clocks(fpmb200_ctx* c, long long* out16) {
  if (!c || !c->stage_clk) return -1;
  cudaDeviceSynchronize();
  cudaMemcpy(out16, c->stage_clk, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(c->stage_clk, 0, 16 * sizeof(long long));
  return 0;
}
#endif
extern "C" long long fpmb200_kernel_launches(const fpmb200_ctx* c) { return c ? c->launches : 0; }
extern "C" const char* fpmb200_variant(const fpmb200_ctx* c) { return c ? c->variant : ""; }
