/* fpmb200.h -- C ABI of the B200-native FPM reconstruction path (libfpmb200.so).
 *
 * The reference (Xiongda337/fpm-OpenCV) has no plugin/FFI interface: its boundary is the
 * executable `fpmMain <dataset.json> <itrCount>` and two C++ functions,
 *     int16_t loadFPMDataset(FPM_Dataset*);   fpmMain.h:118
 *     void    runFPM(FPM_Dataset*);           fpmMain.h:119
 * This header is what a maintainer's `runFPM()` binds instead of the cv::UMat/cvComplex
 * call sequence of fpmMain.cpp:274-498 (see INTEGRATION.md for the stub).  Plain C:
 * opaque handle, plain pointers and sizes, int status codes, no exceptions, no torch/OpenCV
 * types.  All host pointers are caller-owned; calls on one context are not re-entrant
 * (the reference is single-threaded, fpmMain.cpp:29-33); one context per CUDA device.
 *
 * Conventions
 *   Np      low-res tile edge (`FPM_Dataset::Np`, fpmMain.h:66): any even size in 8..1024 whose prime factors are
 *           2, 3, 5 (the shipped JSONs use 90, 100, 200; BASELINE.json's configs 64, 128, 256)
 *   Nlarge  high-res edge (`Nlarge == Mlarge`, fpmMain.h:70-71, fpmMain.cpp:564-565): even, Np..3584, factors 2, 3, 5
 *   objF    [Nlarge][Nlarge][2] float, DC-at-corner like `FPM_Dataset::objF` (fpmMain.h:92)
 *   objCrop [Nlarge][Nlarge][2] float = IDFT_scaled(objF) (`FPM_Dataset::objCrop`, fpmMain.cpp:481)
 *   pupil   [Np][Np][2] float, DC-at-corner like `FPM_Dataset::pupil` (fpmMain.h:94)
 *   stack   [n_leds][Np][Np] uint16 in UPDATE ORDER (`sortedIndicies`, fpmMain.cpp:246-258,350):
 *           slot k holds `imageStack[sortedIndicies[k]].Image` after loader preprocessing
 *   led tables: `cropXStart/cropYStart` (fpmMain.h:37-39) per slot, same order
 * There is no CPU fallback: every entry point fails with FPMB200_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef FPMB200_H
#define FPMB200_H 1
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpmb200_ctx fpmb200_ctx;

enum {
  FPMB200_OK = 0,
  FPMB200_ERR_ARG = -1,      /* bad argument / unsupported size */
  FPMB200_ERR_CUDA = -2,     /* CUDA runtime error (message in fpmb200_last_error) */
  FPMB200_ERR_STATE = -3     /* call order violated (e.g. run before upload) */
};

/* Last error message of the calling thread ("" if none). */
const char* fpmb200_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int fpmb200_abi_version(void);

/* One context per CUDA device ordinal (use_gpu.sh: OPENCV_OPENCL_DEVICE=GPU:<n>). */
int  fpmb200_create(int device, fpmb200_ctx** out);
void fpmb200_destroy(fpmb200_ctx* ctx);

/* Allocates device state for `n_tiles` independent tiles that share one LED geometry.
 * Replaces the UMat allocations of runFPM (fpmMain.cpp:282-297,330-332). */
int fpmb200_tiles_alloc(fpmb200_ctx* ctx, int n_tiles, int Np, int Nlarge, int n_leds);

/* delta1/delta2/eps of FPM_Dataset (fpmMain.h:88-89,99).  `literal_scalar` selects how the
 * float scalars of cv::add (fpmMain.cpp:390,417,469) broadcast: 1 = to both channels (what the
 * shipped source does against stock OpenCV: complex denominators), 0 = real part only. */
int fpmb200_set_params(fpmb200_ctx* ctx, float delta1, float delta2, float eps, int literal_scalar);

/* CTAs per tile of the update kernel: 0 = let the library choose (default: a cluster of 8 CTAs with the field in
 * distributed shared memory for Np=256; Np=128: 4 CTAs per tile while at most a quarter of the SMs' worth of tiles is
 * allocated, 2 up to half of them when the pupil box is narrow, one CTA per tile otherwise), 1 = one CTA per tile, 2 / 4 (Np=128) or
 * 8 (Np=256) = a thread-block cluster per tile (lower single-tile latency; fewer tiles in flight).  The
 * reference has no counterpart: it is the choice its OpenCL queue makes implicitly (fpmMain.cpp:345-476). */
int fpmb200_set_cluster(fpmb200_ctx* ctx, int ctas_per_tile);

/* Per-slot crop origins (fpmMain.cpp:157-165), already in update order.  Each must satisfy
 * 0 <= start <= Nlarge-Np. */
int fpmb200_upload_leds(fpmb200_ctx* ctx, const int16_t* cropXStart, const int16_t* cropYStart, int n_leds);

/* `FPM_Dataset::pupilSupport` real plane, [Np][Np], DC-at-corner (fpmMain.cpp:304-313). */
int fpmb200_upload_pupil_support(fpmb200_ctx* ctx, const float* mask);

/* Copies the intensity stacks of tiles [tile_first, tile_first+n) host->device on `stream`
 * (cudaStream_t or NULL = the context's stream).  Asynchronous when `stack` is pinned. */
int fpmb200_upload_stack(fpmb200_ctx* ctx, int tile_first, int n, const uint16_t* stack, void* stream);

/* Pupil + spectrum initialisation of fpmMain.cpp:301-343 for tiles [tile_first, tile_first+n):
 * pupil = support, objF = fftShift-placed FFT of sqrt(image of slot `init_led_slot`) filtered by
 * the support.  The reference uses slot 1 (`sortedIndicies.at(1)`, fpmMain.cpp:319). */
int fpmb200_init_tiles(fpmb200_ctx* ctx, int tile_first, int n, int init_led_slot, void* stream);

/* `iters` passes of the inner loops fpmMain.cpp:345-476 over all LED slots, sequential order
 * preserved, for tiles [tile_first, tile_first+n).  Asynchronous on `stream`.  Pass all iterations in
 * one call when there are more tiles than SMs: the run is then cut into balanced passes of one
 * iteration over at most sm_count tiles instead of leaving the last wave partly empty (same results,
 * bit for bit; FPMB200_RUN_BALANCED=0 keeps one launch). */
int fpmb200_run(fpmb200_ctx* ctx, int tile_first, int n, int iters, void* stream);

/* One sub-aperture update (fpmMain.cpp:350-475) of one tile; for per-step parity. */
int fpmb200_step(fpmb200_ctx* ctx, int tile, int led_slot);

/* objCrop = IDFT_scaled(objF) (fpmMain.cpp:481) for tiles [tile_first, tile_first+n). */
int fpmb200_finalize(fpmb200_ctx* ctx, int tile_first, int n, void* stream);

/* State injection / extraction (any pointer may be NULL).  Synchronous. */
int fpmb200_upload_state(fpmb200_ctx* ctx, int tile, const float* objF, const float* pupil);
int fpmb200_download(fpmb200_ctx* ctx, int tile, float* objF, float* objCrop, float* pupil);
/* Asynchronous gather of objCrop of tiles [tile_first,tile_first+n) into pinned host memory. */
int fpmb200_download_objcrop(fpmb200_ctx* ctx, int tile_first, int n, float* objCrop, void* stream);

/* Device pointer of tile `tile`'s buffer for zero-copy hand-off (e.g. the final NCCL gather of a
 * multi-GPU run): which = 0 centred spectrum [Nlarge][Nlarge][2] float, 1 objCrop, 2 pupil,
 * 3 intensity stack as the kernel keeps it (float 1/I in the permuted layout of csrc/fpm_update.cuh
 * `stack_offset`; natural order on the general path), 4 the uint16 stack as uploaded / ingested
 * ([n_leds][Np][Np]).  Consecutive tiles are contiguous. */
int fpmb200_device_buffer(fpmb200_ctx* ctx, int which, int tile, void** ptr, unsigned long long* bytes_per_tile);

/* ---- full field of view (the callers either side of the loop; the reference runs one tile per process) ----
 *
 * Frame ingest = fpmMain.cpp:109-144 for every tile at once: the camera frame of one LED is copied to the device
 * once, each tile's Np x Np ROI (origin set below = the reference's cropX/cropY of that tile) is cut, divided by
 * `divisor` when it is != 1 (darkfieldExpMultiplier of a dark-field LED; cv::divide: round half to even, x/0 = 0,
 * :128-129), and the frame's background value -- cv::mean of the two Np x Np ROIs at (bk1x,bk1y), (bk2x,bk2y),
 * averaged, clamped at bg_threshold, rounded (:131-140) -- is subtracted with saturation (:143-144).  The result
 * becomes LED slot `led_slot` of every tile's stack (the caller passes frames in update order, as for
 * fpmb200_upload_stack).  `frame` is a host pointer ([height][width] uint16; pinned for an asynchronous copy). */
int fpmb200_set_tile_origins(fpmb200_ctx* ctx, const int32_t* roi_x, const int32_t* roi_y, int n_tiles);
int fpmb200_ingest_frame(fpmb200_ctx* ctx, int led_slot, const uint16_t* frame, int width, int height, int divisor,
                         int bk1x, int bk1y, int bk2x, int bk2y, int bg_threshold, void* stream);
/* The same for a context that owns only some of the frame's tiles (several GPUs): `rows` points at frame row `row0`
 * ([n_rows][width] uint16, pinned for an asynchronous copy) and must cover every ROI of this context's tiles; the
 * frame's background value (`FPMimg::bg_val`, computed once per frame by the host: fpmhost / backgroundValue) is passed
 * in, because the two background ROIs need not lie inside these rows. */
int fpmb200_ingest_rows(fpmb200_ctx* ctx, int led_slot, const uint16_t* rows, int width, int row0, int n_rows, int divisor,
                        int bg_val, void* stream);
/* `FPMimg::bg_val` (fpmMain.h:26) of every ingested LED slot, [n_leds]; synchronises. */
int fpmb200_ingest_bg(fpmb200_ctx* ctx, int32_t* bg_val);

/* Amplitude mosaic of a regular nx x ny grid of tiles (tile index = iy*nx + ix, ROI origin = (x0 + ix*step,
 * y0 + iy*step), step <= Np): out[Hm][Wm] float with Wm = ((nx-1)*step + Np) * Nlarge/Np, Hm likewise; overlapping
 * tiles are cross-faded with separable linear ramps.  `tiles_device` = device pointer to nx*ny objCrop images
 * ([Nlarge][Nlarge][2] float each, e.g. the gathered result of several GPUs) or NULL for this context's own
 * tiles after fpmb200_finalize.  `out` is a device pointer if out_on_device, else host memory (async on `stream`). */
int fpmb200_mosaic(fpmb200_ctx* ctx, const void* tiles_device, int nx, int ny, int step, float* out, int out_on_device,
                   void* stream);

/* Final gather of a single-process multi-GPU run: a device buffer on `ctx`'s GPU, and an asynchronous (peer) copy of
 * the objCrop of tiles [tile_first, tile_first+n) of `src` into `dst_ptr` on `dst`'s GPU (ordered on `src`'s
 * stream, or `stream`).  The multi-process variant is an NCCL send/recv from fpmb200_device_buffer (bench.py). */
int fpmb200_device_alloc(fpmb200_ctx* ctx, unsigned long long bytes, void** ptr);
int fpmb200_device_free(fpmb200_ctx* ctx, void* ptr);
int fpmb200_copy_objcrop_to(fpmb200_ctx* src, int tile_first, int n, fpmb200_ctx* dst, void* dst_ptr, void* stream);

/* Page-locked host staging memory for the asynchronous copies above (frames for fpmb200_ingest_frame, stacks for
 * fpmb200_upload_stack, results of fpmb200_download_objcrop).  write_combined != 0 asks for write-combined pages:
 * faster for the device to read over PCIe on some hosts, slow for the CPU to read back -- input buffers only.
 * The reference has no counterpart (cv::UMat::getMat maps pageable memory, fpmMain.cpp:380-381). */
int fpmb200_host_alloc(unsigned long long bytes, int write_combined, void** ptr);
int fpmb200_host_free(void* ptr);

int fpmb200_sync(fpmb200_ctx* ctx);
/* Stream markers for callers that pipeline host buffers against the asynchronous calls above without owning a CUDA
 * runtime: record marker `slot` (0..63) behind everything enqueued so far on `stream` (NULL = the context's stream),
 * and block the calling thread until the work before the last record of `slot` is done (a never-recorded slot
 * returns at once). */
int fpmb200_event_record(fpmb200_ctx* ctx, int slot, void* stream);
int fpmb200_event_sync(fpmb200_ctx* ctx, int slot);

/* Introspection: kernels launched by this context so far, and the name/shape of the update
 * kernel variant selected for the current allocation (for logs and bench.py). */
long long   fpmb200_kernel_launches(const fpmb200_ctx* ctx);
const char* fpmb200_variant(const fpmb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FPMB200_H */
