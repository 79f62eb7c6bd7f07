/* fpmhost.h -- C ABI of the host-side dataset layer (libfpmhost.so; no CUDA dependency).
 *
 * Exposes, for tests and foreign-language hosts, what the reference computes in
 *   main()            fpmMain.cpp:512-584  (dataset*.json keys, derived optics)
 *   loadFPMDataset()  fpmMain.cpp:36-271   (LED geometry, NA filter, crop boxes, LED order,
 *                                           image crop / dark-field divide / background subtract)
 * The C++ interface with the reference's own names (FPM_Dataset, FPMimg, loadFPMDataset, runFPM)
 * is fpm-opencv_b200/host/fpm_dataset.h.
 */
#ifndef FPMHOST_H
#define FPMHOST_H 1
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpmhost_dataset fpmhost_dataset;

typedef struct fpmhost_scalars {      /* FPM_Dataset fields, fpmMain.h:43-101 */
  int32_t Np, Nlarge, Mlarge, resImprovementFactor, naRadius, ledCount, ledUsedCount;
  int32_t cropX, cropY, bk1cropX, bk1cropY, bk2cropX, bk2cropY, darkfieldExpMultiplier;
  int32_t flipIlluminationX, flipIlluminationY, color, itrCount, parse_ok;
  float ps_eff, du, lambda, objectiveNA, maxIlluminationNA, delta1, delta2, bgThreshold, eps, ps;
  double arrayRotation;
} fpmhost_scalars;

typedef struct fpmhost_led {          /* FPMimg fields, fpmMain.h:19-41 */
  int32_t led_num, used;
  double sinTheta_x, sinTheta_y;
  float uled, vled, illumination_na;
  int16_t idx_u, idx_v, cropXStart, cropXEnd, cropYStart, cropYEnd, bg_val, pad_;
} fpmhost_led;

const char* fpmhost_last_error(void);

/* readDatasetJson: parse `json_path` (argv[1]) with iteration count argv[2]. */
int  fpmhost_open(const char* json_path, int itr_count, fpmhost_dataset** out);
void fpmhost_close(fpmhost_dataset* ds);
/* Geometry + LED order as if image files existed for LED numbers first..last (no file I/O).
 * Returns ledUsedCount (>= 0) or a negative error. */
int  fpmhost_geometry(fpmhost_dataset* ds, int first_led, int last_led);
/* loadFPMDataset(): scans datasetRoot for <filePrefix><n><fileExtension> TIFFs. Returns the
 * reference's return value (1 / -1). */
int  fpmhost_load(fpmhost_dataset* ds);
int  fpmhost_get_scalars(const fpmhost_dataset* ds, fpmhost_scalars* out);
/* sortedIndicies (fpmMain.cpp:246-258); returns the count written. */
int  fpmhost_get_order(const fpmhost_dataset* ds, int16_t* order, int capacity);
int  fpmhost_get_led(const fpmhost_dataset* ds, int led_num, fpmhost_led* out);
/* imageStack[led_num].Image, [Np][Np] uint16 after preprocessing. */
int  fpmhost_get_image(const fpmhost_dataset* ds, int led_num, uint16_t* out);
const char* fpmhost_geometry_source(const fpmhost_dataset* ds);
/* pupilSupport real plane (fpmMain.cpp:304-313), DC-at-corner. */
int  fpmhost_pupil_support(int Np, int radius, float* mask);
/* The per-frame preprocessing of loadFPMDataset (fpmMain.cpp:124-144) on one [height][width] uint16 frame: ROI cut at
 * (cropX,cropY), cv::divide by `divisor` when it is != 1, background estimate from the two ROIs clamped at
 * bgThreshold, saturating subtraction.  out = [Np][Np]; *bg_val = FPMimg::bg_val. */
int  fpmhost_preprocess_frame(const uint16_t* frame, int width, int height, int Np, int cropX, int cropY, int bk1x,
                              int bk1y, int bk2x, int bk2y, int divisor, int bgThreshold, uint16_t* out, int* bg_val);
/* Regular tile grid over a width x height frame: tile (ix,iy) has its ROI at (ix*(Np-overlap), iy*(Np-overlap));
 * returns the counts (the reference has one cropX/cropY per run, fpmMain.cpp:532-533). */
int  fpmhost_tile_grid(int width, int height, int Np, int overlap, int* nx, int* ny);
/* CUDA ordinal encoded in OPENCV_OPENCL_DEVICE (use_gpu.sh), -1 for CPU:* (use_cpu.sh). */
int  fpmhost_device_from_env(void);

#ifdef __cplusplus
}
#endif
#endif
