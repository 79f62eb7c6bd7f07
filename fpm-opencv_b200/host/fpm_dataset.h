// fpm_dataset.h -- host-side mirror of the reference's data model and entry points.
//
// Same class and member names as fpmMain.h:19-101 so that code written against the reference
// reads the same; cv::UMat members become plain host vectors (the device state lives behind the
// C ABI of include/fpmb200.h), and the four per-LED scratch mats of FPMimg (Objfcrop, ObjfcropP,
// ObjcropP, Objfup; fpmMain.h:22-25) are gone because the fused kernel never materialises them.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "mini_json.h"

class FPMimg {                      // fpmMain.h:19-41
 public:
  std::vector<uint16_t> Image;      // [Np][Np] after ROI crop / dark-field divide / background subtract
  int led_num = 0;
  double sinTheta_x = 0, sinTheta_y = 0;
  float vled = 0, uled = 0;
  int16_t idx_u = 0, idx_v = 0;
  float illumination_na = 0;
  int16_t bg_val = 0;
  int16_t pupilShiftX = 0, pupilShiftY = 0;
  int16_t cropYStart = 0, cropYEnd = 0, cropXStart = 0, cropXEnd = 0;
};

class FPM_Dataset {                 // fpmMain.h:43-101
 public:
  std::string holeCoordinateFileName;
  mjson::Value holeCoordinates;
  std::string datasetRoot, filePrefix, fileExtension;
  std::vector<FPMimg> imageStack;
  double arrayRotation = 0;
  uint16_t darkfieldExpMultiplier = 1;
  int16_t holeNumberDigits = 4;
  uint16_t ledCount = 508, ledUsedCount = 0;
  float pixelSize = 0, objectiveMag = 0, objectiveNA = 0, maxIlluminationNA = 0, lambda = 0;
  bool color = false, leadingZeros = false;
  int16_t centerLED = 249;
  int16_t cropX = 0, cropY = 0, Np = 0, Np_padded = 0, Mcrop = 0, Ncrop = 0, Nlarge = 0, Mlarge = 0;
  float du = 0;
  std::vector<float> NALedPatternStackX, NALedPatternStackY, illuminationNAList;
  std::vector<float> sortedNALedPatternStackX, sortedNALedPatternStackY;
  std::vector<int16_t> sortedIndicies;
  int16_t bk1cropX = 0, bk1cropY = 0, bk2cropX = 0, bk2cropY = 0;
  float bgThreshold = 0, ps_eff = 0, ps = 0;
  float delta1 = 0, delta2 = 0;
  std::vector<float> objCrop;       // [Nlarge][Mlarge][2]  (cv::UMat objCrop, fpmMain.h:91)
  std::vector<float> objF;          // [Nlarge][Mlarge][2], DC-at-corner (fpmMain.h:92)
  std::vector<float> pupil;         // [Np][Np][2], DC-at-corner; runFPM leaves it fftShifted (fpmMain.cpp:496)
  std::vector<float> pupilSupport;  // [Np][Np] real plane, DC-at-corner (fpmMain.h:95)
  int16_t itrCount = 10;
  bool flipIlluminationX = false, flipIlluminationY = false;
  float eps = 0.0000000001;

  // ---- additions of the B200 build (not in the reference) ----
  int16_t resImprovementFactor = 0;  // local in the reference's main(), fpmMain.cpp:556
  int16_t naRadius = 0;              // local in runFPM, fpmMain.cpp:305
  bool debug = false;
  bool literalScalar = true;         // cv::add(float lvalue) broadcasts to both channels
  int cudaDevice = 0;                // from OPENCV_OPENCL_DEVICE=GPU:<n> (use_gpu.sh)
  std::vector<std::array<float, 3>> ledXYZ;   // geometry when it does not come from `holeCoordinates`
  std::string geometrySource;        // "holeCoordinates" | "ledList:<file>" | "holePositions" | "domeHoleCoordinates"
  std::string jsonDir;               // directory of the dataset JSON (relative geometry file lookup)
  double secondsPerIteration = 0, secondsTotal = 0;
};

// main()'s configuration block, fpmMain.cpp:512-584 (argv[1] = path, argv[2] = itrCount).
// Returns false only when the JSON cannot be opened at all (the reference would read defaults).
bool readDatasetJson(const std::string& path, int itrCount, FPM_Dataset* dataset);

// LED direction -> NA filter -> k-space crop box for one LED number (fpmMain.cpp:77-106,146-168).
// Returns true when the LED passes `illumination_na < maxIlluminationNA`; throws
// std::runtime_error where the reference would throw (no usable geometry).
bool computeLedGeometry(const FPM_Dataset& dataset, int led_num, FPMimg* img);

// Registers `img` like fpmMain.cpp:171-177 (imageStack / NA lists, indexed by LED number).
void registerImage(FPM_Dataset* dataset, const FPMimg& img);
// fpmMain.cpp:52-57: ledCount+1 placeholder entries (index 0 is a dummy).
void allocateImageStack(FPM_Dataset* dataset);
// fpmMain.cpp:246-258: unstable std::sort by illumination NA, first ledUsedCount kept.
void sortLedOrder(FPM_Dataset* dataset);
// ROI crop, dark-field exposure divide, two-ROI background estimate and saturating subtraction
// (fpmMain.cpp:124-144) on a full frame [h][w] (already reduced to one channel).
bool preprocessFrame(const FPM_Dataset& dataset, const uint16_t* frame, int w, int h, FPMimg* img, std::string* err);
// FPMimg::bg_val of a frame (fpmMain.cpp:131-140): cv::mean of the two Np x Np background ROIs, averaged, clamped at
// bgThreshold, rounded.  The ROIs must lie inside the frame (preprocessFrame checks).
int16_t backgroundValue(const FPM_Dataset& dataset, const uint16_t* frame, int w);

int16_t loadFPMDataset(FPM_Dataset* dataset);   // fpmMain.h:118
void runFPM(FPM_Dataset* dataset);              // fpmMain.h:119

// filled disc of cv::circle(centre (Np/2,Np/2), radius, filled) then fftShift (fpmMain.cpp:304-313)
// regular tile grid over a frame: counts of tiles with ROI origin (ix*(Np-overlap), iy*(Np-overlap))
void tileGrid(int width, int height, int Np, int overlap, int* nx, int* ny);
void makePupilSupport(int Np, int radius, std::vector<float>* mask);
// parses OPENCV_OPENCL_DEVICE as exported by use_gpu.sh / use_cpu.sh: returns CUDA ordinal >= 0,
// -1 for "CPU:*" (not served by this build), 0 when unset.
int deviceFromEnv();

// Full field of view (host/run_fov.cpp): every frame of the dataset directory is cut into a regular grid of Np x Np
// tiles (`overlap` pixels shared between neighbours), the tiles are sharded over `devices` (CUDA ordinals), and the
// result is one amplitude mosaic (written to outDir/mosaic_amp.tif when outDir is not empty).  Returns 1 or -1 like
// loadFPMDataset.
int runFPMFullFOV(FPM_Dataset* dataset, int overlap, const std::vector<int>& devices, const std::string& outDir);

