// fpmMain.cpp -- the reference's entry point, `./fpmMain <dataset.json> <itrCount>`
// (fpmMain.cpp:500-592), on the B200 path.  Same argument meaning, same dataset*.json keys, same
// stdout lines ("Dataset Root:", "resImprovementFactor:", "Loading Images...", "Loaded: ...",
// "Iteration k Completed (Time: ... sec)", "FP Processing Completed (Time: ... sec)").
// The reference shows the result in highgui windows (fpmMain.cpp:495-497); this build has no GUI
// and, when a third argument or FPM_OUTPUT_DIR names a directory, writes amplitude / phase of the
// object and of the (fftShifted) pupil there as 32-bit float TIFFs instead.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <string>
#include <vector>

#include "fpm_dataset.h"
#include "tiff_io.h"

static void writeAmpPhase(const std::string& dir, const std::string& name, const std::vector<float>& c, int n, bool shift) {
  std::vector<float> amp((size_t)n * n), ph((size_t)n * n);
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      const int sy = shift ? (y + n / 2) % n : y, sx = shift ? (x + n / 2) % n : x;
      const float re = c[2 * ((size_t)sy * n + sx)], im = c[2 * ((size_t)sy * n + sx) + 1];
      amp[(size_t)y * n + x] = std::sqrt(re * re + im * im);
      ph[(size_t)y * n + x] = std::atan2(im, re);
    }
  std::string err;
  if (!fpmio::writeTiffF32(dir + "/" + name + "_amp.tif", amp.data(), n, n, &err) ||
      !fpmio::writeTiffF32(dir + "/" + name + "_phase.tif", ph.data(), n, n, &err))
    std::cout << "ERROR: " << err << std::endl;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cout << "ERROR: Not enough input argumants.\n Usage: ./fpmMain dataset.json" << std::endl;   // :502-505
    return 0;
  }
  FPM_Dataset mDataset;
  try {
    readDatasetJson(argv[1], atoi(argv[2]), &mDataset);

    std::cout << "Dataset Root: " << mDataset.datasetRoot << std::endl;                              // :541
    char fileName[129];
    snprintf(fileName, sizeof fileName, "%s%04d%s", mDataset.filePrefix.c_str(), mDataset.centerLED,
             mDataset.fileExtension.c_str());
    std::cout << mDataset.datasetRoot + fileName << std::endl;                                       // :546
    std::cout << "resImprovementFactor: " << mDataset.resImprovementFactor << std::endl;             // :560
    std::cout << "LED geometry: " << mDataset.geometrySource << std::endl;

    if (mDataset.cudaDevice < 0) {
      std::cout << "ERROR: OPENCV_OPENCL_DEVICE=CPU:* (use_cpu.sh) -- this build has no CPU reconstruction path.\n"
                   "       The same-host CPU baseline lives in oracle/ (python bench.py --impl reference)." << std::endl;
      return 3;
    }
    // Full field of view (not in the reference, which reconstructs the one ROI cropX/cropY): FPM_FOV_OVERLAP=<pixels>
    // tiles the whole frame; FPM_GPUS=0,1,... shards the tiles over several GPUs of the box.
    if (const char* fov = getenv("FPM_FOV_OVERLAP")) {
      std::vector<int> devices;
      if (const char* gl = getenv("FPM_GPUS")) {
        for (const char* q = gl; *q;) {
          devices.push_back(atoi(q));
          while (*q && *q != ',') ++q;
          if (*q == ',') ++q;
        }
      }
      if (devices.empty()) devices.push_back(mDataset.cudaDevice);
      const char* out = argc > 3 ? argv[3] : getenv("FPM_OUTPUT_DIR");
      return runFPMFullFOV(&mDataset, atoi(fov), devices, out ? out : "") > 0 ? 0 : 1;
    }
    if (loadFPMDataset(&mDataset) > 0) {                                                             // :590-591
      runFPM(&mDataset);
      const char* out = argc > 3 ? argv[3] : getenv("FPM_OUTPUT_DIR");
      if (out && *out) {
        writeAmpPhase(out, "object", mDataset.objCrop, mDataset.Nlarge, false);
        writeAmpPhase(out, "pupil", mDataset.pupil, mDataset.Np, true);                              // :496 fftShift(pupil)
        std::cout << "Wrote object_amp/phase.tif, pupil_amp/phase.tif to " << out << std::endl;
      }
    } else {
      return 1;
    }
  } catch (const std::exception& e) {
    std::cout << "ERROR: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
