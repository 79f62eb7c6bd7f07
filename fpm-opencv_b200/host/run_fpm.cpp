// run_fpm.cpp -- runFPM(): the reference's reconstruction entry point (fpmMain.h:119,
// fpmMain.cpp:274-498) as a thin wrapper that packs FPM_Dataset into the C ABI of
// include/fpmb200.h.  All arithmetic runs in the fused sm_100a kernel; nothing here computes.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <stdexcept>

#include "../../include/fpmb200.h"
#include "fpm_dataset.h"

namespace {
struct Ctx {
  fpmb200_ctx* c = nullptr;
  ~Ctx() { fpmb200_destroy(c); }
};
void ck(int rc, const char* what) {
  if (rc != FPMB200_OK) throw std::runtime_error(std::string(what) + ": " + fpmb200_last_error());
}
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

void runFPM(FPM_Dataset* dataset) {
  const double t3 = now();
  if (dataset->cudaDevice < 0)
    throw std::runtime_error(
        "OPENCV_OPENCL_DEVICE selects CPU (use_cpu.sh): this build has no CPU reconstruction path; "
        "the same-host CPU baseline is oracle/ (python bench.py --impl reference)");
  const int Np = dataset->Np, L = dataset->Nlarge, n = dataset->ledUsedCount;
  if (n < 2) throw std::runtime_error("need at least 2 images: the spectrum is seeded from sortedIndicies.at(1) (fpmMain.cpp:319)");

  // pupil support (fpmMain.cpp:301-313)
  makePupilSupport(Np, dataset->naRadius, &dataset->pupilSupport);

  // pack per-slot tables and the intensity stack in update order (fpmMain.cpp:348-355)
  std::vector<int16_t> cx(n), cy(n);
  std::vector<uint16_t> stack((size_t)n * Np * Np);
  for (int k = 0; k < n; ++k) {
    const FPMimg& im = dataset->imageStack.at(dataset->sortedIndicies.at(k));
    if (im.Image.size() != (size_t)Np * Np) throw std::runtime_error("image of LED " + std::to_string(im.led_num) + " missing");
    cx[k] = im.cropXStart;
    cy[k] = im.cropYStart;
    std::copy(im.Image.begin(), im.Image.end(), stack.begin() + (size_t)k * Np * Np);
  }

  Ctx h;
  ck(fpmb200_create(dataset->cudaDevice, &h.c), "fpmb200_create");
  ck(fpmb200_tiles_alloc(h.c, 1, Np, L, n), "fpmb200_tiles_alloc");
  ck(fpmb200_set_params(h.c, dataset->delta1, dataset->delta2, dataset->eps, dataset->literalScalar ? 1 : 0), "fpmb200_set_params");
  ck(fpmb200_upload_leds(h.c, cx.data(), cy.data(), n), "fpmb200_upload_leds");
  ck(fpmb200_upload_pupil_support(h.c, dataset->pupilSupport.data()), "fpmb200_upload_pupil_support");
  ck(fpmb200_upload_stack(h.c, 0, 1, stack.data(), nullptr), "fpmb200_upload_stack");
  ck(fpmb200_init_tiles(h.c, 0, 1, 1, nullptr), "fpmb200_init_tiles");     // slot 1 = sortedIndicies.at(1)
  ck(fpmb200_sync(h.c), "fpmb200_sync");
  if (dataset->debug) std::cout << "B200 kernel: " << fpmb200_variant(h.c) << std::endl;

  for (int16_t itr = 1; itr <= dataset->itrCount; itr++) {                   // fpmMain.cpp:345
    const double t1 = now();
    ck(fpmb200_run(h.c, 0, 1, 1, nullptr), "fpmb200_run");
    ck(fpmb200_sync(h.c), "fpmb200_sync");
    const float diff = (float)(now() - t1);
    dataset->secondsPerIteration = diff;
    std::cout << "Iteration " << itr << " Completed (Time: " << diff << " sec)" << std::endl;   // :479
  }
  ck(fpmb200_finalize(h.c, 0, 1, nullptr), "fpmb200_finalize");              // :481 (only the last one matters)
  dataset->objF.resize((size_t)L * L * 2);
  dataset->objCrop.resize((size_t)L * L * 2);
  dataset->pupil.resize((size_t)Np * Np * 2);
  ck(fpmb200_download(h.c, 0, dataset->objF.data(), dataset->objCrop.data(), dataset->pupil.data()), "fpmb200_download");
  const float diff = (float)(now() - t3);
  dataset->secondsTotal = diff;
  std::cout << "FP Processing Completed (Time: " << diff << " sec)" << std::endl;                 // :489
}
