// run_fov.cpp -- full-field-of-view reconstruction driven by one dataset JSON (north star: "full-FOV
// reconstructions are split into independent spatial tiles sharded across the GPUs of one box with no collective on
// the inner loop, only a final gather").  The reference reconstructs the single ROI cropX/cropY per process
// (fpmMain.cpp:519,532-533) and re-reads every frame for it (fpmMain.cpp:109-144); here every camera frame is read
// once, cut into all tiles on the device(s), each tile runs the same update kernel, and the objCrop tiles are
// gathered on the first GPU and blended into one amplitude mosaic.
#include <dirent.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

#include "../../include/fpmb200.h"
#include "fpm_dataset.h"
#include "tiff_io.h"

namespace {
void ck(int rc, const char* what) {
  if (rc != FPMB200_OK) throw std::runtime_error(std::string(what) + ": " + fpmb200_last_error());
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Dev {
  fpmb200_ctx* c = nullptr;
  int first = 0, n = 0;       // tile range [first, first+n) of the grid
  Dev() = default;
  Dev(const Dev&) = delete;
  Dev& operator=(const Dev&) = delete;
  ~Dev() { fpmb200_destroy(c); }            // every exit path (ck() throws) releases the device state
};
struct DevBuf {                              // device allocation owned by a context
  fpmb200_ctx* c = nullptr;
  void* p = nullptr;
  ~DevBuf() { if (p) fpmb200_device_free(c, p); }
};
}  // namespace

int runFPMFullFOV(FPM_Dataset* d, int overlap, const std::vector<int>& devices, const std::string& outDir) {
  const double t0 = now();
  const int Np = d->Np, L = d->Nlarge;
  if (overlap < 0 || overlap >= Np) throw std::runtime_error("tile overlap must be in [0, Np)");
  if (devices.empty()) throw std::runtime_error("no CUDA device selected");

  // ---- pass 1: LED geometry of every frame in the directory (fpmMain.cpp:63-106,146-177,246-258), no pixels ----
  allocateImageStack(d);
  DIR* dir = opendir(d->datasetRoot.c_str());
  if (dir == NULL) {
    std::cout << "ERROR: Could not Open Directory.\n";                                       // :268
    return -1;
  }
  std::vector<std::pair<int, std::string>> files;   // (led_num, file name)
  struct dirent* ent;
  while ((ent = readdir(dir)) != NULL) {
    std::string fileName = ent->d_name;
    const size_t el = d->fileExtension.length(), pl = d->filePrefix.length();
    if (fileName == "." || fileName == ".." || fileName.length() < el + pl) continue;
    if (fileName.compare(fileName.length() - el, el, d->fileExtension) != 0 || fileName.find(d->filePrefix) != 0) continue;
    const int led_num = atoi(fileName.substr(pl, fileName.length() - el - pl).c_str());
    FPMimg im;
    bool pass;
    try {
      pass = computeLedGeometry(*d, led_num, &im);
    } catch (const std::exception& e) {                                                      // like loadFPMDataset
      std::cout << "ERROR: " << e.what() << std::endl;
      closedir(dir);
      return -1;
    }
    if (!pass) {
      std::cout << "Skipped LED# " << led_num << std::endl;                                  // :236
      continue;
    }
    if (led_num < 0 || led_num > d->ledCount) {
      closedir(dir);
      throw std::runtime_error("LED # " + std::to_string(led_num) + " exceeds ledCount");
    }
    registerImage(d, im);
    files.push_back({led_num, fileName});
  }
  closedir(dir);
  d->ledUsedCount = (uint16_t)files.size();
  if (files.size() < 2) {
    std::cout << "ERROR - No images found in given directory." << std::endl;                 // :242
    return -1;
  }
  sortLedOrder(d);
  const int n = d->ledUsedCount;
  std::vector<int> slot_of(d->ledCount + 1, -1);
  std::vector<int16_t> cx(n), cy(n);
  for (int k = 0; k < n; ++k) {
    const FPMimg& im = d->imageStack.at(d->sortedIndicies.at(k));
    slot_of[im.led_num] = k;
    cx[k] = im.cropXStart;
    cy[k] = im.cropYStart;
  }

  // ---- frame size from the first file; tile grid; contexts ----
  fpmio::Image16 full;
  std::string err;
  if (!fpmio::readTiff(d->datasetRoot + files[0].second, full, &err)) throw std::runtime_error(err);
  const int W = full.width, H = full.height;
  if (W < Np || H < Np) throw std::runtime_error("frame smaller than one tile");
  int nx = 0, ny = 0;
  tileGrid(W, H, Np, overlap, &nx, &ny);
  const int step = Np - overlap, n_tiles = nx * ny;
  // never more devices than tiles: every context below owns at least one tile (devs[0] is the gather root)
  const int G = (int)devices.size() < n_tiles ? (int)devices.size() : n_tiles;
  std::cout << "Full FOV: " << W << "x" << H << " frame -> " << nx << "x" << ny << " tiles of " << Np << " (overlap " << overlap
            << "), " << n << " LEDs, " << G << " GPU(s)" << std::endl;
  {
    const int mx = W - ((nx - 1) * step + Np), my = H - ((ny - 1) * step + Np);
    if (mx || my) std::cout << "Full FOV: " << mx << " columns on the right and " << my << " rows at the bottom are not covered by the tile grid" << std::endl;
  }
  makePupilSupport(Np, d->naRadius, &d->pupilSupport);
  std::vector<Dev> devs(G);
  for (int g = 0; g < G; ++g) {
    Dev& v = devs[g];
    v.first = (int)((long long)n_tiles * g / G);
    v.n = (int)((long long)n_tiles * (g + 1) / G) - v.first;
    if (v.n == 0) continue;
    ck(fpmb200_create(devices[g], &v.c), "fpmb200_create");
    ck(fpmb200_tiles_alloc(v.c, v.n, Np, L, n), "fpmb200_tiles_alloc");
    ck(fpmb200_set_params(v.c, d->delta1, d->delta2, d->eps, d->literalScalar ? 1 : 0), "fpmb200_set_params");
    ck(fpmb200_upload_leds(v.c, cx.data(), cy.data(), n), "fpmb200_upload_leds");
    ck(fpmb200_upload_pupil_support(v.c, d->pupilSupport.data()), "fpmb200_upload_pupil_support");
    std::vector<int32_t> ox(v.n), oy(v.n);
    for (int t = 0; t < v.n; ++t) {
      ox[t] = ((v.first + t) % nx) * step;
      oy[t] = ((v.first + t) / nx) * step;
    }
    ck(fpmb200_set_tile_origins(v.c, ox.data(), oy.data(), v.n), "fpmb200_set_tile_origins");
    if (d->debug) std::cout << "GPU " << devices[g] << ": tiles [" << v.first << "," << v.first + v.n << ") " << fpmb200_variant(v.c) << std::endl;
  }

  // ---- pass 2: every frame is read once and cut into all tiles on the devices (fpmMain.cpp:109-144) ----
  std::cout << "Loading Images..." << std::endl;                                             // :65
  std::vector<uint16_t> plane;
  for (size_t f = 0; f < files.size(); ++f) {
    if (f > 0 && !fpmio::readTiff(d->datasetRoot + files[f].second, full, &err)) throw std::runtime_error(err);
    if (full.width != W || full.height != H) throw std::runtime_error(files[f].second + ": frame size differs from the first frame");
    const uint16_t* frame = full.pix.data();
    if (full.channels != 1) {
      if (!d->color) throw std::runtime_error(files[f].second + " has several channels but isColor is false");
      plane.resize((size_t)W * H);
      for (size_t k = 0; k < plane.size(); ++k) plane[k] = full.pix[k * full.channels];    // channels[2] of BGR (:112-115)
      frame = plane.data();
    }
    const FPMimg& im = d->imageStack.at(files[f].first);
    const int divisor = (d->darkfieldExpMultiplier != 1 && im.illumination_na > d->objectiveNA) ? d->darkfieldExpMultiplier : 1;
    for (Dev& v : devs)
      if (v.c) {
        ck(fpmb200_ingest_frame(v.c, slot_of[files[f].first], frame, W, H, divisor, d->bk1cropX, d->bk1cropY, d->bk2cropX,
                                d->bk2cropY, (int)d->bgThreshold, nullptr), "fpmb200_ingest_frame");
      }
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_sync(v.c), "fpmb200_sync");                                      // `full` is reused for the next file
    std::cout << "Loaded: " << files[f].second << ", LED # is: " << files[f].first << std::endl;   // :180-181
  }

  // ---- the loop (fpmMain.cpp:345-476) on every tile, all devices concurrently ----
  for (Dev& v : devs)
    if (v.c) ck(fpmb200_init_tiles(v.c, 0, v.n, 1, nullptr), "fpmb200_init_tiles");
  for (int16_t itr = 1; itr <= d->itrCount; itr++) {
    const double t1 = now();
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_run(v.c, 0, v.n, 1, nullptr), "fpmb200_run");
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_sync(v.c), "fpmb200_sync");
    d->secondsPerIteration = now() - t1;
    std::cout << "Iteration " << itr << " Completed (Time: " << (float)d->secondsPerIteration << " sec)" << std::endl;   // :479
  }
  for (Dev& v : devs)
    if (v.c) ck(fpmb200_finalize(v.c, 0, v.n, nullptr), "fpmb200_finalize");               // :481

  // ---- final gather on the first GPU + mosaic ----
  const int f = L / Np;
  const int Wm = ((nx - 1) * step + Np) * f, Hm = ((ny - 1) * step + Np) * f;
  std::vector<float> mosaic((size_t)Wm * Hm);
  Dev& root = devs[0];
  DevBuf gbuf;
  void*& gathered = gbuf.p;
  if (G > 1) {
    gbuf.c = root.c;
    ck(fpmb200_device_alloc(root.c, (unsigned long long)n_tiles * L * L * 8ull, &gathered), "fpmb200_device_alloc");
    for (Dev& v : devs)
      if (v.c)
        ck(fpmb200_copy_objcrop_to(v.c, 0, v.n, root.c, (char*)gathered + (size_t)v.first * L * L * 8, nullptr), "fpmb200_copy_objcrop_to");
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_sync(v.c), "fpmb200_sync");
  }
  ck(fpmb200_mosaic(root.c, gathered, nx, ny, step, mosaic.data(), 0, nullptr), "fpmb200_mosaic");
  ck(fpmb200_sync(root.c), "fpmb200_sync");
  d->secondsTotal = now() - t0;
  std::cout << "FP Processing Completed (Time: " << (float)d->secondsTotal << " sec)" << std::endl;                  // :489
  if (!outDir.empty()) {
    if (!fpmio::writeTiffF32(outDir + "/mosaic_amp.tif", mosaic.data(), Wm, Hm, &err)) std::cout << "ERROR: " << err << std::endl;
    else std::cout << "Wrote mosaic_amp.tif (" << Wm << "x" << Hm << ") to " << outDir << std::endl;
  }
  return 1;
}
