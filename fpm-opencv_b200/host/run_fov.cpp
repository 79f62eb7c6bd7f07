// run_fov.cpp -- full-field-of-view reconstruction driven by one dataset JSON (north star: "full-FOV
// reconstructions are split into independent spatial tiles sharded across the GPUs of one box with no collective on
// the inner loop, only a final gather").  The reference reconstructs the single ROI cropX/cropY per process
// (fpmMain.cpp:519,532-533) and re-reads every frame for it (fpmMain.cpp:109-144); here every camera frame is read
// once, cut into all tiles on the device(s), each tile runs the same update kernel, and the objCrop tiles are
// gathered on the first GPU and blended into one amplitude mosaic.
#include <dirent.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
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
  int row0 = 0, rows = 0;     // frame rows [row0, row0+rows) cover the ROIs of these tiles
  Dev() = default;
  Dev(const Dev&) = delete;
  Dev& operator=(const Dev&) = delete;
  ~Dev() { fpmb200_destroy(c); }            // every exit path (ck() throws) releases the device state
};
struct PinnedRing {                          // K page-locked frame buffers (fpmb200_host_alloc)
  std::vector<void*> p;
  PinnedRing(int k, size_t bytes) : p(k, nullptr) {
    for (int i = 0; i < k; ++i) ck(fpmb200_host_alloc(bytes, 0, &p[i]), "fpmb200_host_alloc");
  }
  PinnedRing(const PinnedRing&) = delete;
  ~PinnedRing() { for (void* q : p) fpmb200_host_free(q); }
  uint16_t* slot(int i) { return (uint16_t*)p[i]; }
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

  // ---- frame size from the first file's header; tile grid; contexts ----
  std::string err;
  std::vector<uint8_t> scratch0;
  int W = 0, H = 0, ch0 = 0;
  if (!fpmio::readTiffPlane(d->datasetRoot + files[0].second, nullptr, 0, &W, &H, &ch0, scratch0, &err)) throw std::runtime_error(err);
  if (W < Np || H < Np) throw std::runtime_error("frame smaller than one tile");
  {
    auto inside = [&](int x, int y) { return x >= 0 && y >= 0 && x + Np <= W && y + Np <= H; };
    if (!inside(d->bk1cropX, d->bk1cropY) || !inside(d->bk2cropX, d->bk2cropY))
      throw std::runtime_error("a background ROI leaves the " + std::to_string(W) + "x" + std::to_string(H) + " frame");
  }
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
  // staging memory: a ring of page-locked frame buffers for the reader threads and the mosaic's landing buffer --
  // allocated (cudaHostAlloc is slow: ~0.3 ms per MB) while the devices are being set up
  const int n_files = (int)files.size();
  int n_readers = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("FPM_READERS")) n_readers = atoi(e);
  n_readers = std::max(1, std::min(std::min(n_readers, 8), n_files));
  const int K = std::min(n_readers + 4, 32);                     // ring slots (= marker slots, < 64): readers + frames in flight
  const size_t frame_elems = (size_t)W * H;
  const int fz = L / Np;
  const int Wm = ((nx - 1) * step + Np) * fz, Hm = ((ny - 1) * step + Np) * fz;
  std::unique_ptr<PinnedRing> ring_p, mosaic_p;
  double t_pin = 0, t_geom = now() - t0;
  {
    // one setup thread per device: CUDA context creation and the allocations of the devices overlap
    std::vector<std::string> errs(G);
    std::string pin_err;
    std::vector<std::thread> th;
    th.emplace_back([&] {
      try {
        const double tp = now();
        ring_p.reset(new PinnedRing(K, frame_elems * sizeof(uint16_t)));
        mosaic_p.reset(new PinnedRing(1, (size_t)Wm * Hm * sizeof(float)));
        t_pin = now() - tp;
      } catch (const std::exception& e) { pin_err = e.what(); }
    });
    for (int g = 0; g < G; ++g)
      th.emplace_back([&, g] {
        try {
          Dev& v = devs[g];
          v.first = (int)((long long)n_tiles * g / G);
          v.n = (int)((long long)n_tiles * (g + 1) / G) - v.first;
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
          // the frame rows this device's tiles need: only those are sent to it
          v.row0 = oy[0];
          v.rows = oy[v.n - 1] + Np - v.row0;
          ck(fpmb200_set_tile_origins(v.c, ox.data(), oy.data(), v.n), "fpmb200_set_tile_origins");
        } catch (const std::exception& e) {
          errs[g] = e.what();
        }
      });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
      if (!errs[g].empty()) throw std::runtime_error("GPU " + std::to_string(devices[g]) + ": " + errs[g]);
    if (!pin_err.empty()) throw std::runtime_error(pin_err);
    if (d->debug)
      for (int g = 0; g < G; ++g)
        std::cout << "GPU " << devices[g] << ": tiles [" << devs[g].first << "," << devs[g].first + devs[g].n << ") rows [" << devs[g].row0
                  << "," << devs[g].row0 + devs[g].rows << ") " << fpmb200_variant(devs[g].c) << std::endl;
  }
  const double t_setup = now() - t0;

  // ---- pass 2 (fpmMain.cpp:109-144 for every tile at once): a pool of reader threads fills a ring of page-locked
  //      frame buffers (pread straight into them), the main thread hands every frame, in file order, to the devices --
  //      each gets only the rows of its tiles, asynchronously -- and a buffer returns to the readers when every device
  //      has passed the marker recorded behind its copy.  Reading, host-to-device copies and the cut / divide /
  //      background kernels overlap; every frame is read once, whatever the number of tiles.
  std::cout << "Loading Images..." << std::endl;                                             // :65
  const double t_load0 = now();
  PinnedRing& ring = *ring_p;
  std::vector<int16_t> frame_bg(n_files, 0);
  std::vector<std::string> frame_err(n_files);
  std::vector<char> state(n_files, 0);                           // 0 = not read, 1 = ready, 2 = failed
  std::mutex mu;
  std::condition_variable cv;
  int slots_released = K;                                        // frames [0, slots_released) may be written by the readers
  std::atomic<int> next_file{0};
  bool abort_all = false;
  auto reader = [&]() {
    std::vector<uint8_t> scratch;
    for (;;) {
      const int f = next_file.fetch_add(1);
      if (f >= n_files) return;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return f < slots_released || abort_all; });
        if (abort_all) return;
      }
      uint16_t* buf = ring.slot(f % K);
      int w = 0, h = 0, ch = 0;
      std::string e;
      bool ok = fpmio::readTiffPlane(d->datasetRoot + files[f].second, buf, frame_elems, &w, &h, &ch, scratch, &e);
      if (ok && (w != W || h != H)) { ok = false; e = files[f].second + ": frame size differs from the first frame"; }
      if (ok && ch != 1 && !d->color) { ok = false; e = files[f].second + " has several channels but isColor is false"; }
      if (ok) frame_bg[f] = backgroundValue(*d, buf, W);                                   // :131-140, once per frame
      {
        std::lock_guard<std::mutex> lk(mu);
        frame_err[f] = e;
        state[f] = ok ? 1 : 2;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> pool;
  for (int r = 0; r < n_readers; ++r) pool.emplace_back(reader);
  struct Joiner {                                                // every exit path stops and joins the readers
    std::vector<std::thread>& pool; std::mutex& mu; std::condition_variable& cv; bool& abort_all;
    ~Joiner() {
      { std::lock_guard<std::mutex> lk(mu); abort_all = true; }
      cv.notify_all();
      for (auto& t : pool) if (t.joinable()) t.join();
    }
  } joiner{pool, mu, cv, abort_all};
  for (int f = 0; f < n_files; ++f) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return state[f] != 0; });
      if (state[f] == 2) throw std::runtime_error(frame_err[f]);
    }
    const FPMimg& im = d->imageStack.at(files[f].first);
    const int divisor = (d->darkfieldExpMultiplier != 1 && im.illumination_na > d->objectiveNA) ? d->darkfieldExpMultiplier : 1;
    const uint16_t* frame = ring.slot(f % K);
    for (Dev& v : devs) {
      ck(fpmb200_ingest_rows(v.c, slot_of[files[f].first], frame + (size_t)v.row0 * W, W, v.row0, v.rows, divisor, frame_bg[f], nullptr),
         "fpmb200_ingest_rows");
      ck(fpmb200_event_record(v.c, f % K, nullptr), "fpmb200_event_record");
    }
    // Two frames of device work stay in flight; once every device has passed the marker behind frame g = f - 2, the
    // frames up to g are consumed (stream order) and the readers may overwrite their slots with frames up to g + K.
    const int g = f - 2;
    if (g >= 0) {
      for (Dev& v : devs) ck(fpmb200_event_sync(v.c, g % K), "fpmb200_event_sync");
      { std::lock_guard<std::mutex> lk(mu); slots_released = g + K + 1; }
      cv.notify_all();
    }
    std::cout << "Loaded: " << files[f].second << ", LED # is: " << files[f].first << std::endl;   // :180-181
  }
  for (Dev& v : devs) ck(fpmb200_sync(v.c), "fpmb200_sync");
  const double t_load = now() - t_load0;

  // ---- the loop (fpmMain.cpp:345-476) on every tile, all devices concurrently ----
  const double t_rec0 = now();
  for (Dev& v : devs)
    if (v.c) ck(fpmb200_init_tiles(v.c, 0, v.n, 1, nullptr), "fpmb200_init_tiles");
  // All iterations in one call per device: with more tiles than SMs the library re-cuts the run into balanced passes
  // (fpmb200_run), which a call per iteration would defeat (320 tiles = 3 partly empty waves per iteration).  The
  // reference's per-iteration lines (fpmMain.cpp:479) are printed afterwards with the mean time per iteration.
  {
    const double t1 = now();
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_run(v.c, 0, v.n, d->itrCount, nullptr), "fpmb200_run");
    for (Dev& v : devs)
      if (v.c) ck(fpmb200_sync(v.c), "fpmb200_sync");
    d->secondsPerIteration = d->itrCount > 0 ? (now() - t1) / d->itrCount : 0.0;
    for (int16_t itr = 1; itr <= d->itrCount; itr++)
      std::cout << "Iteration " << itr << " Completed (Time: " << (float)d->secondsPerIteration << " sec)" << std::endl;   // :479
  }
  for (Dev& v : devs)
    if (v.c) ck(fpmb200_finalize(v.c, 0, v.n, nullptr), "fpmb200_finalize");               // :481

  for (Dev& v : devs) ck(fpmb200_sync(v.c), "fpmb200_sync");
  const double t_rec = now() - t_rec0;

  // ---- final gather on the first GPU + mosaic ----
  const double t_mos0 = now();
  float* mosaic = (float*)mosaic_p->slot(0);
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
  ck(fpmb200_mosaic(root.c, gathered, nx, ny, step, mosaic, 0, nullptr), "fpmb200_mosaic");
  ck(fpmb200_sync(root.c), "fpmb200_sync");
  const double t_mos = now() - t_mos0;
  d->secondsTotal = now() - t0;
  std::cout << "FP Processing Completed (Time: " << (float)d->secondsTotal << " sec)" << std::endl;                  // :489
  std::cout << "Full FOV setup: directory scan + LED geometry " << (float)t_geom << " s, page-locked staging (" << K << " frame buffers + mosaic) "
            << (float)t_pin << " s in parallel with " << G << " device contexts + allocations" << std::endl;
  std::cout << "Full FOV timing: geometry+device setup " << (float)t_setup << " s, load+ingest " << (float)t_load << " s (" << n_readers << " reader threads), reconstruction "
            << (float)t_rec << " s, gather+mosaic " << (float)t_mos << " s, " << n_tiles << " tiles x " << n << " LEDs x "
            << d->itrCount << " iterations = " << (long long)n_tiles * n * d->itrCount << " updates" << std::endl;
  if (!outDir.empty()) {
    if (!fpmio::writeTiffF32(outDir + "/mosaic_amp.tif", mosaic, Wm, Hm, &err)) std::cout << "ERROR: " << err << std::endl;
    else std::cout << "Wrote mosaic_amp.tif (" << Wm << "x" << Hm << ") to " << outDir << std::endl;
  }
  return 1;
}
