// fpmhost_c.cpp -- C ABI (include/fpmhost.h) over fpm_dataset.h.
#include <cstring>
#include <exception>
#include <string>

#include "../../include/fpmhost.h"
#include "fpm_dataset.h"

struct fpmhost_dataset {
  FPM_Dataset d;
  bool parse_ok = false;
};

static thread_local std::string g_err;
static int fail(int rc, const std::string& m) { g_err = m; return rc; }

extern "C" const char* fpmhost_last_error(void) { return g_err.c_str(); }

extern "C" int fpmhost_open(const char* path, int itr, fpmhost_dataset** out) {
  if (!path || !out) return fail(-1, "NULL argument");
  *out = nullptr;
  try {
    auto* h = new fpmhost_dataset();
    h->parse_ok = readDatasetJson(path, itr, &h->d);
    if (!h->parse_ok) { delete h; return fail(-2, std::string("cannot open ") + path); }
    *out = h;
    return 0;
  } catch (const std::exception& e) { return fail(-3, e.what()); }
}

extern "C" void fpmhost_close(fpmhost_dataset* h) { delete h; }

extern "C" int fpmhost_geometry(fpmhost_dataset* h, int first, int last) {
  if (!h) return fail(-1, "NULL dataset");
  try {
    FPM_Dataset& d = h->d;
    allocateImageStack(&d);
    int n = 0;
    for (int led = first; led <= last; ++led) {
      FPMimg im;
      if (computeLedGeometry(d, led, &im)) { registerImage(&d, im); ++n; }
    }
    d.ledUsedCount = (uint16_t)n;
    sortLedOrder(&d);
    return n;
  } catch (const std::exception& e) { return fail(-3, e.what()); }
}

extern "C" int fpmhost_load(fpmhost_dataset* h) {
  if (!h) return fail(-1, "NULL dataset");
  try { return loadFPMDataset(&h->d); } catch (const std::exception& e) { return fail(-3, e.what()); }
}

extern "C" int fpmhost_get_scalars(const fpmhost_dataset* h, fpmhost_scalars* o) {
  if (!h || !o) return fail(-1, "NULL argument");
  const FPM_Dataset& d = h->d;
  memset(o, 0, sizeof *o);
  o->Np = d.Np; o->Nlarge = d.Nlarge; o->Mlarge = d.Mlarge; o->resImprovementFactor = d.resImprovementFactor;
  o->naRadius = d.naRadius; o->ledCount = d.ledCount; o->ledUsedCount = d.ledUsedCount;
  o->cropX = d.cropX; o->cropY = d.cropY; o->bk1cropX = d.bk1cropX; o->bk1cropY = d.bk1cropY;
  o->bk2cropX = d.bk2cropX; o->bk2cropY = d.bk2cropY; o->darkfieldExpMultiplier = d.darkfieldExpMultiplier;
  o->flipIlluminationX = d.flipIlluminationX; o->flipIlluminationY = d.flipIlluminationY; o->color = d.color;
  o->itrCount = d.itrCount; o->parse_ok = h->parse_ok;
  o->ps_eff = d.ps_eff; o->du = d.du; o->lambda = d.lambda; o->objectiveNA = d.objectiveNA;
  o->maxIlluminationNA = d.maxIlluminationNA; o->delta1 = d.delta1; o->delta2 = d.delta2;
  o->bgThreshold = d.bgThreshold; o->eps = d.eps; o->ps = d.ps; o->arrayRotation = d.arrayRotation;
  return 0;
}

extern "C" int fpmhost_get_order(const fpmhost_dataset* h, int16_t* order, int cap) {
  if (!h || !order) return fail(-1, "NULL argument");
  int n = (int)h->d.sortedIndicies.size();
  if (n > cap) n = cap;
  for (int k = 0; k < n; ++k) order[k] = h->d.sortedIndicies[k];
  return n;
}

extern "C" int fpmhost_get_led(const fpmhost_dataset* h, int led, fpmhost_led* o) {
  if (!h || !o) return fail(-1, "NULL argument");
  if (led < 0 || (size_t)led >= h->d.imageStack.size()) return fail(-1, "LED number out of range");
  const FPMimg& im = h->d.imageStack[led];
  memset(o, 0, sizeof *o);
  o->led_num = im.led_num; o->used = h->d.illuminationNAList[led] != 99.0f;
  o->sinTheta_x = im.sinTheta_x; o->sinTheta_y = im.sinTheta_y; o->uled = im.uled; o->vled = im.vled;
  o->illumination_na = im.illumination_na; o->idx_u = im.idx_u; o->idx_v = im.idx_v;
  o->cropXStart = im.cropXStart; o->cropXEnd = im.cropXEnd; o->cropYStart = im.cropYStart; o->cropYEnd = im.cropYEnd;
  o->bg_val = im.bg_val;
  return 0;
}

extern "C" int fpmhost_get_image(const fpmhost_dataset* h, int led, uint16_t* out) {
  if (!h || !out) return fail(-1, "NULL argument");
  if (led < 0 || (size_t)led >= h->d.imageStack.size()) return fail(-1, "LED number out of range");
  const auto& im = h->d.imageStack[led].Image;
  if (im.size() != (size_t)h->d.Np * h->d.Np) return fail(-2, "no image loaded for this LED");
  memcpy(out, im.data(), im.size() * sizeof(uint16_t));
  return 0;
}

extern "C" const char* fpmhost_geometry_source(const fpmhost_dataset* h) { return h ? h->d.geometrySource.c_str() : ""; }

extern "C" int fpmhost_pupil_support(int Np, int radius, float* mask) {
  if (!mask || Np <= 0) return fail(-1, "bad argument");
  std::vector<float> m;
  makePupilSupport(Np, radius, &m);
  memcpy(mask, m.data(), m.size() * sizeof(float));
  return 0;
}

extern "C" int fpmhost_preprocess_frame(const uint16_t* frame, int width, int height, int Np, int cropX, int cropY, int bk1x,
                                        int bk1y, int bk2x, int bk2y, int divisor, int bgThreshold, uint16_t* out,
                                        int* bg_val) {
  if (!frame || !out) return fail(-1, "bad argument");
  FPM_Dataset d;
  d.Np = Np; d.cropX = cropX; d.cropY = cropY; d.bk1cropX = bk1x; d.bk1cropY = bk1y; d.bk2cropX = bk2x; d.bk2cropY = bk2y;
  d.bgThreshold = bgThreshold;
  d.darkfieldExpMultiplier = divisor;           // applied when illumination_na > objectiveNA (fpmMain.cpp:128)
  d.objectiveNA = 0.f;
  FPMimg im;
  im.illumination_na = 1.f;
  std::string err;
  if (!preprocessFrame(d, frame, width, height, &im, &err)) return fail(-1, err);
  memcpy(out, im.Image.data(), sizeof(uint16_t) * (size_t)Np * Np);
  if (bg_val) *bg_val = im.bg_val;
  return 0;
}

extern "C" int fpmhost_tile_grid(int width, int height, int Np, int overlap, int* nx, int* ny) {
  if (!nx || !ny || Np <= 0 || overlap < 0 || overlap >= Np || width < Np || height < Np) return fail(-1, "bad argument");
  tileGrid(width, height, Np, overlap, nx, ny);
  return 0;
}

extern "C" int fpmhost_device_from_env(void) { return deviceFromEnv(); }
