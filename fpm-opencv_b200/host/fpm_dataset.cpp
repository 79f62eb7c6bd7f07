// fpm_dataset.cpp -- configuration, LED geometry, LED order and the image loader.
// Restates main() (fpmMain.cpp:512-584) and loadFPMDataset() (fpmMain.cpp:36-271) without OpenCV
// or jsoncpp.  The integer results (idx_u/idx_v, crop boxes, sortedIndicies) must be bit-exact,
// so every expression keeps the reference's operand types and this file
// is compiled with -ffp-contract=off.
#include "fpm_dataset.h"

#include <dirent.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "tiff_io.h"

namespace {
const float kDomeHoleTable[508 * 3] = {
#include "dome_table.inc"
};

using mjson::Value;

std::string dirOf(const std::string& p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string("") : p.substr(0, k + 1);
}

bool loadLedList(const std::string& file, const std::string& jsonDir, std::vector<std::array<float, 3>>* out,
                 std::string* used) {
  // ledArrayMaps format (ledArrayMaps/fLED-c.json:9-591): {"ledList":[{"ledNum":n,"x":..,"y":..,"z":..},..]}
  for (const std::string& cand : {file, jsonDir + file}) {
    Value root;
    mjson::parseFile(cand, root);
    const Value none;
    const Value& list = root.isObject() ? root.get("ledList", none) : none;
    if (!list.isArray() || list.size() == 0) continue;
    size_t maxn = 0;
    for (size_t k = 0; k < list.size(); ++k) {
      const Value& e = list.at((long long)k);
      if (!e.isObject()) continue;
      int n = e.get("ledNum", Value((int)(k + 1))).asInt();
      if (n >= 1) maxn = std::max(maxn, (size_t)n);
    }
    out->assign(maxn, {0.f, 0.f, 0.f});
    for (size_t k = 0; k < list.size(); ++k) {
      const Value& e = list.at((long long)k);
      if (!e.isObject()) continue;
      int n = e.get("ledNum", Value((int)(k + 1))).asInt();
      if (n < 1) continue;
      (*out)[n - 1] = {e.get("x", Value(0)).asFloat(), e.get("y", Value(0)).asFloat(), e.get("z", Value(0)).asFloat()};
    }
    *used = cand;
    return true;
  }
  return false;
}
}  // namespace

bool readDatasetJson(const std::string& path, int itrCount, FPM_Dataset* d) {
  Value datasetJson;
  std::string perr;
  bool opened = true;
  {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) opened = false; else fclose(f);
  }
  mjson::parseFile(path, datasetJson, &perr);           // result ignored, fpmMain.cpp:515
  d->jsonDir = dirOf(path);

  d->filePrefix = datasetJson.get("filePrefix", Value("iLED_")).asString();                  // :517
  d->fileExtension = datasetJson.get("fileExtension", Value(".tif")).asString();
  d->Np = datasetJson.get("cropSizeX", Value(90)).asInt();
  d->datasetRoot = datasetJson.get("datasetRoot", Value(".")).asString();
  d->pixelSize = datasetJson.get("pixelSize", Value(6.5)).asDouble();
  d->objectiveMag = datasetJson.get("objectiveMag", Value(8)).asDouble();
  d->objectiveNA = datasetJson.get("objectiveNA", Value(0.2)).asDouble();
  d->maxIlluminationNA = datasetJson.get("maxIlluminationNA", Value(0.7604)).asDouble();
  d->color = datasetJson.get("isColor", Value(false)).asBool();
  d->centerLED = datasetJson.get("centerLED", Value(249)).asInt();
  d->lambda = datasetJson.get("lambda", Value(0.5)).asDouble();
  d->ps_eff = d->pixelSize / (float)d->objectiveMag;                                         // :529
  d->du = (1 / d->ps_eff) / (float)d->Np;                                                    // :530
  d->leadingZeros = datasetJson.get("leadingZeros", Value(false)).asBool();
  d->cropX = datasetJson.get("cropX", Value(1)).asInt();
  d->cropY = datasetJson.get("cropY", Value(1)).asInt();
  d->arrayRotation = datasetJson.get("arrayRotation", Value(0)).asInt();                     // :534
  d->bk1cropX = datasetJson.get("bk1cropX", Value(1)).asInt();
  d->bk1cropY = datasetJson.get("bk1cropY", Value(1)).asInt();
  d->bk2cropX = datasetJson.get("bk2cropX", Value(1)).asInt();
  d->bk2cropY = datasetJson.get("bk2cropY", Value(1)).asInt();
  d->holeNumberDigits = datasetJson.get("holeNumberDigits", Value(4)).asInt();
  d->resImprovementFactor =
      1 + (int16_t)std::ceil(2 * d->ps_eff * (d->maxIlluminationNA + d->objectiveNA) / d->lambda);   // :556-558
  d->bgThreshold = datasetJson.get("bgThresh", Value(1000)).asInt();                         // :561
  d->Mcrop = d->Np;
  d->Ncrop = d->Np;
  d->Nlarge = d->Ncrop * d->resImprovementFactor;
  d->Mlarge = d->Mcrop * d->resImprovementFactor;
  d->ps = d->ps_eff / (float)d->resImprovementFactor;
  d->delta1 = datasetJson.get("delta1", Value(5)).asInt();                                   // :567
  d->delta2 = datasetJson.get("delta2", Value(10)).asInt();
  d->itrCount = (int16_t)itrCount;                                                           // :569
  d->ledCount = datasetJson.get("ledCount", Value(508)).asInt();
  d->flipIlluminationX = datasetJson.get("flipDatasetX", Value(false)).asBool();
  d->flipIlluminationY = datasetJson.get("flipDatasetY", Value(false)).asBool();
  d->darkfieldExpMultiplier = datasetJson.get("darkfieldExpMultiplier", Value(1)).asInt();
  d->holeCoordinateFileName = datasetJson.get("holeCoordinateFileName", Value("null")).asString();
  d->holeCoordinates = datasetJson.get("holeCoordinates", Value(0));                         // :575
  d->debug = datasetJson.get("debug", Value(false)).asBool();                                // :584
  d->naRadius = (int16_t)std::ceil(d->objectiveNA * d->ps_eff * d->Np / d->lambda);          // :305-306

  // ---- geometry source.  HEAD of the reference only reads `holeCoordinates`
  // and throws otherwise; the other LED formats the repository ships are accepted here. ----
  d->ledXYZ.clear();
  if (d->holeCoordinates.isArray()) {
    d->geometrySource = "holeCoordinates";
  } else {
    std::string fn = d->holeCoordinateFileName;
    if (fn == "null") fn = datasetJson.get("holeCoordinatFile", Value("null")).asString();  // key of dataset_fLED-c.json:27
    std::string used;
    const Value none;
    const Value& hp = datasetJson.isObject() ? datasetJson.get("holePositions", none) : none;
    if (fn != "null" && loadLedList(fn, d->jsonDir, &d->ledXYZ, &used)) {
      d->geometrySource = "ledList:" + used;
    } else if (hp.isArray() && hp.size() > 0) {
      // dataset_cellscope2.json:25 -- same dome as include/domeHoleCoordinates.h with the columns
      // permuted: header (c0,c1,c2) == json (z,y,x); the optical axis is json-x.
      if (fn != "null") std::cerr << "warning: LED map '" << fn << "' not found; using holePositions" << std::endl;
      d->ledXYZ.resize(hp.size());
      for (size_t k = 0; k < hp.size(); ++k) {
        const Value& row = hp.at((long long)k);
        d->ledXYZ[k] = {row.at(2).get("z", Value(0)).asFloat(), row.at(1).get("y", Value(0)).asFloat(),
                        row.at(0).get("x", Value(0)).asFloat()};
      }
      d->geometrySource = "holePositions";
    } else {
      if (fn != "null") std::cerr << "warning: LED map '" << fn << "' not found; using the built-in dome table" << std::endl;
      d->ledXYZ.resize(508);
      for (int k = 0; k < 508; ++k)
        d->ledXYZ[k] = {kDomeHoleTable[3 * k], kDomeHoleTable[3 * k + 1], kDomeHoleTable[3 * k + 2]};
      d->geometrySource = "domeHoleCoordinates";
    }
  }
  d->cudaDevice = deviceFromEnv();
  return opened;
}

bool computeLedGeometry(const FPM_Dataset& ds, int led_num, FPMimg* im) {
  float posX, posY, posZ;
  if (ds.holeCoordinates.isArray()) {                                                         // :77-79
    const Value& row = ds.holeCoordinates.at((long long)led_num - 1);
    posX = row.at(0).get("x", Value(0)).asFloat();
    posY = row.at(1).get("y", Value(0)).asFloat();
    posZ = row.at(2).get("z", Value(0)).asFloat();
  } else if (!ds.ledXYZ.empty()) {
    if (led_num >= 1 && (size_t)led_num <= ds.ledXYZ.size()) {
      posX = ds.ledXYZ[led_num - 1][0]; posY = ds.ledXYZ[led_num - 1][1]; posZ = ds.ledXYZ[led_num - 1][2];
    } else posX = posY = posZ = 0.f;
  } else {
    throw std::runtime_error("no LED geometry: holeCoordinates is not an array");
  }
  const double angle = ds.arrayRotation;                                                     // :60-61
  const double R[3][3] = {{std::cos(angle * M_PI / 180), -std::sin(angle * M_PI / 180), 0},
                          {std::sin(angle * M_PI / 180), std::cos(angle * M_PI / 180), 0},
                          {0, 0, 1}};
  const double in[3] = {posX, posY, posZ};
  double hc[3];
  for (int j = 0; j < 3; ++j) {                  // 1x3 * 3x3 double product, k = 0,1,2 accumulation (:85)
    double s = 0;
    for (int k = 0; k < 3; ++k) s += in[k] * R[k][j];
    hc[j] = s;
  }
  double flip[3] = {1, 1, 1};                                                                // :88-93
  if (ds.flipIlluminationX) { flip[0] = -1; flip[1] = 1; }
  if (ds.flipIlluminationY) { flip[0] = 1; flip[1] = -1; }
  for (int j = 0; j < 3; ++j) hc[j] *= flip[j];

  im->led_num = led_num;
  im->sinTheta_x = std::sin(std::atan2(hc[0], hc[2]));                                       // :95-99
  im->sinTheta_y = std::sin(std::atan2(hc[1], hc[2]));
  im->illumination_na = std::sqrt(im->sinTheta_x * im->sinTheta_x + im->sinTheta_y * im->sinTheta_y);  // :101-103
  if (!(im->illumination_na < ds.maxIlluminationNA)) return false;                            // :106
  im->uled = im->sinTheta_x / ds.lambda;                                                     // :146-147
  im->vled = im->sinTheta_y / ds.lambda;
  im->idx_u = (int16_t)std::round(im->uled / ds.du);                                         // :150-151
  im->idx_v = (int16_t)std::round(im->vled / ds.du);
  im->pupilShiftX = im->idx_u;
  im->pupilShiftY = im->idx_v;
  im->cropXStart = (int16_t)std::round(ds.Nlarge / 2) + im->pupilShiftX - (int16_t)std::round(ds.Ncrop / 2);       // :157-159
  im->cropXEnd = (int16_t)std::round(ds.Nlarge / 2) + im->pupilShiftX + (int16_t)std::round(ds.Ncrop / 2) - 1;
  im->cropYStart = (int16_t)std::round(ds.Mlarge / 2) + im->pupilShiftY - (int16_t)std::round(ds.Ncrop / 2);       // :163-165
  im->cropYEnd = (int16_t)std::round(ds.Mlarge / 2) + im->pupilShiftY + (int16_t)std::round(ds.Ncrop / 2) - 1;
  return true;
}

void allocateImageStack(FPM_Dataset* d) {                                                    // :52-57
  d->imageStack.assign((size_t)d->ledCount + 1, FPMimg());
  d->illuminationNAList.assign((size_t)d->ledCount + 1, 99.0f);
  d->NALedPatternStackX.assign((size_t)d->ledCount + 1, -1.0f);
  d->NALedPatternStackY.assign((size_t)d->ledCount + 1, -1.0f);
  d->sortedIndicies.clear();
  d->sortedNALedPatternStackX.clear();
  d->sortedNALedPatternStackY.clear();
  d->ledUsedCount = 0;
}

void registerImage(FPM_Dataset* d, const FPMimg& im) {                                       // :171-177
  d->imageStack.at(im.led_num) = im;               // throws std::out_of_range like the reference
  d->illuminationNAList.at(im.led_num) = im.illumination_na;
  d->NALedPatternStackX.at(im.led_num) = im.sinTheta_x;
  d->NALedPatternStackY.at(im.led_num) = im.sinTheta_y;
}

void sortLedOrder(FPM_Dataset* d) {                                                          // fpmMain.h:103-115, fpmMain.cpp:246-258
  const std::vector<float>& v = d->illuminationNAList;
  std::vector<size_t> idx(v.size());
  for (size_t i = 0; i != idx.size(); ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&v](size_t i1, size_t i2) { return v[i1] < v[i2]; });
  int16_t indexIncr = 1;
  d->sortedIndicies.clear();
  for (auto i : idx) {
    if (indexIncr <= d->ledUsedCount) {
      d->sortedIndicies.push_back((int16_t)i);
      d->sortedNALedPatternStackX.push_back(d->NALedPatternStackX[i]);
      d->sortedNALedPatternStackY.push_back(d->NALedPatternStackY[i]);
      indexIncr++;
    }
  }
}

int16_t backgroundValue(const FPM_Dataset& d, const uint16_t* frame, int w) {
  const int Np = d.Np;
  auto roiMean = [&](int x0, int y0) {                                                        // cv::mean: sum * (1/N)
    unsigned long long s = 0;
    for (int y = 0; y < Np; ++y)
      for (int x = 0; x < Np; ++x) s += frame[(size_t)(y0 + y) * w + x0 + x];
    return (double)s * (1.0 / ((double)Np * Np));
  };
  const double bk1 = roiMean(d.bk1cropX, d.bk1cropY), bk2 = roiMean(d.bk2cropX, d.bk2cropY);  // :131-134
  double bg_val = (bk2 + bk1) / 2;                                                            // :136-138
  if (bg_val > d.bgThreshold) bg_val = d.bgThreshold;
  return (int16_t)(int)std::round(bg_val);                                                    // :140
}

bool preprocessFrame(const FPM_Dataset& d, const uint16_t* frame, int w, int h, FPMimg* im, std::string* err) {
  const int Np = d.Np;
  auto inside = [&](int x, int y) { return x >= 0 && y >= 0 && x + Np <= w && y + Np <= h; };
  if (!inside(d.cropX, d.cropY) || !inside(d.bk1cropX, d.bk1cropY) || !inside(d.bk2cropX, d.bk2cropY)) {
    if (err) *err = "crop / background ROI leaves the " + std::to_string(w) + "x" + std::to_string(h) + " frame";
    return false;
  }
  im->Image.resize((size_t)Np * Np);                                                          // :124-125
  for (int y = 0; y < Np; ++y)
    memcpy(&im->Image[(size_t)y * Np], frame + (size_t)(d.cropY + y) * w + d.cropX, sizeof(uint16_t) * Np);
  if (d.darkfieldExpMultiplier != 1 && im->illumination_na > d.objectiveNA) {                 // :128-129 cv::divide
    const double m = (double)d.darkfieldExpMultiplier;
    for (auto& p : im->Image) {
      double q = m == 0 ? 0.0 : std::nearbyint((double)p / m);      // cvRound: half to even; x/0 -> 0
      p = (uint16_t)(q < 0 ? 0 : q > 65535 ? 65535 : q);
    }
  }
  im->bg_val = backgroundValue(d, frame, w);                                                  // :131-140
  for (auto& p : im->Image) {                                                                 // :143-144 saturating
    int v = (int)p - (int)im->bg_val;
    p = (uint16_t)(v < 0 ? 0 : v > 65535 ? 65535 : v);
  }
  return true;
}

int16_t loadFPMDataset(FPM_Dataset* d) {
  allocateImageStack(d);
  DIR* dir = opendir(d->datasetRoot.c_str());
  if (dir == NULL) {
    std::cout << "ERROR: Could not Open Directory.\n";                                       // :268
    return -1;
  }
  int16_t num_images = 0;
  std::cout << "Loading Images..." << std::endl;                                             // :65
  struct dirent* ent;
  while ((ent = readdir(dir)) != NULL) {
    std::string fileName = ent->d_name;
    const size_t el = d->fileExtension.length(), pl = d->filePrefix.length();
    if (fileName == "." || fileName == ".." || fileName.length() < el + pl) continue;
    if (fileName.compare(fileName.length() - el, el, d->fileExtension) != 0 || fileName.find(d->filePrefix) != 0) continue;  // :69-70
    std::string holeNum = fileName.substr(pl, fileName.length() - el - pl);                  // :71-73
    FPMimg currentImage;
    const int led_num = atoi(holeNum.c_str());                                                // :75
    bool pass;
    try {
      pass = computeLedGeometry(*d, led_num, &currentImage);
    } catch (const std::exception& e) {
      std::cout << "ERROR: " << e.what() << std::endl;
      closedir(dir);
      return -1;
    }
    std::cout << "NA:" << std::sqrt(currentImage.sinTheta_x * currentImage.sinTheta_x +
                                    currentImage.sinTheta_y * currentImage.sinTheta_y) << std::endl;   // :105
    if (!pass) {
      std::cout << "Skipped LED# " << holeNum << std::endl;                                  // :236
      continue;
    }
    if (led_num < 0 || led_num > d->ledCount) {
      std::cout << "ERROR: LED # " << led_num << " exceeds ledCount " << d->ledCount
                << " (the reference throws std::out_of_range here, fpmMain.cpp:171)" << std::endl;
      closedir(dir);
      return -1;
    }
    fpmio::Image16 full;
    std::string err;
    if (!fpmio::readTiff(d->datasetRoot + fileName, full, &err)) {                            // :110,119 (no separator added)
      std::cout << "ERROR: " << err << std::endl;
      closedir(dir);
      return -1;
    }
    std::vector<uint16_t> plane;
    const uint16_t* frame = full.pix.data();
    if (full.channels != 1) {
      if (!d->color) {
        std::cout << "ERROR: " << fileName << " has " << full.channels << " channels but isColor is false" << std::endl;
        closedir(dir);
        return -1;
      }
      // imread gives BGR, the reference keeps channels[2] (fpmMain.cpp:112-115) = TIFF sample 0
      plane.resize((size_t)full.width * full.height);
      for (size_t k = 0; k < plane.size(); ++k) plane[k] = full.pix[k * full.channels];
      frame = plane.data();
    }
    if (!preprocessFrame(*d, frame, full.width, full.height, &currentImage, &err)) {
      std::cout << "ERROR: " << fileName << ": " << err << std::endl;
      closedir(dir);
      return -1;
    }
    registerImage(d, currentImage);
    num_images++;
    std::cout << "Loaded: " << fileName << ", LED # is: " << currentImage.led_num << std::endl;   // :180-181
  }
  d->ledUsedCount = num_images;                                                               // :238
  closedir(dir);
  if (num_images <= 0) {
    std::cout << "ERROR - No images found in given directory." << std::endl;                  // :242
    return -1;
  }
  sortLedOrder(d);
  return 1;
}

void tileGrid(int width, int height, int Np, int overlap, int* nx, int* ny) {
  const int step = Np - overlap;
  *nx = (width - Np) / step + 1;
  *ny = (height - Np) / step + 1;
}

void makePupilSupport(int Np, int radius, std::vector<float>* mask) {
  // cv::circle(filled) == {dx^2 + dy^2 <= r^2} (checked against cv2.circle for r = 1..199 in
  // tests/test_oracle.py), then fftShift = circular shift by Np/2 (fpmMain.cpp:304-310)
  mask->assign((size_t)Np * Np, 0.f);
  const int c = Np / 2;
  for (int y = 0; y < Np; ++y)
    for (int x = 0; x < Np; ++x)
      if ((x - c) * (x - c) + (y - c) * (y - c) <= radius * radius)
        (*mask)[(size_t)((y + c) % Np) * Np + (x + c) % Np] = 1.f;
}

int deviceFromEnv() {
  const char* e = getenv("OPENCV_OPENCL_DEVICE");        // use_gpu.sh:1 / use_cpu.sh:1
  if (!e || !*e) return 0;
  std::string s(e);
  if (s.compare(0, 3, "CPU") == 0) return -1;
  size_t k = s.find_last_of(':');
  if (k == std::string::npos || k + 1 >= s.size()) return 0;
  int n = atoi(s.c_str() + k + 1);
  return n < 0 ? 0 : n;
}
