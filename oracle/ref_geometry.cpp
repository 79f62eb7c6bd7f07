// oracle/ref_geometry.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Golden-vector generator for the bit-exact integer path.  It is compiled against the
// reference's own vendored jsoncpp *where it lies* (/root/reference/include/jsoncpp.cpp,
// see oracle/Makefile; output goes to oracle/_ref/) and performs, with the real
// Json::Value accessors and libstdc++'s std::sort, the statements of
//   main()            fpmMain.cpp:512-575   (config keys, derived optics)
//   loadFPMDataset()  fpmMain.cpp:52-61,77-106,146-168,238-258 (LED geometry + order)
//   runFPM()          fpmMain.cpp:305-306   (naRadius)
// with the same C types as fpmMain.h:19-101.  OpenCV is not available here, so the
// 1x3 * 3x3 cv::Mat_<double> product (fpmMain.cpp:85) is written out as the k=0,1,2
// accumulation cv::gemm performs; tests/golden/make_golden.py cross-checks that order
// against cv2.gemm.  "Which image files exist" is replaced by an explicit LED list:
//   ref_geometry <dataset.json> <Np override or 0> <first led> <last led>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <algorithm>
#include "json.h"

using namespace std;

template <typename T>
std::vector<size_t> sort_indexes(const std::vector<T> &v) {   // fpmMain.h:103-115
  std::vector<size_t> idx(v.size());
  for (size_t i = 0; i != idx.size(); ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&v](size_t i1, size_t i2) { return v[i1] < v[i2]; });
  return idx;
}

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: ref_geometry dataset.json Np first last\n"); return 2; }
  Json::Value datasetJson;
  Json::Reader reader;
  ifstream jsonFile(argv[1]);
  bool parse_ok = reader.parse(jsonFile, datasetJson);          // :515 (result ignored there)

  int16_t Np = datasetJson.get("cropSizeX", 90).asInt();        // :519
  if (atoi(argv[2]) > 0) Np = (int16_t)atoi(argv[2]);
  float pixelSize = datasetJson.get("pixelSize", 6.5).asDouble();
  float objectiveMag = datasetJson.get("objectiveMag", 8).asDouble();
  float objectiveNA = datasetJson.get("objectiveNA", 0.2).asDouble();
  float maxIlluminationNA = datasetJson.get("maxIlluminationNA", 0.7604).asDouble();
  float lambda = datasetJson.get("lambda", 0.5).asDouble();
  float ps_eff = pixelSize / (float)objectiveMag;               // :529
  float du = (1 / ps_eff) / (float)Np;                          // :530
  double arrayRotation = datasetJson.get("arrayRotation", 0).asInt();   // :534
  int16_t resImprovementFactor =
      1 + (int16_t)ceil(2 * ps_eff * (maxIlluminationNA + objectiveNA) / lambda);  // :556-558
  float bgThreshold = datasetJson.get("bgThresh", 1000).asInt();
  int16_t Ncrop = Np, Mcrop = Np;
  int16_t Nlarge = Ncrop * resImprovementFactor;
  int16_t Mlarge = Mcrop * resImprovementFactor;
  float delta1 = datasetJson.get("delta1", 5).asInt();
  float delta2 = datasetJson.get("delta2", 10).asInt();
  uint16_t ledCount = datasetJson.get("ledCount", 508).asInt();
  bool flipIlluminationX = datasetJson.get("flipDatasetX", false).asBool();
  bool flipIlluminationY = datasetJson.get("flipDatasetY", false).asBool();
  uint16_t darkfieldExpMultiplier = datasetJson.get("darkfieldExpMultiplier", 1).asInt();
  Json::Value holeCoordinates = datasetJson.get("holeCoordinates", 0);
  int16_t naRadius = (int16_t)ceil(objectiveNA * ps_eff * Np / lambda);   // :305-306
  size_t hc_size = holeCoordinates.isArray() ? holeCoordinates.size() : 0;

  std::vector<float> illuminationNAList;
  for (int16_t ledIdx = 0; ledIdx <= ledCount; ledIdx++) illuminationNAList.push_back(99.0);  // :52-57

  double angle = arrayRotation;                                  // :60-61
  double R[3][3] = {{cos(angle * M_PI / 180), -sin(angle * M_PI / 180), 0},
                    {sin(angle * M_PI / 180), cos(angle * M_PI / 180), 0},
                    {0, 0, 1}};

  int first = atoi(argv[3]), last = atoi(argv[4]);
  int16_t num_images = 0;
  printf("{\"parse_ok\":%s,\"Np\":%d,\"factor\":%d,\"Nlarge\":%d,\"Mlarge\":%d,\"naRadius\":%d,"
         "\"ps_eff_bits\":%u,\"du_bits\":%u,\"lambda_bits\":%u,\"objectiveNA_bits\":%u,"
         "\"maxIlluminationNA_bits\":%u,\"delta1\":%.9g,\"delta2\":%.9g,\"bgThreshold\":%.9g,"
         "\"ledCount\":%u,\"darkfieldExpMultiplier\":%u,\"arrayRotation\":%.17g,\"holeCoordinatesSize\":%zu,"
         "\"flipX\":%s,\"flipY\":%s,\n\"leds\":[",
         parse_ok ? "true" : "false", Np, resImprovementFactor, Nlarge, Mlarge, naRadius, fbits(ps_eff),
         fbits(du), fbits(lambda), fbits(objectiveNA), fbits(maxIlluminationNA), delta1, delta2,
         bgThreshold, ledCount, darkfieldExpMultiplier, arrayRotation, hc_size,
         flipIlluminationX ? "true" : "false", flipIlluminationY ? "true" : "false");
  bool firstOut = true;
  for (int led_num = first; led_num <= last; led_num++) {
    float posX = holeCoordinates[led_num - 1][0].get("x", 0).asFloat();   // :77-79
    float posY = holeCoordinates[led_num - 1][1].get("y", 0).asFloat();
    float posZ = holeCoordinates[led_num - 1][2].get("z", 0).asFloat();
    double in[3] = {posX, posY, posZ};
    double hc[3];
    for (int j = 0; j < 3; j++) {                                 // :85
      double s = 0;
      for (int k = 0; k < 3; k++) s += in[k] * R[k][j];
      hc[j] = s;
    }
    double flip[3] = {1, 1, 1};                                   // :88-93
    if (flipIlluminationX) { flip[0] = -1; flip[1] = 1; }
    if (flipIlluminationY) { flip[0] = 1; flip[1] = -1; }
    for (int j = 0; j < 3; j++) hc[j] *= flip[j];
    double sinTheta_x = sin(atan2(hc[0], hc[2]));                 // :95-99
    double sinTheta_y = sin(atan2(hc[1], hc[2]));
    float illumination_na = sqrt(sinTheta_x * sinTheta_x + sinTheta_y * sinTheta_y);  // :101-103
    if (sqrt(illumination_na < maxIlluminationNA)) {              // :106
      float uled = sinTheta_x / lambda;                           // :146-147
      float vled = sinTheta_y / lambda;
      int16_t idx_u = (int16_t)round(uled / du);                  // :150-151
      int16_t idx_v = (int16_t)round(vled / du);
      int16_t pupilShiftX = idx_u, pupilShiftY = idx_v;
      int16_t cropXStart = (int16_t)round(Nlarge / 2) + pupilShiftX - (int16_t)round(Ncrop / 2);  // :157-159
      int16_t cropYStart = (int16_t)round(Mlarge / 2) + pupilShiftY - (int16_t)round(Ncrop / 2);  // :163-165
      illuminationNAList.at(led_num) = illumination_na;           // :172-173
      num_images++;
      printf("%s\n{\"n\":%d,\"na_bits\":%u,\"idx_u\":%d,\"idx_v\":%d,\"cropX\":%d,\"cropY\":%d}",
             firstOut ? "" : ",", led_num, fbits(illumination_na), idx_u, idx_v, cropXStart, cropYStart);
      firstOut = false;
    }
  }
  uint16_t ledUsedCount = num_images;                             // :238
  printf("],\n\"ledUsedCount\":%u,\"order\":[", ledUsedCount);
  int16_t indexIncr = 1;                                          // :246-258
  for (auto i : sort_indexes(illuminationNAList)) {
    if (indexIncr <= ledUsedCount) {
      printf("%s%zu", indexIncr == 1 ? "" : ",", i);
      indexIncr++;
    }
  }
  printf("]}\n");
  return 0;
}
