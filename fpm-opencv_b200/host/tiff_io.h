// tiff_io.h -- minimal baseline-TIFF reader/writer for the raw frames the reference loads with
// cv::imread(..., ANYDEPTH) (fpmMain.cpp:110-119).  The old profile shows libtiff's
// DumpModeDecode, i.e. uncompressed strips (output.svg:13,133,645): that is what is supported --
// 8/16-bit, 1..4 samples per pixel (chunky), strips, II or MM byte order.  Anything else is an
// error (no silent fallback).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fpmio {

struct Image16 {
  int width = 0, height = 0, channels = 0;   // samples are converted to uint16 (8-bit kept as 0..255)
  int bits = 0;
  std::vector<uint16_t> pix;                  // [height][width][channels]
};

bool readTiff(const std::string& path, Image16& out, std::string* err);
// One plane of a frame straight into caller memory (e.g. a page-locked staging buffer): [height][width] uint16, sample
// 0 of every pixel (grey frames: the pixel; colour frames: the red plane = channels[2] of OpenCV's BGR, the one the
// reference keeps, fpmMain.cpp:112-115).  Strips are read with pread -- little-endian 16-bit grey frames land in `dst`
// without an intermediate copy; `scratch` is reused between calls for the other layouts.  dst == nullptr: header only.
bool readTiffPlane(const std::string& path, uint16_t* dst, size_t dst_elems, int* width, int* height, int* channels,
                   std::vector<uint8_t>& scratch, std::string* err);
// single-channel 16-bit uncompressed little-endian TIFF, one strip
bool writeTiff16(const std::string& path, const uint16_t* pix, int width, int height, std::string* err);
// single-channel 32-bit float TIFF (SampleFormat = IEEE float), used for amplitude / phase output
bool writeTiffF32(const std::string& path, const float* pix, int width, int height, std::string* err);

}  // namespace fpmio
