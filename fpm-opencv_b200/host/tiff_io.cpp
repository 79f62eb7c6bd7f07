#include "tiff_io.h"

#include <cstdio>
#include <cstring>
#include <fstream>

namespace fpmio {

namespace {
struct Rd {
  const std::vector<uint8_t>& b;
  bool be;
  bool ok = true;
  uint8_t u8(size_t o) {
    if (o + 1 > b.size()) { ok = false; return 0; }
    return b[o];
  }
  uint16_t u16(size_t o) {
    if (o + 2 > b.size()) { ok = false; return 0; }
    return be ? (uint16_t)((b[o] << 8) | b[o + 1]) : (uint16_t)(b[o] | (b[o + 1] << 8));
  }
  uint32_t u32(size_t o) {
    if (o + 4 > b.size()) { ok = false; return 0; }
    return be ? ((uint32_t)b[o] << 24) | (b[o + 1] << 16) | (b[o + 2] << 8) | b[o + 3]
              : ((uint32_t)b[o + 3] << 24) | (b[o + 2] << 16) | (b[o + 1] << 8) | b[o];
  }
};
bool fail(std::string* err, const std::string& m) { if (err) *err = m; return false; }
}  // namespace

bool readTiff(const std::string& path, Image16& out, std::string* err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(err, "cannot open " + path);
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 8) return fail(err, path + ": not a TIFF");
  bool be;
  if (buf[0] == 'I' && buf[1] == 'I') be = false;
  else if (buf[0] == 'M' && buf[1] == 'M') be = true;
  else return fail(err, path + ": not a TIFF");
  Rd r{buf, be};
  if (r.u16(2) != 42) return fail(err, path + ": not a classic TIFF (BigTIFF unsupported)");
  size_t ifd = r.u32(4);
  int n = r.u16(ifd);
  uint32_t width = 0, height = 0, bits = 1, comp = 1, spp = 1, rps = 0xFFFFFFFFu, planar = 1, fmt = 1;
  std::vector<uint32_t> offs, counts;
  auto values = [&](size_t e, std::vector<uint32_t>& v) {
    uint16_t type = r.u16(e + 2);
    uint32_t cnt = r.u32(e + 4);
    size_t sz = type == 3 ? 2 : type == 4 ? 4 : type == 1 ? 1 : 0;
    if (!sz) { r.ok = false; return; }
    // a value array can never be longer than the file it lives in (guards the allocation below)
    if ((uint64_t)sz * cnt > buf.size()) { r.ok = false; return; }
    size_t base = (sz * cnt <= 4) ? e + 8 : r.u32(e + 8);
    v.resize(cnt);
    for (uint32_t k = 0; k < cnt && r.ok; ++k) v[k] = sz == 2 ? r.u16(base + 2 * k) : sz == 4 ? r.u32(base + 4 * k) : r.u8(base + k);
  };
  for (int k = 0; k < n && r.ok; ++k) {
    size_t e = ifd + 2 + 12 * (size_t)k;
    uint16_t tag = r.u16(e);
    std::vector<uint32_t> v;
    switch (tag) {
      case 256: values(e, v); if (!v.empty()) width = v[0]; break;
      case 257: values(e, v); if (!v.empty()) height = v[0]; break;
      case 258: values(e, v); if (!v.empty()) bits = v[0]; break;
      case 259: values(e, v); if (!v.empty()) comp = v[0]; break;
      case 273: values(e, offs); break;
      case 277: values(e, v); if (!v.empty()) spp = v[0]; break;
      case 278: values(e, v); if (!v.empty()) rps = v[0]; break;
      case 279: values(e, counts); break;
      case 284: values(e, v); if (!v.empty()) planar = v[0]; break;
      case 339: values(e, v); if (!v.empty()) fmt = v[0]; break;
      case 322: case 323: case 324: case 325: return fail(err, path + ": tiled TIFF unsupported");
      default: break;
    }
  }
  if (!r.ok || !width || !height || offs.empty()) return fail(err, path + ": malformed TIFF directory");
  if (comp != 1) return fail(err, path + ": compressed TIFF unsupported (only uncompressed strips)");
  if (bits != 8 && bits != 16) return fail(err, path + ": only 8/16-bit samples supported");
  if (fmt != 1) return fail(err, path + ": only unsigned-integer samples supported");
  if (spp < 1 || spp > 4 || (spp > 1 && planar != 1)) return fail(err, path + ": unsupported sample layout");
  if (rps > height) rps = height;
  if (rps == 0) return fail(err, path + ": RowsPerStrip is 0");
  const size_t bps = bits / 8, rowb = (size_t)width * spp * bps;
  // uncompressed: the pixels are in the file, so a directory that promises more samples than the file has bytes is
  // malformed (and must not drive a multi-gigabyte allocation)
  if ((uint64_t)width * height * spp * bps > buf.size()) return fail(err, path + ": image larger than the file");
  if (!counts.empty() && counts.size() != offs.size()) return fail(err, path + ": StripByteCounts does not match StripOffsets");
  out.width = (int)width; out.height = (int)height; out.channels = (int)spp; out.bits = (int)bits;
  out.pix.assign((size_t)width * height * spp, 0);
  size_t row = 0;
  for (size_t s = 0; s < offs.size() && row < height; ++s) {
    size_t rows = (height - row) < rps ? (height - row) : rps;
    size_t need = rows * rowb;
    if ((size_t)offs[s] + need > buf.size()) return fail(err, path + ": strip exceeds file");
    if (!counts.empty() && counts[s] < need) return fail(err, path + ": StripByteCounts smaller than the strip");
    const uint8_t* src = buf.data() + offs[s];
    uint16_t* dst = out.pix.data() + row * width * spp;
    const size_t cnt = rows * width * spp;
    if (bits == 8) for (size_t k = 0; k < cnt; ++k) dst[k] = src[k];
    else if (be) for (size_t k = 0; k < cnt; ++k) dst[k] = (uint16_t)((src[2 * k] << 8) | src[2 * k + 1]);
    else memcpy(dst, src, cnt * 2);
    row += rows;
  }
  if (row != height) return fail(err, path + ": missing strips");
  return true;
}

namespace {
void put16(std::vector<uint8_t>& b, uint16_t v) { b.push_back(v & 255); b.push_back(v >> 8); }
void put32(std::vector<uint8_t>& b, uint32_t v) { for (int k = 0; k < 4; ++k) b.push_back((v >> (8 * k)) & 255); }
void entry(std::vector<uint8_t>& b, uint16_t tag, uint16_t type, uint32_t cnt, uint32_t val) {
  put16(b, tag); put16(b, type); put32(b, cnt);
  if (type == 3 && cnt == 1) { put16(b, (uint16_t)val); put16(b, 0); } else put32(b, val);
}
bool writeTiff(const std::string& path, const void* pix, int w, int h, int bits, int fmt, std::string* err) {
  const uint32_t nbytes = (uint32_t)w * h * (bits / 8);
  std::vector<uint8_t> hd;
  hd.push_back('I'); hd.push_back('I'); put16(hd, 42); put32(hd, 8 + nbytes);   // IFD after the pixels
  std::vector<uint8_t> ifd;
  put16(ifd, 10);
  entry(ifd, 256, 4, 1, w); entry(ifd, 257, 4, 1, h); entry(ifd, 258, 3, 1, bits); entry(ifd, 259, 3, 1, 1);
  entry(ifd, 262, 3, 1, 1); entry(ifd, 273, 4, 1, 8); entry(ifd, 277, 3, 1, 1); entry(ifd, 278, 4, 1, h);
  entry(ifd, 279, 4, 1, nbytes); entry(ifd, 339, 3, 1, fmt);
  put32(ifd, 0);
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return fail(err, "cannot write " + path);
  bool ok = fwrite(hd.data(), 1, hd.size(), f) == hd.size() && fwrite(pix, 1, nbytes, f) == nbytes &&
            fwrite(ifd.data(), 1, ifd.size(), f) == ifd.size();
  fclose(f);
  return ok ? true : fail(err, "short write " + path);
}
}  // namespace

bool writeTiff16(const std::string& path, const uint16_t* pix, int w, int h, std::string* err) {
  return writeTiff(path, pix, w, h, 16, 1, err);
}
bool writeTiffF32(const std::string& path, const float* pix, int w, int h, std::string* err) {
  return writeTiff(path, pix, w, h, 32, 3, err);
}

}  // namespace fpmio
