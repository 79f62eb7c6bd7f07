#include "tiff_io.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

namespace fpmio {

namespace {
bool fail(std::string* err, const std::string& m) { if (err) *err = m; return false; }
}  // namespace

namespace {
// Directory of the first image of a classic TIFF, read through `get(offset, dst, n)` (whole-file buffer or pread).
struct Dir {
  uint32_t width = 0, height = 0, bits = 1, comp = 1, spp = 1, rps = 0xFFFFFFFFu, planar = 1, fmt = 1;
  bool be = false;
  std::vector<uint32_t> offs, counts;
};
template <class Get> bool parseDir(Get&& get, size_t file_size, const std::string& path, Dir& d, std::string* err) {
  uint8_t hd[8];
  if (file_size < 8 || !get(0, hd, 8)) return fail(err, path + ": not a TIFF");
  if (hd[0] == 'I' && hd[1] == 'I') d.be = false;
  else if (hd[0] == 'M' && hd[1] == 'M') d.be = true;
  else return fail(err, path + ": not a TIFF");
  auto r16 = [&](const uint8_t* q) { return d.be ? (uint16_t)((q[0] << 8) | q[1]) : (uint16_t)(q[0] | (q[1] << 8)); };
  auto r32 = [&](const uint8_t* q) {
    return d.be ? ((uint32_t)q[0] << 24) | (q[1] << 16) | (q[2] << 8) | q[3] : ((uint32_t)q[3] << 24) | (q[2] << 16) | (q[1] << 8) | q[0];
  };
  if (r16(hd + 2) != 42) return fail(err, path + ": not a classic TIFF (BigTIFF unsupported)");
  const size_t ifd = r32(hd + 4);
  uint8_t nb[2];
  if (ifd + 2 > file_size || !get(ifd, nb, 2)) return fail(err, path + ": malformed TIFF directory");
  const int n = r16(nb);
  std::vector<uint8_t> ent((size_t)n * 12);
  if (ifd + 2 + ent.size() > file_size || (n && !get(ifd + 2, ent.data(), ent.size()))) return fail(err, path + ": malformed TIFF directory");
  bool ok = true;
  auto values = [&](const uint8_t* e, std::vector<uint32_t>& v) {
    const uint16_t type = r16(e + 2);
    const uint32_t cnt = r32(e + 4);
    const size_t sz = type == 3 ? 2 : type == 4 ? 4 : type == 1 ? 1 : 0;
    // a value array can never be longer than the file it lives in (guards the allocation below)
    if (!sz || (uint64_t)sz * cnt > file_size) { ok = false; return; }
    std::vector<uint8_t> raw((size_t)sz * cnt);
    if (sz * cnt <= 4) memcpy(raw.data(), e + 8, raw.size());
    else {
      const size_t base = r32(e + 8);
      if (base + raw.size() > file_size || !get(base, raw.data(), raw.size())) { ok = false; return; }
    }
    v.resize(cnt);
    for (uint32_t k = 0; k < cnt; ++k) v[k] = sz == 2 ? r16(&raw[2 * k]) : sz == 4 ? r32(&raw[4 * k]) : raw[k];
  };
  for (int k = 0; k < n && ok; ++k) {
    const uint8_t* e = &ent[(size_t)k * 12];
    std::vector<uint32_t> v;
    switch (r16(e)) {
      case 256: values(e, v); if (!v.empty()) d.width = v[0]; break;
      case 257: values(e, v); if (!v.empty()) d.height = v[0]; break;
      case 258: values(e, v); if (!v.empty()) d.bits = v[0]; break;
      case 259: values(e, v); if (!v.empty()) d.comp = v[0]; break;
      case 273: values(e, d.offs); break;
      case 277: values(e, v); if (!v.empty()) d.spp = v[0]; break;
      case 278: values(e, v); if (!v.empty()) d.rps = v[0]; break;
      case 279: values(e, d.counts); break;
      case 284: values(e, v); if (!v.empty()) d.planar = v[0]; break;
      case 339: values(e, v); if (!v.empty()) d.fmt = v[0]; break;
      case 322: case 323: case 324: case 325: return fail(err, path + ": tiled TIFF unsupported");
      default: break;
    }
  }
  if (!ok || !d.width || !d.height || d.offs.empty()) return fail(err, path + ": malformed TIFF directory");
  if (d.comp != 1) return fail(err, path + ": compressed TIFF unsupported (only uncompressed strips)");
  if (d.bits != 8 && d.bits != 16) return fail(err, path + ": only 8/16-bit samples supported");
  if (d.fmt != 1) return fail(err, path + ": only unsigned-integer samples supported");
  if (d.spp < 1 || d.spp > 4 || (d.spp > 1 && d.planar != 1)) return fail(err, path + ": unsupported sample layout");
  if (d.rps > d.height) d.rps = d.height;
  if (d.rps == 0) return fail(err, path + ": RowsPerStrip is 0");
  // uncompressed: the pixels are in the file, so a directory that promises more samples than the file has bytes is
  // malformed (and must not drive a multi-gigabyte allocation)
  if ((uint64_t)d.width * d.height * d.spp * (d.bits / 8) > file_size) return fail(err, path + ": image larger than the file");
  if (!d.counts.empty() && d.counts.size() != d.offs.size()) return fail(err, path + ": StripByteCounts does not match StripOffsets");
  const size_t rowb = (size_t)d.width * d.spp * (d.bits / 8);
  size_t row = 0;
  for (size_t s = 0; s < d.offs.size() && row < d.height; ++s) {
    const size_t rows = (d.height - row) < d.rps ? (d.height - row) : d.rps;
    if ((size_t)d.offs[s] + rows * rowb > file_size) return fail(err, path + ": strip exceeds file");
    if (!d.counts.empty() && d.counts[s] < rows * rowb) return fail(err, path + ": StripByteCounts smaller than the strip");
    row += rows;
  }
  if (row != d.height) return fail(err, path + ": missing strips");
  return true;
}
}  // namespace

bool readTiff(const std::string& path, Image16& out, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return fail(err, "cannot open " + path);
  std::vector<uint8_t> buf;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long sz = ftell(f);
    if (sz > 0) buf.resize((size_t)sz);
    rewind(f);
  }
  const bool rd = !buf.empty() && fread(buf.data(), 1, buf.size(), f) == buf.size();
  fclose(f);
  if (!rd) return fail(err, path + ": not a TIFF");
  Dir d;
  auto get = [&](size_t off, void* dst, size_t n) { if (off + n > buf.size()) return false; memcpy(dst, buf.data() + off, n); return true; };
  if (!parseDir(get, buf.size(), path, d, err)) return false;
  const size_t spp = d.spp, width = d.width, height = d.height;
  out.width = (int)width; out.height = (int)height; out.channels = (int)spp; out.bits = (int)d.bits;
  out.pix.assign(width * height * spp, 0);
  size_t row = 0;
  for (size_t s = 0; s < d.offs.size() && row < height; ++s) {
    const size_t rows = (height - row) < d.rps ? (height - row) : d.rps;
    const uint8_t* src = buf.data() + d.offs[s];
    uint16_t* dst = out.pix.data() + row * width * spp;
    const size_t cnt = rows * width * spp;
    if (d.bits == 8) for (size_t k = 0; k < cnt; ++k) dst[k] = src[k];
    else if (d.be) for (size_t k = 0; k < cnt; ++k) dst[k] = (uint16_t)((src[2 * k] << 8) | src[2 * k + 1]);
    else memcpy(dst, src, cnt * 2);
    row += rows;
  }
  return true;
}

bool readTiffPlane(const std::string& path, uint16_t* dst, size_t dst_elems, int* width, int* height, int* channels,
                   std::vector<uint8_t>& scratch, std::string* err) {
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return fail(err, "cannot open " + path);
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 8) { close(fd); return fail(err, path + ": not a TIFF"); }
  const size_t file_size = (size_t)st.st_size;
  auto get = [&](size_t off, void* p, size_t n) {
    size_t done = 0;
    while (done < n) {
      const ssize_t r = pread(fd, (char*)p + done, n - done, (off_t)(off + done));
      if (r <= 0) return false;
      done += (size_t)r;
    }
    return true;
  };
  Dir d;
  if (!parseDir(get, file_size, path, d, err)) { close(fd); return false; }
  *width = (int)d.width; *height = (int)d.height; *channels = (int)d.spp;
  if (dst == nullptr) { close(fd); return true; }                    // header only
  if ((size_t)d.width * d.height > dst_elems) { close(fd); return fail(err, path + ": frame larger than the staging buffer"); }
  const size_t spp = d.spp, bps = d.bits / 8, width_ = d.width;
  size_t row = 0;
  bool ok = true;
  for (size_t s = 0; s < d.offs.size() && row < d.height && ok; ++s) {
    const size_t rows = (d.height - row) < d.rps ? (d.height - row) : d.rps;
    uint16_t* out = dst + row * width_;
    if (spp == 1 && bps == 2 && !d.be) {
      ok = get(d.offs[s], out, rows * width_ * 2);                    // little-endian grey 16-bit: straight into the buffer
    } else {
      scratch.resize(rows * width_ * spp * bps);
      ok = get(d.offs[s], scratch.data(), scratch.size());
      const uint8_t* src = scratch.data();
      const size_t cnt = rows * width_;
      // sample 0 of every pixel: the plane the reference keeps of a colour frame (fpmMain.cpp:112-115)
      if (bps == 1) for (size_t k = 0; k < cnt; ++k) out[k] = src[k * spp];
      else if (d.be) for (size_t k = 0; k < cnt; ++k) out[k] = (uint16_t)((src[2 * k * spp] << 8) | src[2 * k * spp + 1]);
      else for (size_t k = 0; k < cnt; ++k) out[k] = (uint16_t)(src[2 * k * spp] | (src[2 * k * spp + 1] << 8));
    }
    row += rows;
  }
  close(fd);
  if (!ok) return fail(err, path + ": short read");
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
