// png_io.cpp -- minimal PNG codec on zlib (replaces stb_image_write / stb_image used by the
// reference at path_trace.cu:233 and inv_scene.h:56).  Writer: 8-bit RGB, filter 0, one IDAT.
// Reader: non-interlaced, bit depth 8 or 16, colour types 0/2/3/4/6, converted to 8-bit RGB
// (stbi_load(..., 3) semantics: alpha dropped, grey replicated).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "scene_host.h"

namespace ipt {

namespace {
void put32(std::vector<unsigned char>& b, unsigned v) {
  b.push_back((v >> 24) & 255); b.push_back((v >> 16) & 255); b.push_back((v >> 8) & 255); b.push_back(v & 255);
}
void chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t n) {
  put32(out, (unsigned)n);
  size_t s = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  put32(out, (unsigned)crc32(0L, &out[s], (uInt)(n + 4)));
}
unsigned get32(const unsigned char* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
  int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

bool write_png_rgb8(const char* path, const unsigned char* rgb, int w, int h, std::string* err) {
  std::vector<unsigned char> raw((size_t)h * (3 * (size_t)w + 1));
  for (int y = 0; y < h; y++) {
    raw[(size_t)y * (3 * w + 1)] = 0;
    memcpy(&raw[(size_t)y * (3 * w + 1) + 1], rgb + (size_t)y * 3 * w, 3 * (size_t)w);
  }
  uLongf zn = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zn);
  if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) { *err = "png: deflate failed"; return false; }
  std::vector<unsigned char> out;
  static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  out.insert(out.end(), sig, sig + 8);
  std::vector<unsigned char> ihdr;
  put32(ihdr, (unsigned)w); put32(ihdr, (unsigned)h);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(out, "IHDR", ihdr.data(), ihdr.size());
  chunk(out, "IDAT", z.data(), zn);
  chunk(out, "IEND", nullptr, 0);
  FILE* f = fopen(path, "wb");
  if (!f) { *err = std::string("png: cannot open for writing: ") + path; return false; }
  size_t wr = fwrite(out.data(), 1, out.size(), f);
  fclose(f);
  if (wr != out.size()) { *err = std::string("png: short write: ") + path; return false; }
  return true;
}

static bool read_png_rgb8_impl(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err);
bool read_png_rgb8(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err) {
  try {
    return read_png_rgb8_impl(path, rgb, w, h, err);
  } catch (const std::exception& e) {   // bad_alloc and friends must not cross the C ABI
    *err = std::string("png: ") + e.what();
    return false;
  }
}
static bool read_png_rgb8_impl(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err) {
  FILE* f = fopen(path, "rb");
  if (!f) { *err = std::string("png: cannot open: ") + path; return false; }
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) { *err = std::string("png: bad signature: ") + path; return false; }
  size_t pos = 8;
  int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<unsigned char> idat, palette;
  while (pos + 8 <= buf.size()) {
    unsigned len = get32(&buf[pos]);
    const unsigned char* type = &buf[pos + 4];
    const unsigned char* data = &buf[pos + 8];
    if (pos + 12 + (size_t)len > buf.size()) { *err = "png: truncated chunk"; return false; }
    if (!memcmp(type, "IHDR", 4)) {
      if (len < 13) { *err = "png: short IHDR chunk"; return false; }
      W = (int)get32(data); H = (int)get32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (W <= 0 || H <= 0) { *err = "png: missing IHDR"; return false; }
  // files come from callers (createGraph takes a path): a corrupt header must not turn into a multi-gigabyte allocation
  if (W > 32768 || H > 32768 || (long long)W * H > (1ll << 28)) { *err = "png: image larger than 32768 x 32768 / 2^28 pixels"; return false; }
  if (interlace != 0) { *err = "png: interlaced images are not supported"; return false; }
  if (depth != 8 && depth != 16) { *err = "png: only 8/16-bit samples are supported"; return false; }
  int ch = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
  if (!ch || (ctype == 3 && depth != 8)) { *err = "png: unsupported colour type"; return false; }
  const int bpp = ch * depth / 8;
  const size_t stride = (size_t)W * bpp;
  std::vector<unsigned char> raw((stride + 1) * (size_t)H);
  uLongf rn = (uLongf)raw.size();
  if (uncompress(raw.data(), &rn, idat.data(), (uLong)idat.size()) != Z_OK || rn != raw.size()) { *err = "png: inflate failed"; return false; }
  std::vector<unsigned char> img(stride * (size_t)H);
  for (int y = 0; y < H; y++) {
    const unsigned char* in = &raw[(stride + 1) * (size_t)y];
    unsigned char* cur = &img[stride * (size_t)y];
    const unsigned char* up = y ? &img[stride * (size_t)(y - 1)] : nullptr;
    int ft = in[0];
    in++;
    for (size_t x = 0; x < stride; x++) {
      int a = x >= (size_t)bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
      int v = in[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: *err = "png: bad filter type"; return false;
      }
      cur[x] = (unsigned char)v;
    }
  }
  rgb->resize((size_t)W * H * 3);
  const int step = depth / 8;   // 16-bit: keep the high byte
  for (size_t i = 0; i < (size_t)W * H; i++) {
    const unsigned char* p = &img[i * bpp];
    unsigned char r, g, b;
    if (ctype == 0 || ctype == 4) { r = g = b = p[0]; }
    else if (ctype == 3) {
      size_t k = (size_t)p[0] * 3;
      r = g = b = 0;
      if (k + 2 < palette.size()) { r = palette[k]; g = palette[k + 1]; b = palette[k + 2]; }
    } else { r = p[0]; g = p[step]; b = p[2 * step]; }
    (*rgb)[3 * i] = r; (*rgb)[3 * i + 1] = g; (*rgb)[3 * i + 2] = b;
  }
  *w = W; *h = H;
  return true;
}

}  // namespace ipt
