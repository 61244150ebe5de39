// Host-side PNG decoder for the patch loader (no device code): what `Image.open(f).convert('RGB')` yields for the patch
// files 1_WSI2Patches.py writes (/root/reference/1_HistoPathology/models.py:280-284, 1_WSI2Patches.py:88-124), decoded
// straight into (pinned) uint8 HWC batches by a pool of threads - the reference decodes in 20 DataLoader workers through
// PIL.  Supports 8-bit, non-interlaced grey / RGB / palette / grey+alpha / RGBA (alpha dropped, like convert('RGB')).
// zlib does the inflate; scanline filters per the PNG specification (RFC 2083 section 6).
#include <zlib.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mmbs {

static inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

static inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// -> 0 on success; `err` receives a short reason otherwise
static int png_decode_rgb8(const uint8_t* data, size_t n, uint8_t* out, int expect_h, int expect_w, const char** err) {
  static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 + 25 || memcmp(data, SIG, 8) != 0) { *err = "not a PNG file"; return 1; }
  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  uint8_t palette[256 * 3];
  memset(palette, 0, sizeof(palette));
  std::vector<uint8_t> idat;
  bool seen_end = false;
  while (pos + 12 <= n && !seen_end) {
    const uint32_t len = be32(data + pos);
    const uint8_t* type = data + pos + 4;
    const uint8_t* body = data + pos + 8;
    if (size_t(len) > n - pos - 12) { *err = "truncated chunk"; return 1; }
    if (memcmp(type, "IHDR", 4) == 0 && len >= 13) {
      w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (memcmp(type, "PLTE", 4) == 0) {
      memcpy(palette, body, len < sizeof(palette) ? len : sizeof(palette));
    } else if (memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), body, body + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      seen_end = true;
    }
    pos += size_t(len) + 12;
  }
  if (ctype < 0 || idat.empty()) { *err = "missing IHDR / IDAT"; return 1; }
  if (depth != 8 || interlace != 0) { *err = "only 8-bit non-interlaced PNG is supported"; return 1; }
  if (int(h) != expect_h || int(w) != expect_w) { *err = "image size differs from the batch's patch size"; return 1; }
  int ch;
  switch (ctype) {
    case 0: ch = 1; break;   // grey
    case 2: ch = 3; break;   // RGB
    case 3: ch = 1; break;   // palette index
    case 4: ch = 2; break;   // grey + alpha
    case 6: ch = 4; break;   // RGBA
    default: *err = "unknown colour type"; return 1;
  }
  const size_t stride = size_t(w) * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf raw_len = uLongf(raw.size());
  if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) {
    *err = "zlib inflate failed";
    return 1;
  }
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* line = raw.data() + (stride + 1) * y;
    const int filter = line[0];
    const uint8_t* src = line + 1;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= size_t(ch) ? cur[i - ch] : 0, b = prev[i], c = i >= size_t(ch) ? prev[i - ch] : 0;
      int v = src[i];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: *err = "bad scanline filter"; return 1;
      }
      cur[i] = uint8_t(v);
    }
    uint8_t* dst = out + size_t(y) * w * 3;
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* px = cur.data() + size_t(x) * ch;
      if (ctype == 2 || ctype == 6) { dst[3 * x] = px[0]; dst[3 * x + 1] = px[1]; dst[3 * x + 2] = px[2]; }
      else if (ctype == 3) { memcpy(dst + 3 * x, palette + 3 * px[0], 3); }
      else { dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = px[0]; }
    }
    prev.swap(cur);
  }
  return 0;
}

static int read_file(const char* path, std::vector<uint8_t>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz <= 0) { fclose(f); return 1; }
  buf.resize(size_t(sz));
  const size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  return got == buf.size() ? 0 : 1;
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_png_decode(const uint8_t* file_bytes, size_t nbytes, uint8_t* out_hwc, int h, int w) {
  MMBS_REQUIRE(file_bytes && out_hwc && h > 0 && w > 0, "mmbs_png_decode: bad argument");
  const char* err = "";
  if (png_decode_rgb8(file_bytes, nbytes, out_hwc, h, w, &err)) {
    set_error("mmbs_png_decode: %s", err);
    return MMBS_ERR_ARG;
  }
  return MMBS_OK;
}

extern "C" int mmbs_png_decode_files(const char* const* paths, int64_t count, uint8_t* out_hwc, int h, int w, int threads) {
  MMBS_REQUIRE(paths && out_hwc && count >= 0 && h > 0 && w > 0, "mmbs_png_decode_files: bad argument");
  if (count == 0) return MMBS_OK;
  int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (int64_t(nt) > count) nt = int(count);
  std::atomic<int64_t> next(0), failed(-1);
  std::vector<const char*> reason(size_t(nt), "");
  auto work = [&](int t) {
    std::vector<uint8_t> buf;
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= count || failed.load() >= 0) return;
      const char* err = "cannot read the file";
      if (read_file(paths[i], buf) || png_decode_rgb8(buf.data(), buf.size(), out_hwc + size_t(i) * h * w * 3, h, w, &err)) {
        int64_t none = -1;
        if (failed.compare_exchange_strong(none, i)) reason[size_t(t)] = err;
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  const int64_t bad = failed.load();
  if (bad >= 0) {
    const char* why = "";
    for (const char* r : reason)
      if (r[0]) why = r;
    set_error("mmbs_png_decode_files: %s: %s", paths[bad], why);
    return MMBS_ERR_ARG;
  }
  return MMBS_OK;
}
