// Multi-threaded writer for the feature matrices of the extraction scripts (host code).
//
// Replaces  np.savetxt(path, features, delimiter=",")  of
//   /root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:184-192 and
//   /root/reference/2_GeneExpression/3_GeneExpress_extractfeatures.py:143-149
// byte for byte: numpy's default format '%.18e', ',' between columns, '\n' after every row.  np.savetxt formats one
// value at a time in the interpreter (n_cases x 2048 values: ~1.2 us per value); here every thread formats a band of
// rows into its own buffer with snprintf (same correctly-rounded decimal digits as Python's % operator) and the bands
// are written out in order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int mmbs_write_matrix_csv(const double* data, int64_t rows, int64_t cols, const char* path, int32_t threads) {
  using namespace mmbs;
  MMBS_REQUIRE(data != nullptr || rows * cols == 0, "mmbs_write_matrix_csv: null data");
  MMBS_REQUIRE(path != nullptr && rows >= 0 && cols >= 1, "mmbs_write_matrix_csv: bad arguments (rows=%lld cols=%lld)",
               (long long)rows, (long long)cols);
  FILE* f = fopen(path, "wb");
  if (f == nullptr) {
    set_error("mmbs_write_matrix_csv: cannot open %s", path);
    return MMBS_ERR_ARG;
  }
  int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
  nt = std::max(1, std::min<int>(nt, 64));
  const int64_t band = 256;   // rows per work item
  const int64_t n_bands = (rows + band - 1) / band;
  std::vector<std::string> out(static_cast<size_t>(n_bands));
  auto format_bands = [&](int t) {
    char tmp[64];
    for (int64_t b = t; b < n_bands; b += nt) {
      std::string& s = out[size_t(b)];
      const int64_t r0 = b * band, r1 = std::min(rows, r0 + band);
      s.reserve(size_t((r1 - r0) * cols * 26));
      for (int64_t r = r0; r < r1; ++r) {
        const double* row = data + r * cols;
        for (int64_t c = 0; c < cols; ++c) {
          const int len = snprintf(tmp, sizeof(tmp), "%.18e", row[c]);
          s.append(tmp, size_t(len));
          s.push_back(c + 1 == cols ? '\n' : ',');
        }
      }
    }
  };
  // bands are formatted in waves of nt * 8 so that memory stays bounded for very large matrices
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(format_bands, t);
  format_bands(0);
  for (auto& th : pool) th.join();
  bool ok = true;
  for (const std::string& s : out) ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    set_error("mmbs_write_matrix_csv: short write to %s", path);
    return MMBS_ERR_ARG;
  }
  return MMBS_OK;
}
