// Two-pass risk-set sort (MSD partition by an equalised-CDF bucket map + block-local sort with the forward gather
// fused): see cox_sort.cu.  Used by the Cox loss for 2048 < n <= FS_MAX_N.
#pragma once

#include "radix_sort.cuh"

namespace mmbs {

constexpr int FS_BINS = 4096;             // histogram of the top 12 key bits
constexpr int FS_MAX_BUCKETS = 1024;
constexpr int FS_CAP = 16384;             // samples one block sorts in shared memory
constexpr int FS_TARGET = 8192;           // aimed bucket size (n <= 8.4 M); 10 M samples -> 9766 per bucket
constexpr int FS_PART_TILE = 8192;        // keys per partition tile
constexpr int FS_MAX_WORK = 2048;         // >= FS_MAX_BUCKETS + FS_MAX_N / FS_CAP + 1
constexpr int64_t FS_MAX_N = 12000000;    // beyond: mean bucket size too close to FS_CAP -> LSD sort

static inline int fs_num_buckets(int64_t n) {
  const int64_t nb = (n + FS_TARGET - 1) / FS_TARGET;
  return int(nb < 1 ? 1 : (nb > FS_MAX_BUCKETS ? FS_MAX_BUCKETS : nb));
}
static inline int64_t fs_tiles(int64_t n) { return n > 0 ? (n + FS_PART_TILE - 1) / FS_PART_TILE : 1; }

struct FastSortWs {
  // zeroed by the caller before every sort
  uint32_t* hist12;        // [FS_BINS]
  uint32_t* bucket_count;  // [FS_MAX_BUCKETS]
  uint32_t* counters;      // [4]: blocks done (histogram, count), partition tile ticket
  int32_t* fallback;       // [1] set when the bucket map could not balance the input: the LSD sort must run
  uint32_t* lookback;      // [fs_tiles][FS_MAX_BUCKETS]
  // not zeroed
  uint2* lut;              // [FS_BINS] (exclusive prefix, count) of every 12-bit bin
  uint32_t* bucket_base;   // [FS_MAX_BUCKETS + 1]
  uint4* work;             // [FS_MAX_WORK] (bucket, first element, count, bucket size)
  uint32_t* params;        // [4]: number of work items
  uint32_t* keys;          // [n] partitioned keys
  uint32_t* vals;          // [n] partitioned payloads (index | event << 31)
};

// Enqueue histogram (+ max / NaN flag over `scores` when given), count, partition and local sort.  Writes
// perm_out (index | event bit) and, when `scores` is given, saved_s = scores[perm] - max and the list of argmax
// positions.  On inputs the bucket map cannot balance, *w.fallback is set instead (outputs undefined).
int fs_sort_enqueue(const float* times, const float* status, const float* scores, int64_t n, const FastSortWs& w,
                    uint32_t* max_enc, int32_t* nan_flag, int32_t* nonbinary_flag, int32_t* perm_out, float* saved_s,
                    int32_t* max_count, int32_t* max_list, int max_list_cap, cudaStream_t stream);

}  // namespace mmbs
