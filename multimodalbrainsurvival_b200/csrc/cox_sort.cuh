// Bucketed risk-set pipeline of the Cox loss for 2048 < n <= FS_MAX_N (see cox_sort.cu): equalised-CDF partition into
// buckets of ~5 K samples, one block per bucket sorts it in shared memory and runs the forward pass on it, one block
// per bucket runs the backward pass.
#pragma once

#include "radix_sort.cuh"

namespace mmbs {

constexpr int FS_BINS = 4096;              // histogram of the top 12 key bits
constexpr int FS_MAX_BUCKETS = 4096;
constexpr int FS_CAP = 8192;               // samples one block sorts in shared memory (= slots per bucket region)
constexpr int FS_TARGET = 6144;            // aimed bucket size (12 of the 16 rows of a block): 1.33x head-room below FS_CAP
constexpr int FS_LOG_S = 12;               // 4096 sub-buckets per bucket (~1.25 samples each)
constexpr int FS_SQ_BUDGET = 1 << 20;      // rank-by-comparison finish: accepted sum of (sub-bucket size)^2 over the sub-buckets
                                           // of more than 8 samples (one run of ~1000 tied times, or the bucket that ends at
                                           // t -> 0 and spans a hundred binades); beyond: the LSD pipeline
constexpr int FS_PART = FS_CAP / 32;        // partial sums of exp(s~) per bucket: one per 32 sorted positions
constexpr int FS_MAX_LIST = 1024;          // argmax positions remembered for the gradient through max(scores)
constexpr int64_t FS_MAX_N = int64_t(FS_MAX_BUCKETS) * FS_TARGET;   // 25.2 M; beyond: LSD sort

// MMBS_COX_HIST_SAMPLE=<shift> overrides (0 = read everything)
int fs_sample_shift(int64_t n);
// aimed bucket size: FS_TARGET unless MMBS_COX_TARGET=<samples> overrides it (experiments)
int fs_target();

struct FsPlan {
  int nb;             // buckets
  int sample_shift;   // the histogram reads one 1024-sample chunk out of 2^sample_shift (large n: the CDF table only has
                      // to balance the buckets, ~5 % noise on 5 K-sample buckets is far inside the 1.6x head-room)
};
static inline FsPlan fs_plan(int64_t n) {
  FsPlan p;
  const int64_t target = fs_target();
  int64_t nb = (n + target - 1) / target;
  p.nb = int(nb < 1 ? 1 : (nb > FS_MAX_BUCKETS ? FS_MAX_BUCKETS : nb));
  p.sample_shift = fs_sample_shift(n);
  return p;
}

struct FsEdge {       // the two 12-bit bins that hold the smallest / largest key are interpolated over the occupied
  uint32_t lo_bin;    // key range only (a bin half covered by the data would otherwise give buckets of twice the size)
  uint32_t lo_base;
  float lo_scale;
  uint32_t hi_bin;
  uint32_t hi_base;
  float hi_scale;
  uint32_t mult;      // floor(nb * 2^32 / m): rank estimate in [0, m) -> bucket; m = samples the histogram counted
  uint32_t pad;
};

struct FastSortWs {
  // zeroed by the caller before every forward pass
  uint32_t* hist12;        // [FS_BINS]
  uint32_t* kext;          // [2]: max(~key), max(key)
  uint32_t* counters;      // [8]: 0 histogram blocks done, 1 bucket-sort blocks done
  uint32_t* cursor;        // [FS_MAX_BUCKETS] samples written to every bucket region
  int32_t* fallback;       // [1] set when this pipeline gave up: the LSD-sort pipeline must (re)do the work
  // not zeroed
  uint2* lut;              // [FS_BINS] (exclusive prefix, count) of every 12-bit bin
  FsEdge* edge;            // [1]
  double* agg_val;         // [FS_MAX_BUCKETS] sum of exp(s~) over the bucket
  double* exp_prefix;      // [FS_MAX_BUCKETS] sum of exp(s~) over all earlier buckets           (kept for backward)
  double* wsum;            // [FS_MAX_BUCKETS] sum of w = status / (C + eps) over all LATER buckets (kept for backward)
  uint32_t* bucket_base;   // [FS_MAX_BUCKETS] first sorted position of the bucket               (kept for backward)
  uint32_t* bucket_cnt;    // [FS_MAX_BUCKETS]                                                    (kept for backward)
  uint2* pairs;            // [nb * FS_CAP] partitioned (key, index | event << 31)
  float* part32;           // [FS_MAX_BUCKETS * FS_PART] sums of exp(s~) over 32 sorted positions       (kept for backward)
  double* row_loss;        // [FS_MAX_BUCKETS * 16] per warp row: loss terms, w, gradient sums
  double* row_w;           //                                                                    (kept for backward)
  double* row_g;
  float* sc_part;          // [nb * FS_CAP] the samples' scores in the same slots (the bucket sort never gathers)
};

// Forward: histogram (+ max / NaN flag over `scores`) -> partition -> per-bucket sort + gather + scan + loss.
// With scores == nullptr only the permutation is produced (mmbs_risk_order).
int fs_forward_enqueue(const float* times, const float* status, const float* scores, int64_t n, const FastSortWs& w,
                       uint32_t* max_enc, int32_t* nan_flag, int32_t* nonbinary_flag, int32_t* perm_out, float* saved_s,
                       int32_t* max_count, int32_t* max_list, float* loss_out, int32_t* flags_out, cudaStream_t stream);

// Backward: per bucket, recompute C and w from s~, suffix sums, gradient scatter; the last block applies the gradient
// through max(scores) when its argmax positions fit the list (otherwise cox_maxfix_full_kernel does, cox.cu).
int fs_backward_enqueue(const float* status, const int32_t* perm, const float* saved_s, const float* grad_loss, int64_t n,
                        const FastSortWs& w, const int32_t* nonbinary_flag, const int32_t* max_count,
                        const int32_t* max_list, double* gsum_total, float* grad_scores, cudaStream_t stream);

}  // namespace mmbs
