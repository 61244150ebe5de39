// Bucketed risk-set pipeline of the Cox loss for 2048 < n <= FS_MAX_N (see cox_sort.cu): equalised-CDF partition into
// buckets of ~6 K samples, one block per bucket sorts it in shared memory, one warp per 512 sorted positions runs the
// forward / backward scans, one-block passes in between carry the sums across buckets.
#pragma once

#include "radix_sort.cuh"

namespace mmbs {

constexpr int FS_BINS = 4096;              // histogram of the top 12 key bits
constexpr int FS_MAX_BUCKETS = 4096;
constexpr int FS_CAP = 8192;               // samples one block sorts in shared memory (= slots per bucket region)
constexpr int FS_TARGET = 6144;            // aimed bucket size (12 of the 16 rows of a block): 1.33x head-room below FS_CAP
#ifndef MMBS_FS_LOG_S
#define MMBS_FS_LOG_S 12
#endif
constexpr int FS_LOG_S = MMBS_FS_LOG_S;    // 2^12 = 4096 sub-buckets per bucket (~1.5 samples each)
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

#ifdef __CUDACC__
constexpr int FS_FINAL_THREADS = 512;    // block size of the one-block finishing passes (<= 128 registers per thread: the
                                         // dispatch kernels call the separately compiled rs_sort_tail_launch, 106 registers)
constexpr int FS_ROWS = 16;              // warp rows per bucket (cox_sort.cu R_ROWS)

// inclusive scan of one double per thread over the block, forward (lower threads first) or reverse; *total = block sum
template <int NW, bool REVERSE>
__device__ __forceinline__ double block_scan_f64(double v, double* s_red, int lane, int warp, double* total) {
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = REVERSE ? __shfl_down_sync(0xffffffffu, incl, o) : __shfl_up_sync(0xffffffffu, incl, o);
    if (REVERSE ? (lane + o < 32) : (lane >= o)) incl += t;
  }
  __syncthreads();
  if (lane == (REVERSE ? 0 : 31)) s_red[warp] = incl;
  __syncthreads();
  double off = 0.0, tot = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const double c = s_red[w];
    if (REVERSE ? (w > warp) : (w < warp)) off += c;
    tot += c;
  }
  *total = tot;
  return incl + off;
}

template <int NW>
__device__ __forceinline__ double block_sum_f64(double v, double* s_red, int lane, int warp) {
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += s_red[w];
  return t;
}


// Finish of the forward pass, ONE block of FS_FINAL_THREADS threads: loss = sum of the row partials / n (.mean() over N,
// models.py:111); wsuffix[b] = sum of w over all LATER buckets.  Element e = 16 b + r of the row arrays goes to thread
// e mod 1024: coalesced, every load independent; a bucket's 16 rows sit in one half-warp.
__device__ __forceinline__ void fs_loss_finalize_block(const double* __restrict__ row_loss, const double* __restrict__ row_w,
                                                       int nb, int64_t n, double* __restrict__ wsuffix,
                                                       const int32_t* __restrict__ nan_flag, float* __restrict__ loss_out,
                                                       int32_t* __restrict__ flags_out) {
  constexpr int PER = FS_MAX_BUCKETS / FS_FINAL_THREADS;
  __shared__ double s_red[FS_FINAL_THREADS / 32];
  __shared__ double s_wb[FS_MAX_BUCKETS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = nb * FS_ROWS;
  constexpr int BATCH = 8;
  double lsum = 0.0;
  for (int e0 = 0; e0 < total; e0 += BATCH * FS_FINAL_THREADS) {   // block-uniform trip count (shuffles inside)
    double l[BATCH], w[BATCH];
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int e = e0 + k * FS_FINAL_THREADS + tid;
      l[k] = e < total ? __ldg(row_loss + e) : 0.0;
      w[k] = e < total ? __ldg(row_w + e) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      lsum += l[k];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);   // over the 16 rows of the bucket
      const int e = e0 + k * FS_FINAL_THREADS + tid;
      if ((lane & 15) == 0 && e < FS_MAX_BUCKETS * FS_ROWS) s_wb[e >> 4] = w[k];
    }
  }
  __syncthreads();
  double w[PER], wsum = 0.0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int q = tid * PER + k;
    w[k] = q < nb ? s_wb[q] : 0.0;
    wsum += w[k];
  }
  double tot;
  double later = block_scan_f64<FS_FINAL_THREADS / 32, true>(wsum, s_red, lane, warp, &tot) - wsum;   // higher threads only
#pragma unroll
  for (int k = PER - 1; k >= 0; --k) {
    const int q = tid * PER + k;
    if (q < nb) wsuffix[q] = later;
    later += w[k];
  }
  const double t = block_sum_f64<FS_FINAL_THREADS / 32>(lsum, s_red, lane, warp);
  if (tid == 0) {
    const int f = *nan_flag;
    loss_out[0] = f ? __int_as_float(0x7fc00000) : float(t / double(n));
    if (flags_out) flags_out[0] = f;
  }
}

// Finish of the backward pass, ONE block: gradient through "- max(scores)": every argmax position receives
// -(sum_k g~_k) / count (when the positions fit the list; otherwise cox_maxfix_full_kernel reads gsum_total)
__device__ __forceinline__ void fs_backward_finalize_block(const double* __restrict__ row_g, int nb,
                                                           const int32_t* __restrict__ max_count,
                                                           const int32_t* __restrict__ max_list,
                                                           double* __restrict__ gsum_total, float* grad_scores) {
  __shared__ double s_red[FS_FINAL_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double g = 0.0;
  const double2* pg = reinterpret_cast<const double2*>(row_g);
  for (int i0 = 0; i0 < nb * (FS_ROWS / 2); i0 += 8 * FS_FINAL_THREADS) {
    double2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * FS_FINAL_THREADS + tid;
      a[k] = i < nb * (FS_ROWS / 2) ? __ldg(pg + i) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) g += a[k].x + a[k].y;
  }
  const double t = block_sum_f64<FS_FINAL_THREADS / 32>(g, s_red, lane, warp);
  if (tid == 0) gsum_total[0] = t;
  const int mc = *max_count;
  if (mc <= FS_MAX_LIST) {
    const float fix = float(t / double(mc));
    for (int i = tid; i < mc; i += FS_FINAL_THREADS) grad_scores[max_list[i]] -= fix;
  }
}
#endif

}  // namespace mmbs
