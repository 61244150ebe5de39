// Per-patient aggregation as a segmented mean (sm_100a).
//
// Replaces the aggregation tails of extract_features()
// (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:75-89: for every case,
//  boolean-mask . features / count - O(cases*N*2048) on the CPU) and of get_survival_CI()
// (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:126-152: per-id mean of the
//  scores, "last row seen" survival/vital).
//
// The rows are grouped by a stable radix sort of (segment id, row index) - the same
// Onesweep passes as the Cox sort - so every segment owns a contiguous, ascending
// list of row indices; one pass over `values` (coalesced 128-bit row reads) then
// accumulates each segment in ascending row order in fp32.  Bytes moved:
// N*(4*D+4) + G*4*D plus 20 B/row of sort traffic.
#include <algorithm>

#include "radix_sort.cuh"

namespace mmbs {

__global__ void __launch_bounds__(256) seg_count_kernel(const int32_t* __restrict__ seg, int64_t n,
                                                        int64_t n_seg, int32_t* __restrict__ counts,
                                                        int32_t* __restrict__ bad) {
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
    const int32_t g = seg[i];
    if (g < 0 || g >= n_seg) {
      atomicOr(bad, 1);
    } else {
      atomicAdd(counts + g, 1);
    }
  }
}

// single-block exclusive scan counts -> offsets[n_seg+1].
// A segment id outside [0, n_seg) (flag `bad`) poisons the whole result instead of shifting every later
// segment's row list silently: all segments become empty (mean = 0/0 = NaN, last_row = -1) and counts = -1.
__global__ void __launch_bounds__(1024) seg_offsets_kernel(int32_t* __restrict__ counts,
                                                           int64_t n_seg, int32_t* __restrict__ offsets,
                                                           const int32_t* __restrict__ bad) {
  __shared__ int32_t s_w[32];
  __shared__ int32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  const bool poisoned = (*bad != 0);
  __syncthreads();
  for (int64_t b = 0; b < n_seg; b += 1024) {
    const int64_t i = b + tid;
    if (poisoned && i < n_seg) counts[i] = -1;
    const int32_t c = (i < n_seg && !poisoned) ? counts[i] : 0;
    int32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int32_t off = s_carry;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (i < n_seg) offsets[i] = off + incl - c;
    __syncthreads();
    if (tid == 1023) s_carry = off + incl;
    __syncthreads();
  }
  if (tid == 0) offsets[n_seg] = s_carry;
}

// d % 4 == 0: one thread per float4 column, rows of the segment in ascending order.
__global__ void __launch_bounds__(128) seg_mean_wide_kernel(
    const float* __restrict__ values, const int32_t* __restrict__ rows,
    const int32_t* __restrict__ offsets, int64_t d4, float* __restrict__ out,
    int32_t* __restrict__ last_row) {
  const int64_t g = blockIdx.x;
  const int64_t col = int64_t(blockIdx.y) * 128 + threadIdx.x;
  const int32_t beg = offsets[g], end = offsets[g + 1];
  if (blockIdx.y == 0 && threadIdx.x == 0 && last_row) last_row[g] = end > beg ? rows[end - 1] : -1;
  if (col >= d4) return;
  const float4* v = reinterpret_cast<const float4*>(values);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int32_t r = beg;
  for (; r + 4 <= end; r += 4) {  // 4 independent row loads in flight, summed in order
    const int32_t i0 = __ldg(rows + r), i1 = __ldg(rows + r + 1), i2 = __ldg(rows + r + 2), i3 = __ldg(rows + r + 3);
    const float4 a = __ldg(v + int64_t(i0) * d4 + col);
    const float4 b = __ldg(v + int64_t(i1) * d4 + col);
    const float4 c = __ldg(v + int64_t(i2) * d4 + col);
    const float4 e = __ldg(v + int64_t(i3) * d4 + col);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
    acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
  }
  for (; r < end; ++r) {
    const float4 a = __ldg(v + int64_t(__ldg(rows + r)) * d4 + col);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  const float cntf = float(end - beg);  // empty segment -> 0/0 = NaN like numpy
  float4 o;
  o.x = acc.x / cntf; o.y = acc.y / cntf; o.z = acc.z / cntf; o.w = acc.w / cntf;
  reinterpret_cast<float4*>(out)[g * d4 + col] = o;
}

// any d (scores: d == 1): one warp per (segment, column); lanes stride the rows.
__global__ void __launch_bounds__(128) seg_mean_narrow_kernel(
    const float* __restrict__ values, const int32_t* __restrict__ rows,
    const int32_t* __restrict__ offsets, int64_t d, int64_t n_seg, float* __restrict__ out,
    int32_t* __restrict__ last_row) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t(blockIdx.x) * 128 + threadIdx.x) >> 5;
  if (wid >= n_seg * d) return;
  const int64_t g = wid / d, col = wid % d;
  const int32_t beg = offsets[g], end = offsets[g + 1];
  if (col == 0 && lane == 0 && last_row) last_row[g] = end > beg ? rows[end - 1] : -1;
  float acc = 0.f;
  for (int32_t r = beg + lane; r < end; r += 32) acc += __ldg(values + int64_t(__ldg(rows + r)) * d + col);
  acc = warp_sum(acc);
  if (lane == 0) out[g * d + col] = acc / float(end - beg);
}

struct SegWorkspace {
  uint32_t* hist; uint32_t* counters; int32_t* bad; uint32_t* lookback;  // zeroed
  size_t zero_bytes;
  uint32_t* digit_base; int32_t* offsets; int32_t* rows;
  uint32_t* keys_a; uint32_t* keys_b; uint32_t* vals_a; uint32_t* vals_b;
  size_t total_bytes;
};

static SegWorkspace carve_seg(void* base, int64_t n, int64_t n_seg) {
  const int64_t rt = rs_tiles(n);
  Carver c(base);
  SegWorkspace w;
  w.hist = c.take<uint32_t>(4 * RS_RADIX);
  w.counters = c.take<uint32_t>(8);
  w.bad = c.take<int32_t>(1);
  w.lookback = c.take<uint32_t>(size_t(4) * rt * RS_RADIX);
  w.zero_bytes = align_up(c.off, 256);
  w.digit_base = c.take<uint32_t>(4 * RS_RADIX);
  w.offsets = c.take<int32_t>(n_seg + 1);
  w.rows = c.take<int32_t>(n);
  w.keys_a = c.take<uint32_t>(n);
  w.keys_b = c.take<uint32_t>(n);
  w.vals_a = c.take<uint32_t>(n);
  w.vals_b = c.take<uint32_t>(n);
  w.total_bytes = align_up(c.off, 256);
  return w;
}

}  // namespace mmbs

using namespace mmbs;

extern "C" size_t mmbs_segmented_mean_workspace_bytes(int64_t n, int64_t n_seg) {
  return carve_seg(nullptr, std::max<int64_t>(n, 1), std::max<int64_t>(n_seg, 1)).total_bytes;
}

extern "C" int mmbs_segmented_mean(const float* values, const int32_t* seg_ids, int64_t n, int64_t d,
                                   int64_t n_seg, float* out, int32_t* counts, int32_t* last_row,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(values && seg_ids && out && counts && workspace, "mmbs_segmented_mean: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N && d >= 1 && n_seg >= 1 && n_seg <= RS_MAX_N,
               "mmbs_segmented_mean: bad sizes n=%lld d=%lld n_seg=%lld", (long long)n, (long long)d,
               (long long)n_seg);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const SegWorkspace w = carve_seg(workspace, n, n_seg);
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_segmented_mean: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  MMBS_CUDA_TRY(cudaMemsetAsync(w.hist, 0, w.zero_bytes, stream));
  MMBS_CUDA_TRY(cudaMemsetAsync(counts, 0, size_t(n_seg) * sizeof(int32_t), stream));
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256 * 4), int64_t(sm_count()) * 8)));
  seg_count_kernel<<<grid, 256, 0, stream>>>(seg_ids, n, n_seg, counts, w.bad);
  MMBS_LAUNCH_CHECK();
  seg_offsets_kernel<<<1, 1024, 0, stream>>>(counts, n_seg, w.offsets, w.bad);
  MMBS_LAUNCH_CHECK();
  int passes = 1;
  while (passes < 4 && (uint64_t(n_seg - 1) >> (8 * passes)) != 0) ++passes;
  if (int rc = rs_histogram_enqueue(seg_ids, KEY_U32, n, passes, w.hist, w.digit_base, nullptr, nullptr,
                                    nullptr, stream))
    return rc;
  SortWorkspace s;
  s.keys_a = w.keys_a; s.keys_b = w.keys_b; s.vals_a = w.vals_a; s.vals_b = w.vals_b;
  s.hist = w.hist; s.digit_base = w.digit_base; s.counters = w.counters; s.lookback = w.lookback;
  if (int rc = rs_sort_enqueue(seg_ids, KEY_U32, n, passes, s, w.rows, stream)) return rc;
  if (d % 4 == 0 && (reinterpret_cast<uintptr_t>(values) % 16 == 0) &&
      (reinterpret_cast<uintptr_t>(out) % 16 == 0)) {
    const int64_t d4 = d / 4;
    dim3 g(unsigned(n_seg), unsigned(ceil_div(d4, 128)));
    MMBS_REQUIRE(n_seg <= 0x7fffffff && ceil_div(d4, 128) <= 65535, "mmbs_segmented_mean: grid too large");
    seg_mean_wide_kernel<<<g, 128, 0, stream>>>(values, w.rows, w.offsets, d4, out, last_row);
  } else {
    const int64_t warps = n_seg * d;
    seg_mean_narrow_kernel<<<unsigned(ceil_div(warps, 4)), 128, 0, stream>>>(values, w.rows, w.offsets, d,
                                                                            n_seg, out, last_row);
  }
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
