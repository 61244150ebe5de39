// Cox negative log partial likelihood, forward and backward, for sm_100a.
//
// Replaces cox_loss() of the reference (/root/reference/1_HistoPathology/models.py:90-111
// and its three textual copies):
//   sort(-times) -> gather -> subtract max -> exp -> cumsum -> log(.+1e-5) -> mask -> mean.
//
// Forward  = histogram(+max of scores)  ->  4 Onesweep radix passes  ->  one chained
//            scan pass over the sorted order (decoupled look-back, fp64 carries).
// Backward = one reverse chained scan pass (suffix sums of status/(C+eps)) that
//            scatters the gradient, + the gradient through max(scores).
// All passes are HBM-bound: coalesced 128-bit loads/stores on the streamed
// arrays, the only random accesses are the 4-byte gathers through the permutation.
#include <algorithm>

#include "radix_sort.cuh"

namespace mmbs {

constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS = 8;
constexpr int CS_TILE = CS_THREADS * CS_ITEMS;  // 2048 sorted positions per tile
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int COX_MAX_LIST = 1024;
constexpr float COX_EPS = 1e-5f;

static inline int64_t cs_tiles(int64_t n) { return n > 0 ? (n + CS_TILE - 1) / CS_TILE : 1; }

// Chained-scan tile status: ONE 64-bit word per tile = a double whose two mantissa LSBs are
// replaced by the state (0 = empty, 1 = tile aggregate, 2 = inclusive prefix).  Value and flag
// travel in a single atomic 8-byte store/load, so the look-back needs no memory fences; the
// carry keeps 50 mantissa bits.
struct ScanState {
  unsigned long long* status;
};
__device__ __forceinline__ unsigned long long pack_status(double v, unsigned flag) {
  return (static_cast<unsigned long long>(__double_as_longlong(v)) & ~3ull) | flag;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Warp-parallel decoupled look-back: returns the sum of all earlier tiles' totals.
__device__ __forceinline__ double lookback_sum(const ScanState& st, int64_t tile, int lane) {
  double excl = 0.0;
  int64_t pred = tile - 1;
  while (true) {
    const int64_t idx = pred - lane;
    unsigned long long w = 2ull;  // tiles before tile 0: inclusive prefix 0
    if (idx >= 0) {
      do {
        w = ld_volatile_u64(st.status + idx);
      } while ((w & 3ull) == 0ull);
    }
    const unsigned incl_mask = __ballot_sync(0xffffffffu, (w & 3ull) == 2ull);
    const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
    const double v = (lane <= first_incl) ? __longlong_as_double(static_cast<long long>(w & ~3ull)) : 0.0;
    excl += warp_sum(v);
    if (incl_mask) break;
    pred -= 32;
  }
  return excl;
}

// ------------------------------------------------------------------ forward scan
__global__ void __launch_bounds__(CS_THREADS) cox_scan_fwd_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ scores,
    const float* __restrict__ status, const uint32_t* __restrict__ max_enc, int64_t n,
    float* __restrict__ saved_e, float* __restrict__ saved_w, ScanState st, uint32_t* tile_counter,
    double* __restrict__ loss_partial, int32_t* nan_flag, int32_t* max_count,
    int32_t* __restrict__ max_list, const int32_t* __restrict__ nonbinary) {
  __shared__ double s_warp_tot[CS_WARPS];
  __shared__ double s_red[CS_WARPS];
  __shared__ double s_block_excl;
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t base = tile * CS_TILE + int64_t(tid) * CS_ITEMS;
  const float smax = float_order_dec(*max_enc);

  int32_t p[CS_ITEMS];
  if (base + CS_ITEMS <= n) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(perm + base));
    const int4 b = __ldg(reinterpret_cast<const int4*>(perm + base) + 1);
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;
    p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) p[j] = (base + j < n) ? __ldg(perm + base + j) : 0;
  }
  // perm words: bits 0-30 original index, bit 31 = event indicator (binary status)
  const bool gather_status = (*nonbinary != 0);
  const uint64_t pol = make_evict_last_policy();
  bool valid[CS_ITEMS];
  float sc[CS_ITEMS], dl[CS_ITEMS], e[CS_ITEMS], c[CS_ITEMS];
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    valid[j] = (base + j < n);
    const uint32_t word = uint32_t(p[j]);
    p[j] = int32_t(word & 0x7fffffffu);
    dl[j] = (valid[j] && (word >> 31)) ? 1.f : 0.f;
  }
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {  // the random 4-byte gather (scores stay L2-resident: evict_last)
    sc[j] = valid[j] ? ld_f32_hint(scores + p[j], pol) : 0.f;
    if (gather_status) dl[j] = valid[j] ? __ldg(status + p[j]) : 0.f;
  }
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    sc[j] -= smax;                                  // s~ (models.py:102)
    e[j] = valid[j] ? expf(sc[j]) : 0.f;            // models.py:103
    run += e[j];
    c[j] = run;
    if (valid[j] && sc[j] == 0.f) {                 // an argmax position (for backward)
      const int pos = atomicAdd(max_count, 1);
      if (pos < COX_MAX_LIST) max_list[pos] = p[j];
    }
  }
  // block-wide exclusive prefix of the per-thread totals, carried in fp64
  const double tsum = double(run);
  double winc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  if (lane == 31) s_warp_tot[warp] = winc;
  __syncthreads();
  double warp_excl = 0.0, block_total = 0.0;
#pragma unroll
  for (int w = 0; w < CS_WARPS; ++w) {
    const double t = s_warp_tot[w];
    if (w < warp) warp_excl += t;
    block_total += t;
  }
  if (warp == 0) {
    if (lane == 0) st_volatile_u64(st.status + tile, pack_status(block_total, tile == 0 ? 2u : 1u));
    double excl = 0.0;
    if (tile > 0) {
      excl = lookback_sum(st, tile, lane);
      if (lane == 0) st_volatile_u64(st.status + tile, pack_status(excl + block_total, 2u));
    }
    if (lane == 0) s_block_excl = excl;
  }
  __syncthreads();
  const double off = s_block_excl + warp_excl + (winc - tsum);

  float w[CS_ITEMS];
  float lsum = 0.f;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const float cj = float(off + double(c[j]));      // cumsum (models.py:104)
    const float den = cj + COX_EPS;
    const float term = -(sc[j] - logf(den)) * dl[j];  // models.py:104-105
    w[j] = dl[j] / den;
    if (valid[j]) {
      lsum += term;
      bad |= (term != term);
    }
  }
  if (base + CS_ITEMS <= n) {
    float4* pe = reinterpret_cast<float4*>(saved_e + base);
    float4* pw = reinterpret_cast<float4*>(saved_w + base);
    pe[0] = make_float4(e[0], e[1], e[2], e[3]);
    pe[1] = make_float4(e[4], e[5], e[6], e[7]);
    pw[0] = make_float4(w[0], w[1], w[2], w[3]);
    pw[1] = make_float4(w[4], w[5], w[6], w[7]);
  } else {
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j)
      if (base + j < n) {
        saved_e[base + j] = e[j];
        saved_w[base + j] = w[j];
      }
  }
  double bl = warp_sum(double(lsum));
  if (lane == 0) s_red[warp] = bl;
  const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0 && any_bad) atomicOr(nan_flag, 1);
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < CS_WARPS; ++w2) t += s_red[w2];
    loss_partial[tile] = t;
  }
}

__global__ void __launch_bounds__(256) cox_finalize_kernel(const double* __restrict__ partial,
                                                           int64_t tiles, int64_t n,
                                                           const int32_t* __restrict__ nan_flag,
                                                           float* __restrict__ loss_out,
                                                           int32_t* __restrict__ flags_out) {
  __shared__ double s_red[8];
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < tiles; i += 256) t += partial[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    const int f = *nan_flag;
    loss_out[0] = f ? __int_as_float(0x7fc00000) : float(s / double(n));  // .mean() over N (models.py:111)
    if (flags_out) flags_out[0] = f;
  }
}

// ------------------------------------------------------------------ backward scan
// Tiles walk the sorted order from the end; tile j covers
// k in [n_pad-(j+1)*T, n_pad-j*T) so every thread's 8 positions stay 32-byte aligned.
__global__ void __launch_bounds__(CS_THREADS) cox_scan_bwd_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ status,
    const float* __restrict__ saved_e, const float* __restrict__ saved_w,
    const float* __restrict__ grad_loss, int64_t n, int64_t n_pad, float* __restrict__ grad_scores,
    ScanState st, uint32_t* tile_counter, double* __restrict__ gsum_partial,
    const int32_t* __restrict__ nonbinary) {
  __shared__ double s_warp_tot[CS_WARPS];
  __shared__ double s_red[CS_WARPS];
  __shared__ double s_block_excl;
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t lo = n_pad - (tile + 1) * CS_TILE;
  const int64_t base = lo + int64_t(CS_THREADS - 1 - tid) * CS_ITEMS;  // thread 0 = highest k
  const float scale = grad_loss[0] / float(n);

  float e[CS_ITEMS], w[CS_ITEMS];
  int32_t p[CS_ITEMS];
  if (base + CS_ITEMS <= n) {
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(saved_e + base));
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(saved_e + base) + 1);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(saved_w + base));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(saved_w + base) + 1);
    const int4 a = __ldg(reinterpret_cast<const int4*>(perm + base));
    const int4 b = __ldg(reinterpret_cast<const int4*>(perm + base) + 1);
    e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w; e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
    w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) {
      const bool v = base + j < n;
      e[j] = v ? __ldg(saved_e + base + j) : 0.f;
      w[j] = v ? __ldg(saved_w + base + j) : 0.f;
      p[j] = v ? __ldg(perm + base + j) : 0;
    }
  }
  const bool gather_status = (*nonbinary != 0);
  bool valid[CS_ITEMS];
  float dl[CS_ITEMS];
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    valid[j] = (base + j < n);
    const uint32_t word = uint32_t(p[j]);
    p[j] = int32_t(word & 0x7fffffffu);
    dl[j] = (valid[j] && (word >> 31)) ? 1.f : 0.f;
    if (gather_status) dl[j] = valid[j] ? __ldg(status + p[j]) : 0.f;
  }

  float suf[CS_ITEMS];  // inclusive suffix sums inside the thread (descending k)
  float run = 0.f;
#pragma unroll
  for (int j = CS_ITEMS - 1; j >= 0; --j) {
    run += w[j];
    suf[j] = run;
  }
  const double tsum = double(run);
  double winc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  if (lane == 31) s_warp_tot[warp] = winc;
  __syncthreads();
  double warp_excl = 0.0, block_total = 0.0;
#pragma unroll
  for (int w2 = 0; w2 < CS_WARPS; ++w2) {
    const double t = s_warp_tot[w2];
    if (w2 < warp) warp_excl += t;
    block_total += t;
  }
  if (warp == 0) {
    if (lane == 0) st_volatile_u64(st.status + tile, pack_status(block_total, tile == 0 ? 2u : 1u));
    double excl = 0.0;
    if (tile > 0) {
      excl = lookback_sum(st, tile, lane);
      if (lane == 0) st_volatile_u64(st.status + tile, pack_status(excl + block_total, 2u));
    }
    if (lane == 0) s_block_excl = excl;
  }
  __syncthreads();
  const double off = s_block_excl + warp_excl + (winc - tsum);

  float gs = 0.f;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    if (valid[j]) {
      const float W = float(off + double(suf[j]));      // sum_{i>=k} status_i/(C_i+eps)
      const float g = -(dl[j] - e[j] * W) * scale;
      grad_scores[p[j]] = g;                            // un-permute
      gs += g;
    }
  }
  double bl = warp_sum(double(gs));
  if (lane == 0) s_red[warp] = bl;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < CS_WARPS; ++w2) t += s_red[w2];
    gsum_partial[tile] = t;
  }
}

// Gradient through "- max(scores)": every argmax position receives
// -(sum_k g~_k)/count  (torch's full-reduction max backward splits evenly).
__global__ void __launch_bounds__(256) cox_maxfix_list_kernel(
    const double* __restrict__ gsum_partial, int64_t tiles, const int32_t* __restrict__ max_count,
    const int32_t* __restrict__ max_list, double* __restrict__ gsum_total,
    float* __restrict__ grad_scores) {
  __shared__ double s_red[8];
  __shared__ double s_total;
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < tiles; i += 256) t += gsum_partial[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    s_total = s;
    gsum_total[0] = s;
  }
  __syncthreads();
  const int cnt = *max_count;
  if (cnt <= COX_MAX_LIST) {
    const float fix = float(s_total / double(cnt));
    for (int i = threadIdx.x; i < cnt; i += 256) grad_scores[max_list[i]] -= fix;
  }
}

__global__ void __launch_bounds__(256) cox_maxfix_full_kernel(
    const float* __restrict__ scores, const uint32_t* __restrict__ max_enc,
    const int32_t* __restrict__ max_count, const double* __restrict__ gsum_total, int64_t n,
    float* __restrict__ grad_scores) {
  const int cnt = *max_count;
  if (cnt <= COX_MAX_LIST) return;  // handled by the list kernel
  const float smax = float_order_dec(*max_enc);
  const float fix = float(gsum_total[0] / double(cnt));
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256)
    if (scores[i] - smax == 0.f) grad_scores[i] -= fix;
}

// ------------------------------------------------------------------ workspace
struct CoxWorkspace {
  // zeroed at the start of forward (contiguous)
  uint32_t* hist;          // [4][256]
  uint32_t* counters;      // [8]: 0-3 radix passes, 4 fwd scan, 5 bwd scan
  uint32_t* max_enc;       // [1]
  int32_t* nan_flag;       // [1]
  int32_t* max_count;      // [1]
  int32_t* nonbinary;      // [1] some status value is neither 0 nor 1
  uint32_t* lookback;      // [4][rs_tiles][256]
  unsigned long long* status_fwd;  // [cs_tiles]
  size_t zero_bytes;
  // zeroed at the start of backward
  unsigned long long* status_bwd;  // [cs_tiles]
  // not zeroed
  uint32_t* digit_base;    // [4][256]
  int32_t* max_list;       // [COX_MAX_LIST]
  double* gsum_total;      // [1]
  double* loss_partial;
  double* gsum_partial;
  uint32_t* keys_a; uint32_t* keys_b; uint32_t* vals_a; uint32_t* vals_b;
  size_t total_bytes;
};

static CoxWorkspace carve_cox(void* base, int64_t n) {
  const int64_t rt = rs_tiles(n), ct = cs_tiles(n);
  Carver c(base);
  CoxWorkspace w;
  w.hist = c.take<uint32_t>(4 * RS_RADIX);
  w.counters = c.take<uint32_t>(8);
  w.max_enc = c.take<uint32_t>(1);
  w.nan_flag = c.take<int32_t>(1);
  w.max_count = c.take<int32_t>(1);
  w.nonbinary = c.take<int32_t>(1);
  w.lookback = c.take<uint32_t>(size_t(4) * rt * RS_RADIX);
  w.status_fwd = c.take<unsigned long long>(ct);
  w.zero_bytes = align_up(c.off, 256);
  w.status_bwd = c.take<unsigned long long>(ct);
  w.digit_base = c.take<uint32_t>(4 * RS_RADIX);
  w.max_list = c.take<int32_t>(COX_MAX_LIST);
  w.gsum_total = c.take<double>(1);
  w.loss_partial = c.take<double>(ct);
  w.gsum_partial = c.take<double>(ct);
  w.keys_a = c.take<uint32_t>(n);
  w.keys_b = c.take<uint32_t>(n);
  w.vals_a = c.take<uint32_t>(n);
  w.vals_b = c.take<uint32_t>(n);
  w.total_bytes = align_up(c.off, 256);
  return w;
}

static SortWorkspace sort_ws(const CoxWorkspace& w) {
  SortWorkspace s;
  s.keys_a = w.keys_a; s.keys_b = w.keys_b; s.vals_a = w.vals_a; s.vals_b = w.vals_b;
  s.hist = w.hist; s.digit_base = w.digit_base; s.counters = w.counters; s.lookback = w.lookback;
  return s;
}

}  // namespace mmbs

using namespace mmbs;

extern "C" size_t mmbs_cox_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return carve_cox(nullptr, n).total_bytes;
}

static int cox_sort_common(const float* scores, const float* times, const float* status, int64_t n,
                           int32_t* perm_out, const CoxWorkspace& w, cudaStream_t stream) {
  MMBS_CUDA_TRY(cudaMemsetAsync(w.hist, 0, w.zero_bytes, stream));
  int rc = rs_histogram_enqueue(times, KEY_NEG_TIME_F32, n, 4, w.hist, w.digit_base, scores,
                                w.max_enc, w.nan_flag, stream);
  if (rc) return rc;
  return rs_sort_enqueue(times, KEY_NEG_TIME_F32, n, 4, sort_ws(w), perm_out, stream, status, w.nonbinary);
}

extern "C" int mmbs_risk_order(const float* times, int64_t n, int32_t* perm_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(times && perm_out && workspace, "mmbs_risk_order: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_risk_order: n=%lld out of range", (long long)n);
  const CoxWorkspace w = carve_cox(workspace, n);
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_risk_order: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  return cox_sort_common(nullptr, times, nullptr, n, perm_out, w, static_cast<cudaStream_t>(stream));
}

extern "C" int mmbs_cox_forward(const float* scores, const float* times, const float* status,
                                int64_t n, int32_t* perm_out, float* saved_e, float* saved_w,
                                float* loss_out, int32_t* flags_out, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(scores && times && status && perm_out && saved_e && saved_w && loss_out && workspace,
               "mmbs_cox_forward: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_cox_forward: n=%lld out of range [1, 2^30)",
               (long long)n);
  MMBS_REQUIRE((reinterpret_cast<uintptr_t>(perm_out) | reinterpret_cast<uintptr_t>(saved_e) |
                reinterpret_cast<uintptr_t>(saved_w)) % 16 == 0,
               "mmbs_cox_forward: perm_out/saved_e/saved_w must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const CoxWorkspace w = carve_cox(workspace, n);
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_cox_forward: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  if (int rc = cox_sort_common(scores, times, status, n, perm_out, w, stream)) return rc;
  const int64_t tiles = cs_tiles(n);
  ScanState st{w.status_fwd};
  cox_scan_fwd_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(
      perm_out, scores, status, w.max_enc, n, saved_e, saved_w, st, w.counters + 4, w.loss_partial,
      w.nan_flag, w.max_count, w.max_list, w.nonbinary);
  MMBS_LAUNCH_CHECK();
  cox_finalize_kernel<<<1, 256, 0, stream>>>(w.loss_partial, tiles, n, w.nan_flag, loss_out, flags_out);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_cox_backward(const float* scores, const float* status, const int32_t* perm,
                                 const float* saved_e, const float* saved_w, const float* grad_loss,
                                 int64_t n, float* grad_scores, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(scores && status && perm && saved_e && saved_w && grad_loss && grad_scores && workspace,
               "mmbs_cox_backward: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_cox_backward: n=%lld out of range", (long long)n);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const CoxWorkspace w = carve_cox(workspace, n);  // must be the forward's workspace (max list)
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_cox_backward: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  const int64_t tiles = cs_tiles(n);
  MMBS_CUDA_TRY(cudaMemsetAsync(w.status_bwd, 0, size_t(tiles) * sizeof(unsigned long long), stream));
  MMBS_CUDA_TRY(cudaMemsetAsync(w.counters + 5, 0, sizeof(uint32_t), stream));
  ScanState st{w.status_bwd};
  cox_scan_bwd_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(
      perm, status, saved_e, saved_w, grad_loss, n, tiles * CS_TILE, grad_scores, st, w.counters + 5,
      w.gsum_partial, w.nonbinary);
  MMBS_LAUNCH_CHECK();
  cox_maxfix_list_kernel<<<1, 256, 0, stream>>>(w.gsum_partial, tiles, w.max_count, w.max_list,
                                               w.gsum_total, grad_scores);
  MMBS_LAUNCH_CHECK();
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256 * 8), int64_t(sm_count()) * 4)));
  cox_maxfix_full_kernel<<<grid, 256, 0, stream>>>(scores, w.max_enc, w.max_count, w.gsum_total, n,
                                                  grad_scores);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
