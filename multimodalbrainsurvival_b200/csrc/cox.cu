// Cox negative log partial likelihood, forward and backward, for sm_100a.
//
// Replaces cox_loss() of the reference (/root/reference/1_HistoPathology/models.py:90-111
// and its three textual copies):
//   sort(-times) -> gather -> subtract max -> exp -> cumsum -> log(.+1e-5) -> mask -> mean.
//
// n <= 2048 (every training config): one fused block forward, one backward (cox_small_*).
// 2048 < n <= FS_MAX_N: the bucketed pipeline of cox_sort.cu (3 kernels forward, 1 backward).
// Larger n, or inputs the bucketed pipeline gave up on (device flag `fallback`, no host sync) - the kernels below:
// Forward  = risk-set sort -> reduce-then-scan over the sorted order.
//            histogram -> 4 Onesweep radix passes (payload = index | event bit) ->
//            cox_gather_kernel (s~ through the permutation).
//               cox_tilesum_kernel : tile sums of exp(s~)
//               cox_tile_scan_kernel: exclusive scan of the 2048-element tile sums (fp64)
//               cox_loss_kernel    : C = cumsum, log, mask, loss partials; saves w = status/(C+eps)
//                                    and the tile sums of w for the backward pass
// Backward = cox_tile_scan_kernel (suffix) -> cox_grad_kernel (suffix sums of w, gradient
//            scatter) -> the gradient through max(scores).
// Every pass streams with 128-bit accesses and has no inter-block dependency (an earlier
// chained-scan version with decoupled look-back spent >40 % of its time waiting on predecessor
// tiles: profiles/r01_ncu_stalls_cox_scan_fwd.txt); fp32 inside a tile, fp64 across tiles.
#include <algorithm>
#include <cstdlib>

#include "cox_sort.cuh"

namespace mmbs {

constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS = 8;
constexpr int CS_TILE = CS_THREADS * CS_ITEMS;  // 2048 sorted positions per tile
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int COX_MAX_LIST = 1024;
constexpr float COX_EPS = 1e-5f;

static inline int64_t cs_tiles(int64_t n) { return n > 0 ? (n + CS_TILE - 1) / CS_TILE : 1; }

__device__ __forceinline__ void load8_i32(const int32_t* p, int64_t base, int64_t n, int32_t (&v)[8]) {
  if (base + 8 <= n) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p + base));
    const int4 b = __ldg(reinterpret_cast<const int4*>(p + base) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j < n) ? __ldg(p + base + j) : 0;
  }
}
__device__ __forceinline__ void load8_f32(const float* p, int64_t base, int64_t n, float (&v)[8]) {
  if (base + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + base));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + base) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j < n) ? __ldg(p + base + j) : 0.f;
  }
}
__device__ __forceinline__ void store8_f32(float* p, int64_t base, int64_t n, const float (&v)[8]) {
  if (base + 8 <= n) {
    float4* q = reinterpret_cast<float4*>(p + base);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (base + j < n) p[base + j] = v[j];
  }
}
// sum over the block, returned to thread 0 (other threads get 0); s_red has CS_WARPS slots
__device__ __forceinline__ double block_sum_to_t0(double v, double* s_red, int lane, int warp) {
  v = warp_sum(v);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int w = 0; w < CS_WARPS; ++w) t += s_red[w];
  }
  return t;
}

// ------------------------------------------------------------------ forward, pass A
// perm words: bits 0-30 original index, bit 31 = event indicator (binary status).
__global__ void __launch_bounds__(CS_THREADS) cox_gather_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ scores,
    const uint32_t* __restrict__ max_enc, int64_t n, float* __restrict__ saved_s,
    int32_t* max_count, int32_t* __restrict__ max_list, const int32_t* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;   // the two-pass sort already wrote s~ (cox_sort.cu)
  const int tid = threadIdx.x;
  const int64_t base = int64_t(blockIdx.x) * CS_TILE + int64_t(tid) * CS_ITEMS;
  const float smax = float_order_dec(*max_enc);
  const uint64_t pol = make_evict_last_policy();
  int32_t p[CS_ITEMS];
  load8_i32(perm, base, n, p);
  float st[CS_ITEMS];
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {  // the random 4-byte gather (scores kept L2-resident: evict_last)
    const bool valid = base + j < n;
    p[j] &= 0x7fffffff;
    st[j] = valid ? ld_f32_hint(scores + p[j], pol) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const bool valid = base + j < n;
    st[j] -= smax;                                   // s~ (models.py:102)
    if (valid && st[j] == 0.f) {                     // an argmax position (for backward)
      const int pos = atomicAdd(max_count, 1);
      if (pos < COX_MAX_LIST) max_list[pos] = p[j];
    }
  }
  store8_f32(saved_s, base, n, st);
}

// tile sums of exp(s~) over the sorted order (models.py:103), one streaming read of s~
__global__ void __launch_bounds__(CS_THREADS) cox_tilesum_kernel(const float* __restrict__ saved_s, int64_t n,
                                                                 double* __restrict__ tile_sum,
                                                                 const int32_t* __restrict__ enable) {
  __shared__ double s_red[CS_WARPS];
  if (enable != nullptr && *enable == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = int64_t(blockIdx.x) * CS_TILE + int64_t(tid) * CS_ITEMS;
  float st[CS_ITEMS];
  load8_f32(saved_s, base, n, st);
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) run += (base + j < n) ? expf(st[j]) : 0.f;
  const double t = block_sum_to_t0(double(run), s_red, lane, warp);
  if (tid == 0) tile_sum[blockIdx.x] = t;
}

// ------------------------------------------------------------------ tile-sum scan (one block)
// out[i] = sum_{t<i} in[t]  (reverse: sum_{t>i} in[t]); fp64; in == out allowed.
__global__ void __launch_bounds__(1024) cox_tile_scan_kernel(const double* in, double* out, int64_t tiles,
                                                             int reverse, const int32_t* __restrict__ enable) {
  __shared__ double s_w[32];
  __shared__ double s_carry;
  if (enable != nullptr && *enable == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0.0;
  __syncthreads();
  for (int64_t b = 0; b < tiles; b += 1024) {
    const int64_t i = b + tid;
    const int64_t src = reverse ? (tiles - 1 - i) : i;
    const double c = (i < tiles) ? in[src] : 0.0;
    double incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    double off = s_carry;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    __syncthreads();               // every thread has read s_carry / s_w before they change
    if (i < tiles) out[src] = off + incl - c;
    if (tid == 1023) s_carry = off + incl;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ forward, pass B
__global__ void __launch_bounds__(CS_THREADS) cox_loss_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ status,
    const float* __restrict__ saved_s, const double* __restrict__ tile_excl, int64_t n,
    float* __restrict__ saved_w, double* __restrict__ loss_partial, double* __restrict__ tile_wsum,
    int32_t* nan_flag, const int32_t* __restrict__ nonbinary, const int32_t* __restrict__ enable) {
  __shared__ double s_warp_tot[CS_WARPS];
  __shared__ double s_red[CS_WARPS];
  if (enable != nullptr && *enable == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = int64_t(blockIdx.x) * CS_TILE + int64_t(tid) * CS_ITEMS;
  const bool gather_status = (*nonbinary != 0);
  int32_t p[CS_ITEMS];
  float st[CS_ITEMS], dl[CS_ITEMS], c[CS_ITEMS];
  load8_i32(perm, base, n, p);
  load8_f32(saved_s, base, n, st);
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const bool valid = base + j < n;
    const uint32_t word = uint32_t(p[j]);
    dl[j] = (valid && (word >> 31)) ? 1.f : 0.f;
    if (gather_status) dl[j] = valid ? __ldg(status + (word & 0x7fffffffu)) : 0.f;
    run += valid ? expf(st[j]) : 0.f;
    c[j] = run;
  }
  // exclusive prefix of the per-thread totals inside the tile, carried in fp64
  const double tsum = double(run);
  double winc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  if (lane == 31) s_warp_tot[warp] = winc;
  __syncthreads();
  double off = tile_excl[blockIdx.x] + (winc - tsum);
#pragma unroll
  for (int w = 0; w < CS_WARPS; ++w)
    if (w < warp) off += s_warp_tot[w];

  float wv[CS_ITEMS];
  float lsum = 0.f, wsum = 0.f;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    const bool valid = base + j < n;
    const float cj = float(off + double(c[j]));        // cumsum (models.py:104)
    const float den = cj + COX_EPS;
    const float term = -(st[j] - logf(den)) * dl[j];   // models.py:104-105
    wv[j] = valid ? dl[j] / den : 0.f;
    if (valid) {
      lsum += term;
      wsum += wv[j];
      bad |= (term != term);
    }
  }
  store8_f32(saved_w, base, n, wv);
  const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0 && any_bad) atomicOr(nan_flag, 1);
  const double tl = block_sum_to_t0(double(lsum), s_red, lane, warp);
  if (tid == 0) loss_partial[blockIdx.x] = tl;
  __syncthreads();
  const double tw = block_sum_to_t0(double(wsum), s_red, lane, warp);
  if (tid == 0) tile_wsum[blockIdx.x] = tw;
}

__global__ void __launch_bounds__(256) cox_finalize_kernel(const double* __restrict__ partial,
                                                           int64_t tiles, int64_t n,
                                                           const int32_t* __restrict__ nan_flag,
                                                           float* __restrict__ loss_out,
                                                           int32_t* __restrict__ flags_out,
                                                           const int32_t* __restrict__ enable) {
  __shared__ double s_red[8];
  if (enable != nullptr && *enable == 0) return;
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

// ------------------------------------------------------------------ backward
// W_k = sum_{i>=k} w_i  = (sum over later tiles) + suffix inside the tile;
// g~_k = -(status_k - exp(s~_k) * W_k) * grad_loss / n, scattered to grad[perm[k]].
__global__ void __launch_bounds__(CS_THREADS) cox_grad_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ status,
    const float* __restrict__ saved_s, const float* __restrict__ saved_w,
    const double* __restrict__ tile_suffix, const float* __restrict__ grad_loss, int64_t n,
    float* __restrict__ grad_scores, double* __restrict__ gsum_partial,
    const int32_t* __restrict__ nonbinary, const int32_t* __restrict__ enable) {
  __shared__ double s_warp_tot[CS_WARPS];
  __shared__ double s_red[CS_WARPS];
  if (enable != nullptr && *enable == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = int64_t(blockIdx.x) * CS_TILE + int64_t(tid) * CS_ITEMS;
  const double scale = double(grad_loss[0]) / double(n);
  const bool gather_status = (*nonbinary != 0);
  int32_t p[CS_ITEMS];
  float st[CS_ITEMS], w[CS_ITEMS], suf[CS_ITEMS];
  load8_i32(perm, base, n, p);
  load8_f32(saved_s, base, n, st);
  load8_f32(saved_w, base, n, w);   // zero beyond n
  float run = 0.f;
#pragma unroll
  for (int j = CS_ITEMS - 1; j >= 0; --j) {
    run += w[j];
    suf[j] = run;                   // inclusive suffix inside the thread
  }
  const double tsum = double(run);
  double winc = tsum;               // inclusive suffix over lanes: lane l holds sum of lanes >= l
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(0xffffffffu, winc, o);
    if (lane + o < 32) winc += t;
  }
  if (lane == 0) s_warp_tot[warp] = winc;
  __syncthreads();
  double off = tile_suffix[blockIdx.x] + (winc - tsum);
#pragma unroll
  for (int w2 = 0; w2 < CS_WARPS; ++w2)
    if (w2 > warp) off += s_warp_tot[w2];

  double gs = 0.0;   // summed in fp64: the argmax position receives -sum(g~), a sum of n rounded terms otherwise
#pragma unroll
  for (int j = 0; j < CS_ITEMS; ++j) {
    if (base + j < n) {
      const uint32_t word = uint32_t(p[j]);
      const int32_t idx = int32_t(word & 0x7fffffffu);
      float dl = (word >> 31) ? 1.f : 0.f;
      if (gather_status) dl = __ldg(status + idx);
      const double W = off + double(suf[j]);
      const double g = -(double(dl) - double(expf(st[j])) * W) * scale;
      grad_scores[idx] = float(g);        // un-permute
      gs += g;
    }
  }
  const double t = block_sum_to_t0(gs, s_red, lane, warp);
  if (tid == 0) gsum_partial[blockIdx.x] = t;
}

// Gradient through "- max(scores)": every argmax position receives
// -(sum_k g~_k)/count  (torch's full-reduction max backward splits evenly).
__global__ void __launch_bounds__(256) cox_maxfix_list_kernel(
    const double* __restrict__ gsum_partial, int64_t tiles, const int32_t* __restrict__ max_count_pair,
    const int32_t* __restrict__ max_list_pair, const int32_t* __restrict__ fallback, double* __restrict__ gsum_total,
    float* __restrict__ grad_scores) {
  // [0]: written by the bucketed pipeline (which also applies it, cox_sort.cu), [1]: by cox_gather_kernel (LSD path)
  if (*fallback == 0) return;
  const int32_t* max_count = max_count_pair + 1;
  const int32_t* max_list = max_list_pair + COX_MAX_LIST;
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
    const int32_t* __restrict__ max_count_pair, const int32_t* __restrict__ fallback,
    const double* __restrict__ gsum_total, int64_t n, float* __restrict__ grad_scores) {
  const int cnt = max_count_pair[(*fallback != 0) ? 1 : 0];
  if (cnt <= COX_MAX_LIST) return;  // handled by the list kernel
  const float smax = float_order_dec(*max_enc);
  const float fix = float(gsum_total[0] / double(cnt));
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256)
    if (scores[i] - smax == 0.f) grad_scores[i] -= fix;
}


// ------------------------------------------------------------------ small risk sets (n <= 2048)
// The training configs call the loss on 128..1024 samples per step: one block does everything
// (bitonic sort of (key,index) words in shared memory - stable because the index breaks ties -
// then the scans), 1 launch forward + 1 launch backward instead of 14.
constexpr int SM_MAX = 2048;
constexpr int SM_THREADS = 1024;

// inclusive scan of one double per thread over the block (1024 threads)
__device__ __forceinline__ double block_scan_incl(double v, double* s_w /*[32]*/, int lane, int warp) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_w[warp] = v;
  __syncthreads();
  double off = 0.0;
  for (int w = 0; w < warp; ++w) off += s_w[w];
  __syncthreads();
  return v + off;
}

__global__ void __launch_bounds__(SM_THREADS) cox_small_fwd_kernel(
    const float* __restrict__ scores, const float* __restrict__ times, const float* __restrict__ status,
    int n, int32_t* __restrict__ perm_out, float* __restrict__ saved_s, float* __restrict__ saved_w,
    float* __restrict__ loss_out, int32_t* __restrict__ flags_out) {
  __shared__ unsigned long long s_kv[SM_MAX];
  __shared__ double s_w[32];
  __shared__ float s_maxw[32];
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int P = 32;
  while (P < n) P <<= 1;
  if (tid == 0) s_flag = 0;
  float vmax = -INFINITY;
  bool has_nan = false;
  for (int i = tid; i < P; i += SM_THREADS) {
    if (i < n) {
      s_kv[i] = (static_cast<unsigned long long>(time_key(times[i])) << 32) | unsigned(i);
      const float sc = scores[i];
      has_nan |= (sc != sc);
      vmax = fmaxf(vmax, sc);
    } else {
      s_kv[i] = ~0ull;
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0) s_maxw[warp] = vmax;
  __syncthreads();
  float smax = s_maxw[0];
  for (int w = 1; w < 32; ++w) smax = fmaxf(smax, s_maxw[w]);
  // bitonic sort, ascending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += SM_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s_kv[i], b = s_kv[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_kv[i] = b; s_kv[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // thread t owns sorted positions 2t, 2t+1
  const int k0 = 2 * tid;
  float st[2] = {0.f, 0.f}, dl[2] = {0.f, 0.f}, e[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = k0 + j;
    if (k < n) {
      const int idx = int(unsigned(s_kv[k] & 0xffffffffull));
      st[j] = scores[idx] - smax;
      dl[j] = status[idx];
      e[j] = expf(st[j]);
      perm_out[k] = int32_t(unsigned(idx) | (dl[j] != 0.f ? 0x80000000u : 0u));
      saved_s[k] = st[j];
    }
  }
  const double incl = block_scan_incl(double(e[0]) + double(e[1]), s_w, lane, warp);
  const double excl = incl - (double(e[0]) + double(e[1]));
  float lsum = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = k0 + j;
    if (k < n) {
      const float cj = float(excl + double(e[0]) + (j ? double(e[1]) : 0.0));
      const float den = cj + COX_EPS;
      const float term = -(st[j] - logf(den)) * dl[j];
      has_nan |= (term != term);
      lsum += term;
      saved_w[k] = dl[j] / den;
    }
  }
  if (has_nan) s_flag = 1;
  const double total = block_scan_incl(double(lsum), s_w, lane, warp);
  if (tid == SM_THREADS - 1) {
    const int f = s_flag;
    loss_out[0] = f ? __int_as_float(0x7fc00000) : float(total / double(n));
    if (flags_out) flags_out[0] = f;
  }
}

__global__ void __launch_bounds__(SM_THREADS) cox_small_bwd_kernel(
    const float* __restrict__ status, const int32_t* __restrict__ perm, const float* __restrict__ saved_s,
    const float* __restrict__ saved_w, const float* __restrict__ grad_loss, int n,
    float* __restrict__ grad_scores) {
  __shared__ double s_w[32];
  __shared__ double s_tot[2];
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cnt = 0;
  const float scale = grad_loss[0] / float(n);
  const int k0 = 2 * tid;
  float st[2] = {0.f, 0.f}, w[2] = {0.f, 0.f}, dl[2] = {0.f, 0.f};
  int idx[2] = {0, 0};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = k0 + j;
    if (k < n) {
      st[j] = saved_s[k];
      w[j] = saved_w[k];
      idx[j] = int(unsigned(perm[k]) & 0x7fffffffu);
      dl[j] = status[idx[j]];
    }
  }
  const double wsum2 = double(w[0]) + double(w[1]);
  const double incl = block_scan_incl(wsum2, s_w, lane, warp);
  if (tid == SM_THREADS - 1) s_tot[0] = incl;
  __syncthreads();
  const double T = s_tot[0];
  // W_k = sum_{i>=k} w_i = T - sum_{i<k} w_i
  const double before0 = incl - wsum2;
  float g[2] = {0.f, 0.f};
  float gs = 0.f;
  int nmax = 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = k0 + j;
    if (k < n) {
      const float W = float(T - (before0 + (j ? double(w[0]) : 0.0)));
      g[j] = -(dl[j] - expf(st[j]) * W) * scale;
      gs += g[j];
      nmax += (st[j] == 0.f);
    }
  }
  if (nmax) atomicAdd(&s_cnt, nmax);
  const double gincl = block_scan_incl(double(gs), s_w, lane, warp);
  if (tid == SM_THREADS - 1) s_tot[1] = gincl;
  __syncthreads();
  const float fix = float(s_tot[1] / double(s_cnt));   // gradient through max(scores)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = k0 + j;
    if (k < n) grad_scores[idx[j]] = g[j] - (st[j] == 0.f ? fix : 0.f);
  }
}

// ------------------------------------------------------------------ workspace
struct CoxWorkspace {
  // zeroed at the start of forward (contiguous, ~50 KB)
  uint32_t* hist;          // [4][256]
  uint32_t* counters;      // [8]: 0-3 radix passes
  uint32_t* max_enc;       // [1]
  int32_t* nan_flag;       // [1]
  int32_t* max_count;      // [2] argmax positions found by {bucketed pipeline, cox_gather_kernel}
  int32_t* nonbinary;      // [1] some status value is neither 0 nor 1
  int32_t* fallback;       // [1] != 0: the LSD pipeline (re)does the work (always when n > FS_MAX_N)
  uint32_t* fs_hist12;     // bucketed pipeline (cox_sort.cuh FastSortWs): [FS_BINS]
  uint32_t* fs_kext;       // [2]
  uint32_t* fs_counters;   // [8]
  uint32_t* fs_cursor;     // [FS_MAX_BUCKETS]
  double* fs_agg_val;      // [FS_MAX_BUCKETS]
  size_t zero_bytes;
  // zeroed by the LSD pipeline itself, only when it runs
  uint32_t* lookback;      // [4][rs_tiles][256]
  size_t lookback_words;
  // not zeroed
  uint32_t* digit_base;    // [4][256]
  uint2* fs_lut;           // [FS_BINS]
  FsEdge* fs_edge;         // [1]
  double* fs_exp_prefix;   // [FS_MAX_BUCKETS] each
  double* fs_wsum;
  uint32_t* fs_bucket_base;
  uint32_t* fs_bucket_cnt;
  float* fs_part32;        // [FS_MAX_BUCKETS * FS_PART]
  double* fs_row_loss;     // [FS_MAX_BUCKETS * 16] each
  double* fs_row_w;
  double* fs_row_g;
  int32_t* max_list;       // [2][COX_MAX_LIST]
  double* gsum_total;      // [1]
  double* tile_sum;        // [cs_tiles] sums of exp(s~), scanned in place
  double* tile_wsum;       // [cs_tiles] sums of w           (kept for backward)
  double* tile_suffix;     // [cs_tiles] suffix scan of tile_wsum (backward)
  double* loss_partial;    // [cs_tiles]
  double* gsum_partial;    // [cs_tiles]
  // the two pipelines never hold sort buffers at the same time: the bucket regions alias the LSD ping-pong arrays
  uint32_t* keys_a; uint32_t* keys_b; uint32_t* vals_a; uint32_t* vals_b;
  uint2* fs_pairs;         // [nb * FS_CAP]
  float* fs_sc_part;       // [nb * FS_CAP]
  size_t total_bytes;
};

static CoxWorkspace carve_cox(void* base, int64_t n) {
  const int64_t rt = rs_tiles(n), ct = cs_tiles(n);
  const bool fast = n <= FS_MAX_N;
  Carver c(base);
  CoxWorkspace w;
  w.hist = c.take<uint32_t>(4 * RS_RADIX);
  w.counters = c.take<uint32_t>(8);
  w.max_enc = c.take<uint32_t>(1);
  w.nan_flag = c.take<int32_t>(1);
  w.max_count = c.take<int32_t>(2);
  w.nonbinary = c.take<int32_t>(1);
  w.fallback = c.take<int32_t>(1);
  w.fs_hist12 = c.take<uint32_t>(FS_BINS);
  w.fs_kext = c.take<uint32_t>(2);
  w.fs_counters = c.take<uint32_t>(8);
  w.fs_cursor = c.take<uint32_t>(FS_MAX_BUCKETS);
  w.fs_agg_val = c.take<double>(FS_MAX_BUCKETS);
  w.zero_bytes = align_up(c.off, 256);
  w.lookback_words = size_t(4) * rt * RS_RADIX;
  w.lookback = c.take<uint32_t>(w.lookback_words);
  w.digit_base = c.take<uint32_t>(4 * RS_RADIX);
  w.fs_lut = c.take<uint2>(FS_BINS);
  w.fs_edge = c.take<FsEdge>(1);
  w.fs_exp_prefix = c.take<double>(FS_MAX_BUCKETS);
  w.fs_wsum = c.take<double>(FS_MAX_BUCKETS);
  w.fs_bucket_base = c.take<uint32_t>(FS_MAX_BUCKETS);
  w.fs_bucket_cnt = c.take<uint32_t>(FS_MAX_BUCKETS);
  w.fs_part32 = c.take<float>(size_t(FS_MAX_BUCKETS) * FS_PART);
  w.fs_row_loss = c.take<double>(size_t(FS_MAX_BUCKETS) * 16);
  w.fs_row_w = c.take<double>(size_t(FS_MAX_BUCKETS) * 16);
  w.fs_row_g = c.take<double>(size_t(FS_MAX_BUCKETS) * 16);
  w.max_list = c.take<int32_t>(2 * COX_MAX_LIST);
  w.gsum_total = c.take<double>(1);
  w.tile_sum = c.take<double>(ct);
  w.tile_wsum = c.take<double>(ct);
  w.tile_suffix = c.take<double>(ct);
  w.loss_partial = c.take<double>(ct);
  w.gsum_partial = c.take<double>(ct);
  const size_t lsd_words = size_t(4) * size_t(n);
  const size_t fs_words = fast ? size_t(fs_plan(n).nb) * FS_CAP * 3 : 0;
  uint32_t* sortbuf = c.take<uint32_t>(std::max(lsd_words, fs_words));
  w.keys_a = sortbuf;
  w.keys_b = sortbuf + n;
  w.vals_a = sortbuf + 2 * n;
  w.vals_b = sortbuf + 3 * n;
  w.fs_pairs = reinterpret_cast<uint2*>(sortbuf);
  w.fs_sc_part = reinterpret_cast<float*>(sortbuf) + (fast ? size_t(fs_plan(n).nb) * FS_CAP * 2 : 0);
  w.total_bytes = align_up(c.off, 256);
  return w;
}

static SortWorkspace sort_ws(const CoxWorkspace& w) {
  SortWorkspace s;
  s.keys_a = w.keys_a; s.keys_b = w.keys_b; s.vals_a = w.vals_a; s.vals_b = w.vals_b;
  s.hist = w.hist; s.digit_base = w.digit_base; s.counters = w.counters; s.lookback = w.lookback;
  return s;
}

static FastSortWs fast_ws(const CoxWorkspace& w) {
  FastSortWs f;
  f.hist12 = w.fs_hist12; f.kext = w.fs_kext; f.counters = w.fs_counters; f.cursor = w.fs_cursor;
  f.agg_val = w.fs_agg_val; f.fallback = w.fallback; f.lut = w.fs_lut; f.edge = w.fs_edge;
  f.exp_prefix = w.fs_exp_prefix; f.wsum = w.fs_wsum;
  f.bucket_base = w.fs_bucket_base; f.bucket_cnt = w.fs_bucket_cnt; f.pairs = w.fs_pairs;
  f.sc_part = w.fs_sc_part; f.part32 = w.fs_part32;
  f.row_loss = w.fs_row_loss; f.row_w = w.fs_row_w; f.row_g = w.fs_row_g;
  return f;
}

// the decoupled look-back words of the LSD sort are cleared by the pipeline that uses them (a no-op otherwise)
__global__ void __launch_bounds__(256) cox_zero_words_kernel(uint32_t* __restrict__ p, size_t words,
                                                             const int32_t* __restrict__ enable) {
  if (*enable == 0) return;
  uint4* q = reinterpret_cast<uint4*>(p);   // Carver hands out 256-byte aligned arrays
  const size_t quads = words / 4;
  for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < quads; i += size_t(gridDim.x) * 256)
    q[i] = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x == 0 && threadIdx.x < words % 4) p[quads * 4 + threadIdx.x] = 0u;
}


// ---- device-side fallback (CUDA dynamic parallelism, tail launches): the bucketed pipeline raises `fallback` on the
// device; ONE tiny kernel behind it enqueues the whole LSD pipeline when - and only when - the flag is set, so the common
// path carries neither a host synchronisation nor a dozen parked no-op launches.
struct CoxLsdForward {
  const float* scores; const float* times; const float* status; int64_t n;
  int32_t* perm_out; float* saved_s; float* saved_w; float* loss_out; int32_t* flags_out;
  SortWorkspace sw; size_t lookback_words;
  uint32_t* max_enc; int32_t* nan_flag; int32_t* max_count; int32_t* max_list; int32_t* nonbinary;
  double* tile_sum; double* tile_wsum; double* loss_partial;
  const int32_t* fallback;
  int zero_grid, hist_grid; unsigned sort_grid; int64_t sort_tiles, scan_tiles;
  // the bucketed pipeline's one-block finish (cox_sort.cuh fs_loss_finalize_block)
  const double* fs_row_loss; const double* fs_row_w; double* fs_wsuffix; int fs_nb;
};

// One block behind the bucketed forward pass: finishes it (loss, suffix sums of w) - or, when the pipeline gave the input
// up, enqueues the LSD pipeline instead.
__global__ void __launch_bounds__(FS_FINAL_THREADS, 1) cox_lsd_forward_dispatch_kernel(CoxLsdForward a) {
  if (*a.fallback == 0) {   // (block-uniform)
    fs_loss_finalize_block(a.fs_row_loss, a.fs_row_w, a.fs_nb, a.n, a.fs_wsuffix, a.nan_flag, a.loss_out, a.flags_out);
    return;
  }
  if (threadIdx.x != 0) return;
  const unsigned tiles = unsigned(a.scan_tiles);
  cox_zero_words_kernel<<<a.zero_grid, 256, 0, cudaStreamTailLaunch>>>(a.sw.lookback, a.lookback_words, a.fallback);
  rs_sort_tail_launch(a.times, int(KEY_NEG_TIME_F32), a.n, 4, a.sw, a.perm_out, a.status, a.nonbinary, a.hist_grid,
                      a.sort_grid, a.sort_tiles);
  cox_gather_kernel<<<tiles, CS_THREADS, 0, cudaStreamTailLaunch>>>(a.perm_out, a.scores, a.max_enc, a.n, a.saved_s,
                                                                    a.max_count + 1, a.max_list + COX_MAX_LIST, nullptr);
  cox_tilesum_kernel<<<tiles, CS_THREADS, 0, cudaStreamTailLaunch>>>(a.saved_s, a.n, a.tile_sum, nullptr);
  cox_tile_scan_kernel<<<1, 1024, 0, cudaStreamTailLaunch>>>(a.tile_sum, a.tile_sum, a.scan_tiles, 0, nullptr);
  cox_loss_kernel<<<tiles, CS_THREADS, 0, cudaStreamTailLaunch>>>(a.perm_out, a.status, a.saved_s, a.tile_sum, a.n, a.saved_w,
                                                                  a.loss_partial, a.tile_wsum, a.nan_flag, a.nonbinary,
                                                                  nullptr);
  cox_finalize_kernel<<<1, 256, 0, cudaStreamTailLaunch>>>(a.loss_partial, a.scan_tiles, a.n, a.nan_flag, a.loss_out,
                                                           a.flags_out, nullptr);
}

struct CoxLsdBackward {
  const float* scores; const float* status; const int32_t* perm; const float* saved_s; const float* saved_w;
  const float* grad_loss; int64_t n; float* grad_scores;
  uint32_t* max_enc; int32_t* max_count; int32_t* max_list; int32_t* nonbinary; const int32_t* fallback;
  double* tile_wsum; double* tile_suffix; double* gsum_partial; double* gsum_total;
  int64_t scan_tiles; int full_grid;
  const double* fs_row_g; int fs_nb;   // the bucketed pipeline's one-block finish (fs_backward_finalize_block)
};

__global__ void __launch_bounds__(FS_FINAL_THREADS, 1) cox_lsd_backward_dispatch_kernel(CoxLsdBackward a) {
  const bool fb = *a.fallback != 0;
  if (!fb) fs_backward_finalize_block(a.fs_row_g, a.fs_nb, a.max_count, a.max_list, a.gsum_total, a.grad_scores);
  if (threadIdx.x != 0) return;
  if (fb) {
    const unsigned tiles = unsigned(a.scan_tiles);
    cox_tile_scan_kernel<<<1, 1024, 0, cudaStreamTailLaunch>>>(a.tile_wsum, a.tile_suffix, a.scan_tiles, 1, nullptr);
    cox_grad_kernel<<<tiles, CS_THREADS, 0, cudaStreamTailLaunch>>>(a.perm, a.status, a.saved_s, a.saved_w, a.tile_suffix,
                                                                    a.grad_loss, a.n, a.grad_scores, a.gsum_partial,
                                                                    a.nonbinary, nullptr);
    cox_maxfix_list_kernel<<<1, 256, 0, cudaStreamTailLaunch>>>(a.gsum_partial, a.scan_tiles, a.max_count, a.max_list,
                                                                a.fallback, a.gsum_total, a.grad_scores);
  }
  // more argmax positions than the list holds (e.g. all scores equal): a pass over `scores`, either pipeline
  if (a.max_count[fb ? 1 : 0] > COX_MAX_LIST)
    cox_maxfix_full_kernel<<<a.full_grid, 256, 0, cudaStreamTailLaunch>>>(a.scores, a.max_enc, a.max_count, a.fallback,
                                                                          a.gsum_total, a.n, a.grad_scores);
}

}  // namespace mmbs

using namespace mmbs;

extern "C" size_t mmbs_cox_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return carve_cox(nullptr, n).total_bytes;
}

static bool use_bucketed(int64_t n) {
  static const bool disabled = []() {
    const char* e = getenv("MMBS_COX_LSD");   // experiments / tests: force the LSD-sort pipeline
    return e && e[0] == '1';
  }();
  return !disabled && n <= FS_MAX_N;
}

static int lsd_zero_grid(const CoxWorkspace& w) {
  return int(std::max<size_t>(1, std::min<size_t>(w.lookback_words / 4 / 256 + 1, size_t(sm_count()) * 4)));
}

// LSD pipeline up to the permutation (+ max(scores) / NaN flag when `scores` is given), enqueued from the host.
static int cox_lsd_sort(const float* scores, const float* times, const float* status, int64_t n, int32_t* perm_out,
                        const CoxWorkspace& w, cudaStream_t stream) {
  cox_zero_words_kernel<<<lsd_zero_grid(w), 256, 0, stream>>>(w.lookback, w.lookback_words, w.fallback);
  MMBS_LAUNCH_CHECK();
  int rc = rs_histogram_enqueue(times, KEY_NEG_TIME_F32, n, 4, w.hist, w.digit_base, scores, w.max_enc, w.nan_flag,
                                stream, nullptr);
  if (rc) return rc;
  return rs_sort_enqueue(times, KEY_NEG_TIME_F32, n, 4, sort_ws(w), perm_out, stream, status, w.nonbinary, nullptr);
}

// mmbs_risk_order on the bucketed pipeline: the permutation only
struct CoxLsdOrder {
  const float* times; int64_t n; int32_t* perm_out; SortWorkspace sw; size_t lookback_words; int32_t* nonbinary;
  const int32_t* fallback; int zero_grid, hist_grid; unsigned sort_grid; int64_t sort_tiles;
};
namespace mmbs {
__global__ void cox_lsd_order_dispatch_kernel(CoxLsdOrder a) {
  if (*a.fallback == 0) return;
  cox_zero_words_kernel<<<a.zero_grid, 256, 0, cudaStreamTailLaunch>>>(a.sw.lookback, a.lookback_words, a.fallback);
  rs_sort_tail_launch(a.times, int(KEY_NEG_TIME_F32), a.n, 4, a.sw, a.perm_out, nullptr, a.nonbinary, a.hist_grid,
                      a.sort_grid, a.sort_tiles);
}
}  // namespace mmbs

extern "C" int mmbs_cox_debug_state(const void* workspace, size_t workspace_bytes, int64_t n, int32_t* out_host) {
  MMBS_REQUIRE(workspace && out_host && n >= 1, "mmbs_cox_debug_state: bad argument");
  const CoxWorkspace w = carve_cox(const_cast<void*>(workspace), n);
  MMBS_REQUIRE(workspace_bytes >= w.total_bytes, "mmbs_cox_debug_state: workspace too small");
  out_host[0] = -1; out_host[1] = 0; out_host[2] = 0; out_host[3] = FS_CAP;
  if (n <= SM_MAX || n > FS_MAX_N) return MMBS_OK;
  MMBS_CUDA_TRY(cudaDeviceSynchronize());
  MMBS_CUDA_TRY(cudaMemcpy(out_host, w.fallback, sizeof(int32_t), cudaMemcpyDeviceToHost));
  static uint32_t host_cursor[FS_MAX_BUCKETS];
  const int nb = fs_plan(n).nb;
  MMBS_CUDA_TRY(cudaMemcpy(host_cursor, w.fs_cursor, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost));
  uint32_t big = 0;
  for (int i = 0; i < nb; ++i) big = std::max(big, host_cursor[i]);
  out_host[1] = int32_t(big);
  out_host[2] = nb;
  return MMBS_OK;
}

extern "C" int mmbs_risk_order(const float* times, int64_t n, int32_t* perm_out, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(times && perm_out && workspace, "mmbs_risk_order: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_risk_order: n=%lld out of range", (long long)n);
  const CoxWorkspace w = carve_cox(workspace, n);
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_risk_order: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMBS_CUDA_TRY(cudaMemsetAsync(w.hist, 0, w.zero_bytes, stream));
  if (use_bucketed(n)) {
    if (int rc = fs_forward_enqueue(times, nullptr, nullptr, n, fast_ws(w), w.max_enc, w.nan_flag, w.nonbinary, perm_out,
                                    nullptr, w.max_count, w.max_list, nullptr, nullptr, stream))
      return rc;
    if (int rc = rs_configure()) return rc;
    CoxLsdOrder a;
    a.times = times; a.n = n; a.perm_out = perm_out; a.sw = sort_ws(w); a.lookback_words = w.lookback_words;
    a.nonbinary = w.nonbinary; a.fallback = w.fallback; a.zero_grid = lsd_zero_grid(w); a.hist_grid = rs_hist_grid(n);
    a.sort_grid = rs_sort_grid(n); a.sort_tiles = rs_tiles(n);
    cox_lsd_order_dispatch_kernel<<<1, 1, 0, stream>>>(a);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  MMBS_CUDA_TRY(cudaMemsetAsync(w.fallback, 0xff, sizeof(int32_t), stream));
  return cox_lsd_sort(nullptr, times, nullptr, n, perm_out, w, stream);
}

extern "C" int mmbs_cox_forward(const float* scores, const float* times, const float* status,
                                int64_t n, int32_t* perm_out, float* saved_s, float* saved_w,
                                float* loss_out, int32_t* flags_out, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(scores && times && status && perm_out && saved_s && saved_w && loss_out,
               "mmbs_cox_forward: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_cox_forward: n=%lld out of range [1, 2^30)",
               (long long)n);
  MMBS_REQUIRE((reinterpret_cast<uintptr_t>(perm_out) | reinterpret_cast<uintptr_t>(saved_s) |
                reinterpret_cast<uintptr_t>(saved_w)) % 16 == 0,
               "mmbs_cox_forward: perm_out/saved_s/saved_w must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= SM_MAX) {  // one fused block; needs no workspace
    cox_small_fwd_kernel<<<1, SM_THREADS, 0, stream>>>(scores, times, status, int(n), perm_out, saved_s,
                                                      saved_w, loss_out, flags_out);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  MMBS_REQUIRE(workspace != nullptr, "mmbs_cox_forward: null workspace");
  const CoxWorkspace w = carve_cox(workspace, n);
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_cox_forward: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  MMBS_CUDA_TRY(cudaMemsetAsync(w.hist, 0, w.zero_bytes, stream));
  const int64_t tiles = cs_tiles(n);
  if (use_bucketed(n)) {
    // bucketed pipeline: saved_w stays unwritten (its backward recomputes w from s~)
    if (int rc = fs_forward_enqueue(times, status, scores, n, fast_ws(w), w.max_enc, w.nan_flag, w.nonbinary, perm_out,
                                    saved_s, w.max_count, w.max_list, loss_out, flags_out, stream))
      return rc;
    // ... unless it raised `fallback`: then this kernel enqueues the LSD pipeline from the device
    if (int rc = rs_configure()) return rc;
    CoxLsdForward a;
    a.scores = scores; a.times = times; a.status = status; a.n = n;
    a.perm_out = perm_out; a.saved_s = saved_s; a.saved_w = saved_w; a.loss_out = loss_out; a.flags_out = flags_out;
    a.sw = sort_ws(w); a.lookback_words = w.lookback_words;
    a.max_enc = w.max_enc; a.nan_flag = w.nan_flag; a.max_count = w.max_count; a.max_list = w.max_list;
    a.nonbinary = w.nonbinary; a.tile_sum = w.tile_sum; a.tile_wsum = w.tile_wsum; a.loss_partial = w.loss_partial;
    a.fallback = w.fallback;
    a.zero_grid = lsd_zero_grid(w); a.hist_grid = rs_hist_grid(n); a.sort_grid = rs_sort_grid(n);
    a.sort_tiles = rs_tiles(n); a.scan_tiles = tiles;
    a.fs_row_loss = w.fs_row_loss; a.fs_row_w = w.fs_row_w; a.fs_wsuffix = w.fs_wsum; a.fs_nb = fs_plan(n).nb;
    cox_lsd_forward_dispatch_kernel<<<1, FS_FINAL_THREADS, 0, stream>>>(a);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  // LSD pipeline, enqueued from the host
  MMBS_CUDA_TRY(cudaMemsetAsync(w.fallback, 0xff, sizeof(int32_t), stream));
  if (int rc = cox_lsd_sort(scores, times, status, n, perm_out, w, stream)) return rc;
  cox_gather_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(perm_out, scores, w.max_enc, n, saved_s,
                                                               w.max_count + 1, w.max_list + COX_MAX_LIST, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_tilesum_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(saved_s, n, w.tile_sum, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_tile_scan_kernel<<<1, 1024, 0, stream>>>(w.tile_sum, w.tile_sum, tiles, 0, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_loss_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(perm_out, status, saved_s, w.tile_sum, n,
                                                             saved_w, w.loss_partial, w.tile_wsum,
                                                             w.nan_flag, w.nonbinary, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_finalize_kernel<<<1, 256, 0, stream>>>(w.loss_partial, tiles, n, w.nan_flag, loss_out, flags_out, nullptr);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_cox_backward(const float* scores, const float* status, const int32_t* perm,
                                 const float* saved_s, const float* saved_w, const float* grad_loss,
                                 int64_t n, float* grad_scores, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(scores && status && perm && saved_s && saved_w && grad_loss && grad_scores,
               "mmbs_cox_backward: null pointer");
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "mmbs_cox_backward: n=%lld out of range", (long long)n);
  MMBS_REQUIRE((reinterpret_cast<uintptr_t>(perm) | reinterpret_cast<uintptr_t>(saved_s) |
                reinterpret_cast<uintptr_t>(saved_w)) % 16 == 0,
               "mmbs_cox_backward: perm/saved_s/saved_w must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= SM_MAX) {
    cox_small_bwd_kernel<<<1, SM_THREADS, 0, stream>>>(status, perm, saved_s, saved_w, grad_loss, int(n),
                                                      grad_scores);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  MMBS_REQUIRE(workspace != nullptr, "mmbs_cox_backward: null workspace");
  const CoxWorkspace w = carve_cox(workspace, n);  // must be the forward's workspace (bucket / tile sums, max list)
  if (workspace_bytes < w.total_bytes) {
    set_error("mmbs_cox_backward: workspace %zu < %zu bytes", workspace_bytes, w.total_bytes);
    return MMBS_ERR_WORKSPACE;
  }
  const int64_t tiles = cs_tiles(n);
  const int full_grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256 * 8), int64_t(sm_count()) * 4)));
  if (use_bucketed(n)) {
    if (int rc = fs_backward_enqueue(status, perm, saved_s, grad_loss, n, fast_ws(w), w.nonbinary, w.max_count,
                                     w.max_list, w.gsum_total, grad_scores, stream))
      return rc;
    CoxLsdBackward a;
    a.scores = scores; a.status = status; a.perm = perm; a.saved_s = saved_s; a.saved_w = saved_w;
    a.grad_loss = grad_loss; a.n = n; a.grad_scores = grad_scores;
    a.max_enc = w.max_enc; a.max_count = w.max_count; a.max_list = w.max_list; a.nonbinary = w.nonbinary;
    a.fallback = w.fallback; a.tile_wsum = w.tile_wsum; a.tile_suffix = w.tile_suffix;
    a.gsum_partial = w.gsum_partial; a.gsum_total = w.gsum_total; a.scan_tiles = tiles; a.full_grid = full_grid;
    a.fs_row_g = w.fs_row_g; a.fs_nb = fs_plan(n).nb;
    cox_lsd_backward_dispatch_kernel<<<1, FS_FINAL_THREADS, 0, stream>>>(a);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  cox_tile_scan_kernel<<<1, 1024, 0, stream>>>(w.tile_wsum, w.tile_suffix, tiles, 1, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_grad_kernel<<<unsigned(tiles), CS_THREADS, 0, stream>>>(perm, status, saved_s, saved_w, w.tile_suffix,
                                                             grad_loss, n, grad_scores, w.gsum_partial,
                                                             w.nonbinary, nullptr);
  MMBS_LAUNCH_CHECK();
  cox_maxfix_list_kernel<<<1, 256, 0, stream>>>(w.gsum_partial, tiles, w.max_count, w.max_list, w.fallback,
                                               w.gsum_total, grad_scores);
  MMBS_LAUNCH_CHECK();
  cox_maxfix_full_kernel<<<full_grid, 256, 0, stream>>>(scores, w.max_enc, w.max_count, w.fallback, w.gsum_total, n,
                                                       grad_scores);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
