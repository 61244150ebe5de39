// Onesweep-style stable radix sort for sm_100a.  See radix_sort.cuh.
#include <algorithm>

#include "radix_sort.cuh"

namespace mmbs {

// ---------------------------------------------------------------- histogram
// One pass over the source: per-digit counts for every radix pass, 4 keys per thread
// (128-bit loads).  Shared-memory atomics, with two guards against skewed bytes
// (survival times share a few exponent values; integer-valued months share zero low
// bytes): a thread whose 4 digits agree adds 4 at once, and a warp whose 128 digits
// agree adds 128 from one lane.  Optionally fused: max over `scores` (order-encoded u32
// atomicMax, loaded with an L2 evict_last policy because the forward scan gathers from
// `scores` again after the sort has streamed through L2) and a NaN flag.
__global__ void __launch_bounds__(256) rs_histogram_kernel(
    const void* __restrict__ src, int kind, int64_t n, int num_passes, uint32_t* __restrict__ hist,
    const float* __restrict__ scores, uint32_t* __restrict__ max_enc, int32_t* __restrict__ nan_flag,
    const int32_t* __restrict__ enable) {
  __shared__ uint32_t s_hist[4][RS_RADIX];
  __shared__ uint32_t s_max[8];
  if (enable != nullptr && *enable == 0) return;   // fallback sort not needed (see cox_sort.cu)
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  for (int i = tid; i < 4 * RS_RADIX; i += 256) (&s_hist[0][0])[i] = 0;
  __syncthreads();

  const uint64_t pol = make_evict_last_policy();
  const bool vec_src = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const bool vec_sc = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
  float vmax = -INFINITY;
  bool has_nan = false;
  // warp-uniform trip count: every lane runs the same number of iterations
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {
    const int64_t base = blk * 1024 + int64_t(tid) * 4;
    const int cnt = int(max((long long)0, min((long long)4, (long long)(n - base))));
    uint32_t k[4] = {0u, 0u, 0u, 0u};
    if (cnt == 4 && vec_src) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint32_t*>(src) + base));
      k[0] = raw.x; k[1] = raw.y; k[2] = raw.z; k[3] = raw.w;
      if (kind == KEY_NEG_TIME_F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = time_key(__uint_as_float(k[i]));
      }
    } else {
      for (int i = 0; i < cnt; ++i) k[i] = rs_load_key(src, kind, base + i);
    }
    if (scores != nullptr) {
      float s4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (cnt == 4 && vec_sc) {
        const float4 v = ld_f32x4_hint(scores + base, pol);
        s4[0] = v.x; s4[1] = v.y; s4[2] = v.z; s4[3] = v.w;
      } else {
        for (int i = 0; i < cnt; ++i) s4[i] = ld_f32_hint(scores + base + i, pol);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        has_nan |= (s4[i] != s4[i]);
        vmax = fmaxf(vmax, s4[i]);
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (p < num_passes) {
        const uint32_t d0 = (k[0] >> (8 * p)) & 0xffu, d1 = (k[1] >> (8 * p)) & 0xffu;
        const uint32_t d2 = (k[2] >> (8 * p)) & 0xffu, d3 = (k[3] >> (8 * p)) & 0xffu;
        const bool same4 = (cnt == 4) && d0 == d1 && d1 == d2 && d2 == d3;
        const uint32_t dl = __shfl_sync(0xffffffffu, d0, 0);
        if (__all_sync(0xffffffffu, same4 && d0 == dl)) {
          if (lane == 0) atomicAdd(&s_hist[p][dl], 128u);
        } else if (same4) {
          atomicAdd(&s_hist[p][d0], 4u);
        } else {
          if (cnt > 0) atomicAdd(&s_hist[p][d0], 1u);
          if (cnt > 1) atomicAdd(&s_hist[p][d1], 1u);
          if (cnt > 2) atomicAdd(&s_hist[p][d2], 1u);
          if (cnt > 3) atomicAdd(&s_hist[p][d3], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < num_passes * RS_RADIX; i += 256) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(hist + i, c);
  }
  if (scores != nullptr) {
    vmax = warp_max(vmax);
    if (lane == 0) s_max[tid >> 5] = float_order_enc(vmax);
    const unsigned any_nan = __ballot_sync(0xffffffffu, has_nan);
    if (lane == 0 && any_nan) atomicOr(nan_flag, 1);
    __syncthreads();
    if (tid == 0) {
      uint32_t m = s_max[0];
      for (int w = 1; w < 8; ++w) m = max(m, s_max[w]);
      atomicMax(max_enc, m);
    }
  }
}

// Exclusive scan of each pass's 256 counters -> first global slot of every digit.
__global__ void __launch_bounds__(RS_RADIX) rs_digit_base_kernel(const uint32_t* __restrict__ hist,
                                                                 uint32_t* __restrict__ digit_base,
                                                                 const int32_t* __restrict__ enable) {
  __shared__ uint32_t s_w[8];
  if (enable != nullptr && *enable == 0) return;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t c = hist[p * RS_RADIX + tid];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t off = 0;
  for (int w = 0; w < warp; ++w) off += s_w[w];
  digit_base[p * RS_RADIX + tid] = off + incl - c;
}

// ---------------------------------------------------------------- one radix pass
__global__ void __launch_bounds__(RS_THREADS, RS_BLOCKS_PER_SM) rs_onesweep_kernel(
    const void* __restrict__ src, int kind, const uint32_t* __restrict__ keys_in,
    const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, int shift, const uint32_t* __restrict__ digit_base,
    uint32_t* lookback, uint32_t* tile_counter, int first, int last,
    const float* __restrict__ status, int32_t* __restrict__ nonbinary_flag, const int32_t* __restrict__ enable) {
  if (enable != nullptr && *enable == 0) return;   // fallback sort not needed (see cox_sort.cu)
  __shared__ uint32_t s_warp_hist[RS_WARPS][RS_RADIX];
  __shared__ uint32_t s_digit_start[RS_RADIX];
  __shared__ uint32_t s_global_base[RS_RADIX];
  __shared__ uint32_t s_wsum[RS_RADIX / 32];
  __shared__ uint32_t s_tile;
  extern __shared__ uint32_t s_dyn[];   // [RS_TILE] keys | [RS_TILE] payloads, in digit order
  uint32_t* s_keys = s_dyn;
  uint32_t* s_vals = s_dyn + RS_TILE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t num_tiles = (n + RS_TILE - 1) / RS_TILE;
  // Persistent blocks; tiles are handed out by an atomic ticket in the order blocks ask for them: a
  // tile only ever waits on tiles whose ticket is already held by a running block, so the look-back
  // cannot deadlock.  The ticket of the NEXT tile is requested one iteration ahead (latency hidden).
  uint32_t next_ticket = 0;
  if (tid == 0) next_ticket = atomicAdd(tile_counter, 1u);
  while (true) {
  if (tid == 0) s_tile = next_ticket;
  for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  if (tile >= num_tiles) break;
  if (RS_PERSISTENT && tid == 0) next_ticket = atomicAdd(tile_counter, 1u);
  const int64_t tile_base = tile * RS_TILE;
  const int n_valid = int(min((long long)RS_TILE, (long long)(n - tile_base)));

  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
  bool nonbinary = false;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int it = warp * (32 * RS_ITEMS) + j * 32 + lane;
    const int64_t g = tile_base + it;
    if (it < n_valid) {
      key[j] = first ? rs_load_key(src, kind, g) : __ldg(keys_in + g);
      // the payload (or, in the first pass, the status it is built from) is not needed before the
      // shared-memory scatter: its load stays in flight during the whole ranking / look-back phase
      val[j] = first ? ((status != nullptr) ? __float_as_uint(__ldg(status + g)) : 0u) : __ldg(vals_in + g);
    } else {
      key[j] = 0xffffffffu;  // padding sorts to the very end of the (last) tile
      val[j] = 0xffffffffu;
    }
  }

  // stable rank of every key among equal digits inside its warp.  The peer mask is built
  // from 8 ballots (one per digit bit): MATCH.ANY issues ~50x slower than VOTE on sm_100.
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const uint32_t d = (key[j] >> shift) & 0xffu;
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const uint32_t bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
    rank[j] = peers;  // parked here until the counter pass below
  }
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const uint32_t d = (key[j] >> shift) & 0xffu;
    const unsigned m = rank[j];
    const int leader = __ffs(m) - 1;
    uint32_t prev = 0;
    if (lane == leader) {
      prev = s_warp_hist[warp][d];
      s_warp_hist[warp][d] = prev + __popc(m);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[j] = prev + __popc(m & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread `tid` < 256 owns digit `tid`: exclusive scan over warps, tile total
  const bool digit_thread = tid < RS_RADIX;
  uint32_t total = 0;
  uint32_t* lb = lookback + tile * RS_RADIX;
  uint32_t incl = 0;
  if (digit_thread) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = s_warp_hist[w][tid];
      s_warp_hist[w][tid] = total;
      total += c;
    }
    st_volatile_u32(lb + tid, (tile == 0 ? RS_FLAG_INCL : RS_FLAG_AGG) | total);
    // tile-local exclusive scan over digits
    incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
  }
  __syncthreads();
  if (digit_thread) {
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_wsum[w];
    s_digit_start[tid] = woff + incl - total;
  }

  __syncthreads();   // s_digit_start complete
  // bring the tile into digit order in shared memory, then stream it out: equal
  // digits leave as contiguous runs (coalesced stores)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const uint32_t d = (key[j] >> shift) & 0xffu;
    const uint32_t pos = s_digit_start[d] + s_warp_hist[warp][d] + rank[j];
    uint32_t v = val[j];
    if (first) {
      // payload = original index; the (binary) event indicator rides in bit 31 so that the scans
      // never gather `status` through the permutation
      const int it = warp * (32 * RS_ITEMS) + j * 32 + lane;
      const float st = __uint_as_float(v);
      v = uint32_t(tile_base + it);
      if (status != nullptr && it < n_valid) {
        if (st != 0.0f) v |= 0x80000000u;
        if (st != 0.0f && st != 1.0f) nonbinary = true;
      }
    }
    s_keys[pos] = key[j];
    s_vals[pos] = v;
  }
  if (nonbinary) atomicOr(nonbinary_flag, 1);
  // (no barrier here: the look-back below ends with one)

  // The tile aggregate was published before the scatter above, so successors are not held up; by
  // now the predecessors have had time to publish theirs and the look-back rarely has to wait.
  // decoupled look-back: sum this digit's counts over all earlier tiles
  // (RS_LOOKBACK predecessors are fetched per round trip: the loads of a round are independent,
  //  so the inclusive-prefix wavefront advances up to RS_LOOKBACK tiles per L2 latency)
  uint32_t excl = 0;
  if (digit_thread && tile > 0) {
    int64_t p = tile - 1;
    bool done = false;
    while (!done) {
      uint32_t v[RS_LOOKBACK];
#pragma unroll
      for (int u = 0; u < RS_LOOKBACK; ++u)
        v[u] = (p - u >= 0) ? ld_volatile_u32(lookback + (p - u) * RS_RADIX + tid) : RS_FLAG_INCL;
#pragma unroll
      for (int u = 0; u < RS_LOOKBACK; ++u) {
        if (!done) {
          while ((v[u] & RS_FLAG_MASK) == 0) {  // predecessor not published yet: back off instead of hammering L2
            __nanosleep(RS_SPIN_NS);
            v[u] = ld_volatile_u32(lookback + (p - u) * RS_RADIX + tid);
          }
          excl += v[u] & RS_VALUE_MASK;
          if (v[u] & RS_FLAG_INCL) done = true;
        }
      }
      p -= RS_LOOKBACK;
    }
    st_volatile_u32(lb + tid, RS_FLAG_INCL | (excl + total));
  }
  if (digit_thread) s_global_base[tid] = digit_base[tid] + excl;
  __syncthreads();

  constexpr int kUnrollOut = RS_UNROLL_OUT;
#pragma unroll kUnrollOut
  for (int j = 0; j < RS_ITEMS; ++j) {   // unrolled: the shared-memory reads of a thread overlap
    const int i = j * RS_THREADS + tid;
    if (i < n_valid) {
      const uint32_t k = s_keys[i];
      const uint32_t d = (k >> shift) & 0xffu;
      const uint32_t dst = s_global_base[d] + (uint32_t(i) - s_digit_start[d]);
      if (!last) keys_out[dst] = k;
      vals_out[dst] = s_vals[i];
    }
  }
  if (!RS_PERSISTENT) break;
  __syncthreads();   // shared memory is reused by the next tile
  }  // persistent tile loop
}

// Device-side twin of rs_histogram_enqueue + rs_sort_enqueue: one thread of a running kernel enqueues the whole sort
// behind its own grid (CUDA dynamic parallelism, tail-launch stream: the grids run in launch order once the launching
// grid has finished and before the next kernel of the host stream starts).  Used by the Cox loss to fall back from the
// bucketed pipeline without a host synchronisation and without parking no-op launches in the stream (cox.cu).
__device__ void rs_sort_tail_launch(const void* src, int kind, int64_t n, int num_passes, SortWorkspace ws,
                                    int32_t* perm_out, const float* status, int32_t* nonbinary_flag, int hist_grid,
                                    unsigned sort_grid, int64_t tiles) {
  rs_histogram_kernel<<<hist_grid, 256, 0, cudaStreamTailLaunch>>>(src, kind, n, num_passes, ws.hist, nullptr, nullptr,
                                                                   nullptr, nullptr);
  rs_digit_base_kernel<<<num_passes, RS_RADIX, 0, cudaStreamTailLaunch>>>(ws.hist, ws.digit_base, nullptr);
  const uint32_t* kin = nullptr;
  const uint32_t* vin = nullptr;
  for (int p = 0; p < num_passes; ++p) {
    const bool first = (p == 0), last = (p == num_passes - 1);
    uint32_t* kout = (p & 1) ? ws.keys_b : ws.keys_a;
    uint32_t* vout = last ? reinterpret_cast<uint32_t*>(perm_out) : ((p & 1) ? ws.vals_b : ws.vals_a);
    rs_onesweep_kernel<<<sort_grid, RS_THREADS, RS_DYN_SMEM, cudaStreamTailLaunch>>>(
        src, kind, kin, vin, kout, vout, n, 8 * p, ws.digit_base + p * RS_RADIX, ws.lookback + int64_t(p) * tiles * RS_RADIX,
        ws.counters + p, first ? 1 : 0, last ? 1 : 0, first ? status : nullptr, nonbinary_flag, nullptr);
    kin = kout;
    vin = vout;
  }
}

int rs_configure() {
  static PerDeviceOnce configured;
  if (configured.first())
    MMBS_CUDA_TRY(cudaFuncSetAttribute(rs_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_DYN_SMEM));
  return MMBS_OK;
}
int rs_hist_grid(int64_t n) {
  const int64_t want = ceil_div(n, 256 * 8);
  return int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(sm_count()) * 8)));
}
unsigned rs_sort_grid(int64_t n) {
  const int64_t tiles = rs_tiles(n);
  return RS_PERSISTENT ? unsigned(std::min<int64_t>(tiles, int64_t(sm_count()) * RS_BLOCKS_PER_SM)) : unsigned(tiles);
}

int rs_histogram_enqueue(const void* src, KeyKind kind, int64_t n, int num_passes, uint32_t* hist,
                         uint32_t* digit_base, const float* scores, uint32_t* max_enc,
                         int32_t* nan_flag, cudaStream_t stream, const int32_t* enable) {
  const int grid = rs_hist_grid(n);
  rs_histogram_kernel<<<grid, 256, 0, stream>>>(src, int(kind), n, num_passes, hist, scores, max_enc,
                                                nan_flag, enable);
  MMBS_LAUNCH_CHECK();
  rs_digit_base_kernel<<<num_passes, RS_RADIX, 0, stream>>>(hist, digit_base, enable);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

int rs_sort_enqueue(const void* src, KeyKind kind, int64_t n, int num_passes,
                    const SortWorkspace& ws, int32_t* perm_out, cudaStream_t stream,
                    const float* status, int32_t* nonbinary_flag, const int32_t* enable) {
  MMBS_REQUIRE(n >= 1 && n <= RS_MAX_N, "radix sort: n=%lld out of range [1, 2^30)", (long long)n);
  MMBS_REQUIRE(num_passes >= 1 && num_passes <= 4, "radix sort: num_passes=%d", num_passes);
  const int64_t tiles = rs_tiles(n);
  if (int rc = rs_configure()) return rc;
  const unsigned grid = rs_sort_grid(n);
  const uint32_t* kin = nullptr;
  const uint32_t* vin = nullptr;
  for (int p = 0; p < num_passes; ++p) {
    const bool first = (p == 0), last = (p == num_passes - 1);
    uint32_t* kout = (p & 1) ? ws.keys_b : ws.keys_a;
    uint32_t* vout = last ? reinterpret_cast<uint32_t*>(perm_out) : ((p & 1) ? ws.vals_b : ws.vals_a);
    rs_onesweep_kernel<<<grid, RS_THREADS, RS_DYN_SMEM, stream>>>(
        src, int(kind), kin, vin, kout, vout, n, 8 * p, ws.digit_base + p * RS_RADIX,
        ws.lookback + int64_t(p) * tiles * RS_RADIX, ws.counters + p, first ? 1 : 0, last ? 1 : 0,
        first ? status : nullptr, nonbinary_flag, enable);
    MMBS_LAUNCH_CHECK();
    kin = kout;
    vin = vout;
  }
  return MMBS_OK;
}

}  // namespace mmbs
