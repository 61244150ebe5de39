// Two-pass risk-set sort for the Cox loss (sm_100a): one stable MSD partition into <= 1024 buckets of ~8-10 K
// samples, then every bucket is sorted inside shared memory by ONE block, which also writes the permutation, the
// event bit and the shifted score s~ = scores[perm] - max (the gather of the forward pass).
//
// Replaces  _, idx = torch.sort(-times); scores[idx]; status[idx]   of cox_loss()
//   /root/reference/1_HistoPathology/models.py:99-101 (and its three textual copies)
// for 2048 < n <= FS_MAX_N; larger risk sets and inputs the bucket map cannot balance run the 4-pass LSD sort
// (radix_sort.cu) - the decision is taken ON THE DEVICE (flag `fallback`), never by a host synchronisation.
//
// Why two passes are enough.  The LSD sort moves every (key, index) pair through HBM/L2 four times and ranks it
// four times.  Here:
//   fs_hist_kernel      one read of `times`: 4096-bin histogram of the top 12 key bits (+ max(scores), NaN flag);
//                       its last block turns the histogram into a piecewise-linear CDF table  lut[bin] = (P, c).
//   bucket_of(key)      = floor(nb * R(key) / n),  R(key) = P[bin] + (low20(key) * c[bin] >> 20)  - monotone in the
//                       key, so buckets are contiguous key ranges, and near-uniform in size for any smooth
//                       distribution of survival times (the float key is linear in t inside a 12-bit bin).
//   fs_count_kernel     exact bucket sizes (one more read of `times`, L2-resident); its last block scans them into
//                       bucket offsets and the work list of the local sort.
//   fs_partition_kernel Onesweep-style stable partition (ballot ranking over the bucket bits, per-warp u16
//                       counters, decoupled look-back over tiles): writes (key, index | event << 31) into the
//                       bucket's exact slot range.  8 B read + 8 B written per sample.
//   fs_local_sort_kernel one block per bucket (<= 16384 samples, 32 per thread in registers): LSD passes of 9 bits
//                       over only the key bits that DIFFER inside the bucket (14-17 bits for 10 M distinct times:
//                       2 passes) entirely in shared memory, then perm / s~ straight to their final positions.
// Stability: the partition keeps the input order inside a bucket and the local passes are stable, so ties keep
// ascending original index - bit-exact with torch.sort(-times, stable=True).
// A bucket larger than one block's capacity is legal when all its keys are equal (heavy ties: nothing to sort, its
// chunks are copied through); otherwise `fallback` is raised and the LSD kernels, which are always enqueued behind
// and exit immediately when the flag is clear, redo the sort.
#include <algorithm>

#include "cox_sort.cuh"

namespace mmbs {

constexpr int FS_HIST_THREADS = 256;
constexpr int P_THREADS = 512;
constexpr int P_ITEMS = 16;
constexpr int P_TILE = P_THREADS * P_ITEMS;   // 8192
constexpr int P_WARPS = P_THREADS / 32;
constexpr int P_DYN_SMEM = 2 * P_TILE * 4;    // staged keys | payloads
constexpr int P_LOOKBACK = 4;
constexpr int L_THREADS = 512;
constexpr int L_ITEMS = FS_CAP / L_THREADS;   // 32
constexpr int L_WARPS = L_THREADS / 32;
constexpr int L_BITS = 9;
constexpr int L_RADIX = 1 << L_BITS;
constexpr int L_DYN_SMEM = 2 * FS_CAP * 4;

__device__ __forceinline__ uint32_t bucket_of(uint32_t key, const uint2* __restrict__ lut, uint32_t mult) {
  const uint2 e = __ldg(lut + (key >> 20));
  const uint32_t r = e.x + uint32_t((uint64_t(key & 0xfffffu) * e.y) >> 20);
  return __umulhi(r, mult);
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// inclusive scan of one u32 per thread over a block of NW warps; s_w has NW slots; returns (inclusive, block total)
template <int NW>
__device__ __forceinline__ uint2 block_scan_u32(uint32_t v, uint32_t* s_w, int lane, int warp) {
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();   // s_w may still be read from a previous scan
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const uint32_t c = s_w[w];
    if (w < warp) off += c;
    tot += c;
  }
  return make_uint2(incl + off, tot);
}

// ------------------------------------------------------------------------------------------ histogram + CDF table
__global__ void __launch_bounds__(FS_HIST_THREADS) fs_hist_kernel(
    const float* __restrict__ times, int64_t n, const float* __restrict__ scores, uint32_t* __restrict__ hist,
    uint32_t* __restrict__ max_enc, int32_t* __restrict__ nan_flag, uint32_t* done_counter, uint2* __restrict__ lut) {
  __shared__ uint32_t s_hist[FS_BINS];
  __shared__ uint32_t s_w[FS_HIST_THREADS / 32];
  __shared__ uint32_t s_max[FS_HIST_THREADS / 32];
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < FS_BINS; i += FS_HIST_THREADS) s_hist[i] = 0;
  __syncthreads();
  const uint64_t pol = make_evict_last_policy();
  const bool vec_t = (reinterpret_cast<uintptr_t>(times) & 15) == 0;
  const bool vec_s = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
  float vmax = -INFINITY;
  bool has_nan = false;
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {   // warp-uniform trip count
    const int64_t base = blk * 1024 + int64_t(tid) * 4;
    const int cnt = int(max((long long)0, min((long long)4, (long long)(n - base))));
    uint32_t d[4] = {0u, 0u, 0u, 0u};
    if (cnt == 4 && vec_t) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(times + base));
      d[0] = time_key(t4.x) >> 20; d[1] = time_key(t4.y) >> 20; d[2] = time_key(t4.z) >> 20; d[3] = time_key(t4.w) >> 20;
    } else {
      for (int i = 0; i < cnt; ++i) d[i] = time_key(__ldg(times + base + i)) >> 20;
    }
    if (scores != nullptr) {
      float s4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (cnt == 4 && vec_s) {
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
    // skewed keys (survival times share a handful of exponents): aggregate equal bins per thread / per warp
    const bool same4 = (cnt == 4) && d[0] == d[1] && d[1] == d[2] && d[2] == d[3];
    const uint32_t dl = __shfl_sync(0xffffffffu, d[0], 0);
    if (__all_sync(0xffffffffu, same4 && d[0] == dl)) {
      if (lane == 0) atomicAdd(&s_hist[dl], 128u);
    } else if (same4) {
      atomicAdd(&s_hist[d[0]], 4u);
    } else {
      for (int i = 0; i < cnt; ++i) atomicAdd(&s_hist[d[i]], 1u);
    }
  }
  __syncthreads();
  for (int i = tid; i < FS_BINS; i += FS_HIST_THREADS) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(hist + i, c);
  }
  if (scores != nullptr) {
    vmax = warp_max(vmax);
    if (lane == 0) s_max[warp] = float_order_enc(vmax);
    const unsigned any_nan = __ballot_sync(0xffffffffu, has_nan);
    if (lane == 0 && any_nan) atomicOr(nan_flag, 1);
    __syncthreads();
    if (tid == 0) {
      uint32_t m = s_max[0];
      for (int w = 1; w < FS_HIST_THREADS / 32; ++w) m = max(m, s_max[w]);
      atomicMax(max_enc, m);
    }
  }
  // the last block to finish builds the CDF table
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  constexpr int PER = FS_BINS / FS_HIST_THREADS;   // 16 consecutive bins per thread
  uint32_t c[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    c[i] = ld_cg_u32(hist + tid * PER + i);
    sum += c[i];
  }
  const uint2 sc = block_scan_u32<FS_HIST_THREADS / 32>(sum, s_w, lane, warp);
  uint32_t run = sc.x - sum;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    lut[tid * PER + i] = make_uint2(run, c[i]);
    run += c[i];
  }
}

// ------------------------------------------------------------------------------------------ exact bucket sizes
__global__ void __launch_bounds__(256) fs_count_kernel(
    const float* __restrict__ times, int64_t n, const uint2* __restrict__ lut, uint32_t mult, int nb,
    uint32_t* __restrict__ bucket_count, uint32_t* done_counter, uint32_t* __restrict__ bucket_base,
    uint4* __restrict__ work, uint32_t* __restrict__ params) {
  __shared__ uint32_t s_cnt[FS_MAX_BUCKETS];
  __shared__ uint32_t s_w[8];
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < FS_MAX_BUCKETS; i += 256) s_cnt[i] = 0;
  __syncthreads();
  const bool vec_t = (reinterpret_cast<uintptr_t>(times) & 15) == 0;
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {
    const int64_t base = blk * 1024 + int64_t(tid) * 4;
    const int cnt = int(max((long long)0, min((long long)4, (long long)(n - base))));
    if (cnt == 4 && vec_t) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(times + base));
      const uint32_t b0 = bucket_of(time_key(t4.x), lut, mult), b1 = bucket_of(time_key(t4.y), lut, mult);
      const uint32_t b2 = bucket_of(time_key(t4.z), lut, mult), b3 = bucket_of(time_key(t4.w), lut, mult);
      if (b0 == b1 && b1 == b2 && b2 == b3) {
        atomicAdd(&s_cnt[b0], 4u);
      } else {
        atomicAdd(&s_cnt[b0], 1u); atomicAdd(&s_cnt[b1], 1u); atomicAdd(&s_cnt[b2], 1u); atomicAdd(&s_cnt[b3], 1u);
      }
    } else {
      for (int i = 0; i < cnt; ++i) atomicAdd(&s_cnt[bucket_of(time_key(__ldg(times + base + i)), lut, mult)], 1u);
    }
  }
  __syncthreads();
  for (int i = tid; i < nb; i += 256) {
    const uint32_t c = s_cnt[i];
    if (c) atomicAdd(bucket_count + i, c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: bucket offsets (exclusive scan) and the work list of the local sort: one item per bucket, or one per
  // FS_CAP-sized chunk of an oversized bucket.  item = (bucket, first element inside the bucket, count, bucket size)
  constexpr int PER = FS_MAX_BUCKETS / 256;   // 4 consecutive buckets per thread
  uint32_t c[PER], k[PER], csum = 0, ksum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int b = tid * PER + i;
    c[i] = (b < nb) ? ld_cg_u32(bucket_count + b) : 0u;
    k[i] = (c[i] + FS_CAP - 1) / FS_CAP;
    csum += c[i];
    ksum += k[i];
  }
  const uint2 cs = block_scan_u32<8>(csum, s_w, lane, warp);
  const uint2 ks = block_scan_u32<8>(ksum, s_w, lane, warp);
  uint32_t crun = cs.x - csum, krun = ks.x - ksum;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int b = tid * PER + i;
    if (b < nb) bucket_base[b] = crun;
    for (uint32_t q = 0; q < k[i]; ++q)
      work[krun + q] = make_uint4(uint32_t(b), q * FS_CAP, min(uint32_t(FS_CAP), c[i] - q * FS_CAP), c[i]);
    crun += c[i];
    krun += k[i];
  }
  if (tid == 255) {
    bucket_base[nb] = cs.y;
    params[0] = ks.y;   // number of work items
  }
}

// ------------------------------------------------------------------------------------------ stable partition
__global__ void __launch_bounds__(P_THREADS, 2) fs_partition_kernel(
    const float* __restrict__ times, const float* __restrict__ status, int64_t n, const uint2* __restrict__ lut,
    uint32_t mult, int nb, int nbits, const uint32_t* __restrict__ bucket_base, uint32_t* lookback,
    uint32_t* tile_counter, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    int32_t* __restrict__ nonbinary_flag) {
  __shared__ uint16_t s_wcnt[P_WARPS][FS_MAX_BUCKETS];   // 32 KB: per-warp bucket counters (a tile has 8192 keys)
  __shared__ uint32_t s_start[FS_MAX_BUCKETS];           // tile-local first slot of every bucket
  __shared__ uint32_t s_gbase[FS_MAX_BUCKETS];           // global slot of the tile's first element of every bucket
  __shared__ uint32_t s_w[P_WARPS];
  __shared__ uint32_t s_tile;
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_keys = s_dyn;
  uint32_t* s_vals = s_dyn + P_TILE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t num_tiles = (n + P_TILE - 1) / P_TILE;
  // tiles are handed out by an atomic ticket in the order blocks ask for them: a tile only waits on tiles whose
  // ticket is held by a running block, so the look-back cannot deadlock
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < P_WARPS * FS_MAX_BUCKETS / 2; i += P_THREADS) reinterpret_cast<uint32_t*>(&s_wcnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  if (tile >= num_tiles) return;
  const int64_t tile_base = tile * P_TILE;
  const int n_valid = int(min((long long)P_TILE, (long long)(n - tile_base)));

  uint32_t key[P_ITEMS], val[P_ITEMS], rd[P_ITEMS];   // rd = rank (bits 0-15) | bucket (bits 16-25)
  bool nonbinary = false;
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    const int it = warp * (32 * P_ITEMS) + j * 32 + lane;
    const int64_t g = tile_base + it;
    if (it < n_valid) {
      key[j] = time_key(__ldg(times + g));
      val[j] = (status != nullptr) ? __float_as_uint(__ldg(status + g)) : 0u;   // needed only at the scatter
    } else {
      key[j] = 0xffffffffu;   // padding (last tile only) ranks behind every real key of the last bucket
      val[j] = 0u;
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    const int it = warp * (32 * P_ITEMS) + j * 32 + lane;
    const uint32_t d = (it < n_valid) ? bucket_of(key[j], lut, mult) : uint32_t(nb - 1);
    // peers = lanes of this warp whose key goes to the same bucket (one ballot per bucket bit)
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 10; ++b) {
      if (b < nbits) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
      }
    }
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if (lane == leader) {
      prev = s_wcnt[warp][d];
      s_wcnt[warp][d] = uint16_t(prev + __popc(peers));
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rd[j] = (prev + __popc(peers & lt_mask)) | (d << 16);
    __syncwarp();
  }
  __syncthreads();

  // thread tid owns buckets tid and tid + 512: exclusive scan over the warps, tile totals, tile-local bucket starts
  uint32_t tot[2];
  uint32_t* lb = lookback + tile * FS_MAX_BUCKETS;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int d = tid + q * P_THREADS;
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < P_WARPS; ++w) {
      const uint32_t c = s_wcnt[w][d];
      s_wcnt[w][d] = uint16_t(total);
      total += c;
    }
    tot[q] = total;
    if (d < nb) st_volatile_u32(lb + d, (tile == 0 ? RS_FLAG_INCL : RS_FLAG_AGG) | total);
  }
  const uint2 sc0 = block_scan_u32<P_WARPS>(tot[0], s_w, lane, warp);
  const uint2 sc1 = block_scan_u32<P_WARPS>(tot[1], s_w, lane, warp);
  s_start[tid] = sc0.x - tot[0];
  s_start[tid + P_THREADS] = sc0.y + sc1.x - tot[1];
  __syncthreads();

  // bring the tile into bucket order in shared memory
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    const uint32_t d = rd[j] >> 16;
    const uint32_t pos = s_start[d] + s_wcnt[warp][d] + (rd[j] & 0xffffu);
    const int it = warp * (32 * P_ITEMS) + j * 32 + lane;
    uint32_t v = uint32_t(tile_base + it);
    if (status != nullptr && it < n_valid) {
      const float st = __uint_as_float(val[j]);
      if (st != 0.0f) v |= 0x80000000u;
      if (st != 0.0f && st != 1.0f) nonbinary = true;
    }
    s_keys[pos] = key[j];
    s_vals[pos] = v;
  }
  if (nonbinary) atomicOr(nonbinary_flag, 1);

  // decoupled look-back over the earlier tiles, both buckets of the thread in lockstep
  uint32_t excl[2] = {0u, 0u};
  if (tile > 0) {
    int64_t p[2] = {tile - 1, tile - 1};
    bool done[2] = {tid >= nb, tid + P_THREADS >= nb};
    while (!(done[0] && done[1])) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (done[q]) continue;
        const int d = tid + q * P_THREADS;
        uint32_t v[P_LOOKBACK];
#pragma unroll
        for (int u = 0; u < P_LOOKBACK; ++u)
          v[u] = (p[q] - u >= 0) ? ld_volatile_u32(lookback + (p[q] - u) * FS_MAX_BUCKETS + d) : RS_FLAG_INCL;
#pragma unroll
        for (int u = 0; u < P_LOOKBACK; ++u) {
          if (!done[q]) {
            while ((v[u] & RS_FLAG_MASK) == 0) {
              __nanosleep(100);
              v[u] = ld_volatile_u32(lookback + (p[q] - u) * FS_MAX_BUCKETS + d);
            }
            excl[q] += v[u] & RS_VALUE_MASK;
            if (v[u] & RS_FLAG_INCL) done[q] = true;
          }
        }
        p[q] -= P_LOOKBACK;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int d = tid + q * P_THREADS;
      if (d < nb) st_volatile_u32(lb + d, RS_FLAG_INCL | (excl[q] + tot[q]));
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int d = tid + q * P_THREADS;
    s_gbase[d] = (d < nb) ? __ldg(bucket_base + d) + excl[q] : 0u;
  }
  __syncthreads();

#pragma unroll 4
  for (int j = 0; j < P_ITEMS; ++j) {
    const int i = j * P_THREADS + tid;
    if (i < n_valid) {
      const uint32_t k = s_keys[i];
      const uint32_t d = bucket_of(k, lut, mult);
      const uint32_t dst = s_gbase[d] + (uint32_t(i) - s_start[d]);
      keys_out[dst] = k;
      vals_out[dst] = s_vals[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ local sort + gather
__global__ void __launch_bounds__(L_THREADS, 1) fs_local_sort_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, const uint32_t* __restrict__ bucket_base,
    const uint4* __restrict__ work, const uint32_t* __restrict__ params, int32_t* __restrict__ perm_out,
    const float* __restrict__ scores, const uint32_t* __restrict__ max_enc, float* __restrict__ saved_s,
    int32_t* max_count, int32_t* __restrict__ max_list, int max_list_cap, int32_t* fallback) {
  __shared__ uint32_t s_wcnt[L_WARPS][L_RADIX];   // 32 KB
  __shared__ uint32_t s_dstart[L_RADIX];
  __shared__ uint32_t s_w[L_WARPS];
  __shared__ uint32_t s_or;
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_keys = s_dyn;
  uint32_t* s_vals = s_dyn + FS_CAP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_work = ld_cg_u32(params);
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float smax = (scores != nullptr) ? float_order_dec(ld_cg_u32(max_enc)) : 0.f;
  const uint64_t pol = make_evict_last_policy();

  for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const uint4 item = work[wi];
    const uint32_t bucket_first = __ldg(bucket_base + item.x);
    const uint32_t base = bucket_first + item.y;
    const int cnt = int(item.z);
    if (tid == 0) s_or = 0;
    uint32_t key[L_ITEMS], val[L_ITEMS];
    const uint32_t key0 = __ldg(keys_in + bucket_first);   // any key of the bucket
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < L_ITEMS; ++j) {
      const int it = warp * (32 * L_ITEMS) + j * 32 + lane;
      if (it < cnt) {
        key[j] = __ldg(keys_in + base + it);
        val[j] = __ldg(vals_in + base + it);
        diff |= key[j] ^ key0;
      } else {
        key[j] = 0xffffffffu;   // padding sorts behind every real key in every pass
        val[j] = 0xffffffffu;
      }
    }
    __syncthreads();            // s_or cleared; previous item's output phase finished with shared memory
    diff = __reduce_or_sync(0xffffffffu, diff);
    if (lane == 0 && diff) atomicOr(&s_or, diff);
    __syncthreads();
    diff = s_or;
    int passes = 0, bpp = L_BITS, lo = 0;
    if (item.w > uint32_t(FS_CAP)) {
      // a chunk of an oversized bucket: legal only when the whole bucket is one key (then it is already in order)
      if (diff != 0 && tid == 0) atomicExch(fallback, 1);
    } else if (diff != 0) {
      lo = __ffs(diff) - 1;
      const int width = 32 - __clz(diff) - lo;   // only these key bits differ inside the bucket
      passes = (width + L_BITS - 1) / L_BITS;
      bpp = (width + passes - 1) / passes;
    }
    if (passes == 0) {   // nothing to sort: identity placement
#pragma unroll
      for (int j = 0; j < L_ITEMS; ++j) {
        const int it = warp * (32 * L_ITEMS) + j * 32 + lane;
        s_vals[it] = val[j];
      }
    }
    for (int ps = 0; ps < passes; ++ps) {
      const int shift = lo + ps * bpp;
      const uint32_t dmask = (1u << bpp) - 1u;
      for (int i = tid; i < L_WARPS * L_RADIX; i += L_THREADS) (&s_wcnt[0][0])[i] = 0;
      __syncthreads();
      uint32_t rank[L_ITEMS];
#pragma unroll
      for (int j = 0; j < L_ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & dmask;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < L_BITS; ++b) {
          if (b < bpp) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
          }
        }
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader) {
          prev = s_wcnt[warp][d];
          s_wcnt[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[j] = prev + __popc(peers & lt_mask);
        __syncwarp();
      }
      __syncthreads();
      {   // thread tid owns digit tid: exclusive scan over the warps, then over the digits
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < L_WARPS; ++w) {
          const uint32_t c = s_wcnt[w][tid];
          s_wcnt[w][tid] = total;
          total += c;
        }
        const uint2 sc = block_scan_u32<L_WARPS>(total, s_w, lane, warp);
        s_dstart[tid] = sc.x - total;
      }
      __syncthreads();
      const bool last = (ps == passes - 1);
#pragma unroll
      for (int j = 0; j < L_ITEMS; ++j) {
        const uint32_t d = (key[j] >> shift) & dmask;
        const uint32_t pos = s_dstart[d] + s_wcnt[warp][d] + rank[j];
        if (!last) s_keys[pos] = key[j];
        s_vals[pos] = val[j];
      }
      __syncthreads();
      if (!last) {
#pragma unroll
        for (int j = 0; j < L_ITEMS; ++j) {
          const int it = warp * (32 * L_ITEMS) + j * 32 + lane;
          key[j] = s_keys[it];
          val[j] = s_vals[it];
        }
      }
    }
    __syncthreads();
    // output: permutation word (index | event << 31) and, fused, the forward gather s~ = scores[index] - max
    // (scores kept L2-resident by the evict_last loads of the histogram pass)
#pragma unroll
    for (int j0 = 0; j0 < L_ITEMS; j0 += 8) {
      uint32_t v[8];
      float st[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = (j0 + u) * L_THREADS + tid;
        v[u] = (i < cnt) ? s_vals[i] : 0u;
      }
      if (scores != nullptr) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = (j0 + u) * L_THREADS + tid;
          st[u] = (i < cnt) ? ld_f32_hint(scores + (v[u] & 0x7fffffffu), pol) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = (j0 + u) * L_THREADS + tid;
        if (i < cnt) {
          perm_out[base + i] = int32_t(v[u]);
          if (scores != nullptr) {
            const float s = st[u] - smax;
            saved_s[base + i] = s;
            if (s == 0.f) {   // an argmax position (gradient through max(scores) in the backward pass)
              const int pos = atomicAdd(max_count, 1);
              if (pos < max_list_cap) max_list[pos] = int32_t(v[u] & 0x7fffffffu);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
int fs_sort_enqueue(const float* times, const float* status, const float* scores, int64_t n, const FastSortWs& w,
                    uint32_t* max_enc, int32_t* nan_flag, int32_t* nonbinary_flag, int32_t* perm_out, float* saved_s,
                    int32_t* max_count, int32_t* max_list, int max_list_cap, cudaStream_t stream) {
  MMBS_REQUIRE(n > 0 && n <= FS_MAX_N, "fast sort: n=%lld out of range", (long long)n);
  static PerDeviceOnce configured;
  if (configured.first()) {
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_DYN_SMEM));
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_DYN_SMEM));
  }
  const int nb = fs_num_buckets(n);
  int nbits = 0;
  while ((1 << nbits) < nb) ++nbits;
  const uint32_t mult = uint32_t((uint64_t(nb) << 32) / uint64_t(n));   // floor: bucket_of(.) <= nb - 1
  const int sms = sm_count();
  const int hist_grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 1024 * 2), int64_t(sms) * 8)));
  fs_hist_kernel<<<hist_grid, FS_HIST_THREADS, 0, stream>>>(times, n, scores, w.hist12, max_enc, nan_flag, w.counters + 0,
                                                            w.lut);
  MMBS_LAUNCH_CHECK();
  fs_count_kernel<<<hist_grid, 256, 0, stream>>>(times, n, w.lut, mult, nb, w.bucket_count, w.counters + 1, w.bucket_base,
                                                 w.work, w.params);
  MMBS_LAUNCH_CHECK();
  const int64_t tiles = fs_tiles(n);
  fs_partition_kernel<<<unsigned(tiles), P_THREADS, P_DYN_SMEM, stream>>>(times, status, n, w.lut, mult, nb, nbits,
                                                                         w.bucket_base, w.lookback, w.counters + 2,
                                                                         w.keys, w.vals, nonbinary_flag);
  MMBS_LAUNCH_CHECK();
  const int64_t max_work = nb + n / FS_CAP + 1;
  const int local_grid = int(std::min<int64_t>(max_work, sms));
  fs_local_sort_kernel<<<local_grid, L_THREADS, L_DYN_SMEM, stream>>>(w.keys, w.vals, w.bucket_base, w.work, w.params,
                                                                     perm_out, scores, max_enc, saved_s, max_count,
                                                                     max_list, max_list_cap, w.fallback);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

}  // namespace mmbs
