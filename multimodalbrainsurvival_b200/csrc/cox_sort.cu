// Bucketed Cox pipeline (sm_100a) for risk sets of 2049 .. FS_MAX_N samples: the whole forward pass is
//   fs_hist_kernel            one read of `times` (+ `scores`): 4096-bin histogram of the top 12 key bits, smallest /
//                             largest key, max(scores), NaN flag; its last block turns the histogram into a
//                             piecewise-linear CDF table  lut[bin] = (P, c)  and the two edge-bin corrections.
//   fs_partition_kernel       every sample goes to bucket  floor(nb * R(key) / n),  R(key) = P[bin] + frac * c[bin]
//                             (monotone in the key, so buckets are contiguous key ranges of ~5 K samples for any smooth
//                             distribution of survival times).  Ranking inside a tile is ONE shared-memory atomic per
//                             sample and a tile claims its slots in a bucket region with one global atomic per bucket:
//                             the partition is NOT stable and does not have to be (below).  8 B read + 8 B written.
//   fs_bucket_forward_kernel  one block per bucket: counting sort over ~4096 sub-buckets of the same rank estimate
//                             (~1.25 samples each, one shared-memory atomic per sample), then every sample finds its
//                             final rank by comparing (key, index) with the few members of its sub-bucket - a total
//                             order, so the result is the stable order whatever the atomics did.  The block then
//                             gathers scores through the sorted indices, writes perm / s~, scans exp(s~), fetches the
//                             sum over all earlier buckets by look-back (buckets are handed out by ticket, so every
//                             predecessor is running), and accumulates the loss terms; the last block reduces the loss.
// and the backward pass is fs_bucket_backward_kernel: per bucket, recompute C and w = status / (C + eps) from s~, suffix
// sums, gradient scatter; the last block applies the gradient through max(scores).
//
// Replaces  _, idx = torch.sort(-times); scores[idx]; status[idx]; exp; cumsum; log; mask; mean  of cox_loss()
//   /root/reference/1_HistoPathology/models.py:99-111 (and its three textual copies).
//
// Inputs this map cannot balance (a bucket over FS_CAP samples, a sub-bucket over FS_CMAX: heavy ties, densities with
// jumps inside a 12-bit bin) raise the device flag `fallback`; the LSD-sort pipeline of radix_sort.cu / cox.cu is
// enqueued behind and only runs when the flag is set - the decision never costs a host synchronisation.
#include <algorithm>

#include "cox_sort.cuh"

namespace mmbs {

constexpr int FS_HIST_THREADS = 256;
constexpr int P_THREADS = 512;
constexpr int P_ITEMS = 16;
constexpr int P_TILE = P_THREADS * P_ITEMS;        // 8192
constexpr int B_THREADS = 512;
constexpr int B_ITEMS = FS_CAP / B_THREADS;        // 16 (the skew below assumes 16)
constexpr int B_SKEW_WORDS = FS_CAP + FS_CAP / 16; // word p lives at p + p/16: 16-word runs start on distinct banks
constexpr float FS_EPS = 1e-5f;
static_assert(B_ITEMS == 16, "skewed shared-memory layout assumes 16 samples per thread");

__device__ __forceinline__ int skew(int p) { return p + (p >> 4); }

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// exp / log of the scan: MUFU-based (2^-21 relative; the loss and its gradient are specified to 1e-5)
__device__ __forceinline__ float fs_exp(float x) { return __expf(x); }
__device__ __forceinline__ float fs_log(float x) { return __logf(x); }

// rank estimate of a key in [0, n): piecewise-linear CDF over the 12-bit bins.  E: the edge-bin record in SHARED memory
// (read only by the few keys that fall into the two edge bins; lo_bin / hi_bin are register copies)
__device__ __forceinline__ uint32_t fs_rank(uint32_t key, const uint2* __restrict__ lut, uint32_t lo_bin, uint32_t hi_bin,
                                            const FsEdge* E) {
  const uint32_t bin = key >> 20;
  const uint2 e = __ldg(lut + bin);
  uint32_t in = uint32_t((uint64_t(key & 0xfffffu) * e.y) >> 20);
  if (bin == lo_bin) in = min(e.y - 1u, __float2uint_rz(__uint2float_rz(key - E->lo_base) * E->lo_scale));
  else if (bin == hi_bin) in = min(e.y - 1u, __float2uint_rz(__uint2float_rz(key - E->hi_base) * E->hi_scale));
  return e.x + in;
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

// ------------------------------------------------------------------------------------------ histogram + CDF table
__global__ void __launch_bounds__(FS_HIST_THREADS, 6) fs_hist_kernel(
    const float* __restrict__ times, int64_t n, const float* __restrict__ scores, uint32_t* __restrict__ hist,
    uint32_t* __restrict__ kext, uint32_t* __restrict__ max_enc, int32_t* __restrict__ nan_flag, uint32_t* done_counter,
    uint2* __restrict__ lut, FsEdge* __restrict__ edge) {
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
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {   // warp-uniform trip count
    const int64_t base = blk * 1024 + int64_t(tid) * 4;
    const int cnt = int(max((long long)0, min((long long)4, (long long)(n - base))));
    uint32_t k[4] = {0u, 0u, 0u, 0u};
    if (cnt == 4 && vec_t) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(times + base));
      k[0] = time_key(t4.x); k[1] = time_key(t4.y); k[2] = time_key(t4.z); k[3] = time_key(t4.w);
    } else {
      for (int i = 0; i < cnt; ++i) k[i] = time_key(__ldg(times + base + i));
    }
    for (int i = 0; i < cnt; ++i) {
      kmin = min(kmin, k[i]);
      kmax = max(kmax, k[i]);
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
    const uint32_t d0 = k[0] >> 20, d1 = k[1] >> 20, d2 = k[2] >> 20, d3 = k[3] >> 20;
    const bool same4 = (cnt == 4) && d0 == d1 && d1 == d2 && d2 == d3;
    const uint32_t dl = __shfl_sync(0xffffffffu, d0, 0);
    if (__all_sync(0xffffffffu, same4 && d0 == dl)) {
      if (lane == 0) atomicAdd(&s_hist[dl], 128u);
    } else if (same4) {
      atomicAdd(&s_hist[d0], 4u);
    } else {
      if (cnt > 0) atomicAdd(&s_hist[d0], 1u);
      if (cnt > 1) atomicAdd(&s_hist[d1], 1u);
      if (cnt > 2) atomicAdd(&s_hist[d2], 1u);
      if (cnt > 3) atomicAdd(&s_hist[d3], 1u);
    }
  }
  __syncthreads();
  for (int i = tid; i < FS_BINS; i += FS_HIST_THREADS) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(hist + i, c);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) {
    atomicMax(kext + 0, ~kmin);
    atomicMax(kext + 1, kmax);
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
  // the last block to finish builds the CDF table (barrier + one cumulative fence order the block's atomics)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
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
  const uint32_t gmin = ~ld_cg_u32(kext + 0), gmax = ld_cg_u32(kext + 1);
  const uint32_t lo_bin = gmin >> 20, hi_bin = gmax >> 20;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const uint32_t bin = uint32_t(tid * PER + i);
    lut[bin] = make_uint2(run, c[i]);
    // edge bins: interpolate over the occupied key range [first, last] of the bin only
    if (bin == lo_bin) {
      const uint32_t last = (lo_bin == hi_bin) ? gmax : ((bin << 20) | 0xfffffu);
      edge->lo_bin = bin;
      edge->lo_base = gmin;
      edge->lo_scale = float(c[i]) / float(last - gmin + 1u);
      if (lo_bin == hi_bin) {
        edge->hi_bin = 0xffffffffu;
        edge->hi_base = 0u;
        edge->hi_scale = 0.f;
      }
    } else if (bin == hi_bin) {
      edge->hi_bin = bin;
      edge->hi_base = bin << 20;
      edge->hi_scale = float(c[i]) / float(gmax - (bin << 20) + 1u);
    }
    run += c[i];
  }
}

// ------------------------------------------------------------------------------------------ partition
// dynamic shared memory: s_cnt[nbp] | s_start[nbp] | s_gbase[nbp] | staged pairs[P_TILE] | staged bucket ids[P_TILE] (u16)
// (nbp = nb rounded up to a multiple of the block size)
__global__ void __launch_bounds__(P_THREADS, 2) fs_partition_kernel(
    const float* __restrict__ times, const float* __restrict__ status, int64_t n, const uint2* __restrict__ lut,
    const FsEdge* __restrict__ edge, uint32_t mult2, int log_s, int nb, uint32_t* __restrict__ cursor,
    uint2* __restrict__ pairs_out, int32_t* __restrict__ nonbinary_flag, int32_t* __restrict__ fallback) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ uint32_t s_w[P_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (nb + P_THREADS - 1) / P_THREADS;
  const int nbp = per * P_THREADS;
  uint32_t* s_cnt = s_dyn;
  uint32_t* s_start = s_dyn + nbp;
  uint32_t* s_gbase = s_dyn + 2 * nbp;
  uint2* s_pairs = reinterpret_cast<uint2*>(s_dyn + 3 * nbp);
  uint16_t* s_bid = reinterpret_cast<uint16_t*>(s_dyn + 3 * nbp + 2 * P_TILE);
  __shared__ FsEdge s_E;
  if (tid < 8) reinterpret_cast<uint32_t*>(&s_E)[tid] = reinterpret_cast<const uint32_t*>(edge)[tid];
  for (int i = tid; i < nbp; i += P_THREADS) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t lo_bin = s_E.lo_bin, hi_bin = s_E.hi_bin;
  const int64_t tile_base = int64_t(blockIdx.x) * P_TILE;
  const int n_valid = int(min((long long)P_TILE, (long long)(n - tile_base)));

  uint32_t key[P_ITEMS], br[P_ITEMS];   // br = bucket << 16 | rank inside (tile, bucket); all ones: no sample
  uint32_t ev = 0;                      // event bits of the thread's samples
  bool nonbinary = false;
  const bool vec = n_valid == P_TILE && (reinterpret_cast<uintptr_t>(times) & 15) == 0 &&
                   (status == nullptr || (reinterpret_cast<uintptr_t>(status) & 15) == 0);
  // sample (q, k) of the thread is tile element 4 * (q * P_THREADS + tid) + k
#pragma unroll
  for (int q = 0; q < P_ITEMS / 4; ++q) {
    const int e0 = 4 * (q * P_THREADS + tid);
    float t4[4], s4[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(times + tile_base + e0));
      t4[0] = a.x; t4[1] = a.y; t4[2] = a.z; t4[3] = a.w;
      if (status != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(status + tile_base + e0));
        s4[0] = b.x; s4[1] = b.y; s4[2] = b.z; s4[3] = b.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = e0 + k < n_valid;
        t4[k] = ok ? __ldg(times + tile_base + e0 + k) : 0.f;
        s4[k] = (ok && status != nullptr) ? __ldg(status + tile_base + e0 + k) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = q * 4 + k;
      key[j] = time_key(t4[k]);
      if (s4[k] != 0.f) ev |= 1u << j;
      nonbinary |= (s4[k] != 0.f && s4[k] != 1.f);
    }
  }
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    const int e = 4 * ((j >> 2) * P_THREADS + tid) + (j & 3);
    if (e < n_valid) {
      const uint32_t b = __umulhi(fs_rank(key[j], lut, lo_bin, hi_bin, &s_E), mult2) >> log_s;
      br[j] = (b << 16) | atomicAdd(&s_cnt[b], 1u);
    } else {
      br[j] = 0xffffffffu;
    }
  }
  if (nonbinary) atomicOr(nonbinary_flag, 1);
  __syncthreads();

  // thread tid owns buckets tid*per .. +per: tile-local starts, and the tile's slots in every bucket region
  {
    uint32_t sum = 0;
    for (int k = 0; k < per; ++k) sum += s_cnt[tid * per + k];
    const uint2 sc = block_scan_u32<P_THREADS / 32>(sum, s_w, lane, warp);
    uint32_t run = sc.x - sum;
    for (int k = 0; k < per; ++k) {
      const int b = tid * per + k;
      const uint32_t c = s_cnt[b];
      s_start[b] = run;
      run += c;
      if (c != 0u) s_gbase[b] = atomicAdd(cursor + b, c);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    if (br[j] != 0xffffffffu) {
      const int e = 4 * ((j >> 2) * P_THREADS + tid) + (j & 3);
      const uint32_t b = br[j] >> 16;
      const uint32_t pos = s_start[b] + (br[j] & 0xffffu);
      s_pairs[pos] = make_uint2(key[j], uint32_t(tile_base + e) | (((ev >> j) & 1u) << 31));
      s_bid[pos] = uint16_t(b);
    }
  }
  __syncthreads();
  bool overflow = false;
#pragma unroll 4
  for (int i = tid; i < n_valid; i += P_THREADS) {
    const uint2 pr = s_pairs[i];
    const uint32_t b = s_bid[i];
    const uint32_t dst = s_gbase[b] + (uint32_t(i) - s_start[b]);
    if (dst < uint32_t(FS_CAP)) pairs_out[size_t(b) * FS_CAP + dst] = pr;
    else overflow = true;
  }
  if (overflow) atomicExch(fallback, 1);
}

// ------------------------------------------------------------------------------------------ per-bucket forward
// dynamic shared memory: s_a[B_SKEW_WORDS] | s_b[B_SKEW_WORDS] | s_cnt[4097]
constexpr int B_DYN_SMEM = (2 * B_SKEW_WORDS + (1 << FS_LOG_S_MAX) + 8) * 4;

__global__ void __launch_bounds__(B_THREADS, 2) fs_bucket_forward_kernel(
    const uint2* __restrict__ pairs, const uint32_t* __restrict__ cursor, const uint2* __restrict__ lut,
    const FsEdge* __restrict__ edge, uint32_t mult2, int log_s, int nb, int64_t n, uint32_t* counters,
    const float* __restrict__ scores, const float* __restrict__ status, const uint32_t* __restrict__ max_enc,
    int32_t* nan_flag, const int32_t* __restrict__ nonbinary_flag, int32_t* __restrict__ perm_out,
    float* __restrict__ saved_s, int32_t* max_count, int32_t* __restrict__ max_list, double* agg_val,
    double* __restrict__ exp_prefix, double* __restrict__ wsum, double* loss_part, uint32_t* __restrict__ bucket_base,
    uint32_t* __restrict__ bucket_cnt, float* __restrict__ loss_out, int32_t* __restrict__ flags_out, int32_t* fallback) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ double s_red[B_THREADS / 32];
  __shared__ uint32_t s_w[B_THREADS / 32];
  __shared__ uint32_t s_w2[B_THREADS / 32];
  __shared__ uint32_t s_misc[2];
  uint32_t* s_a = s_dyn;                       // keys, then the sorted payloads (skewed)
  uint32_t* s_b = s_dyn + B_SKEW_WORDS;        // payloads, then s~ | event << 31 (skewed)
  uint32_t* s_cnt = s_dyn + 2 * B_SKEW_WORDS;  // sub-bucket counters / starts, one sentinel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = B_THREADS / 32;
  // one thread decides for the block (the flag may be raised by a running block at any time) and takes the ticket:
  // buckets are handed out in ticket order, so every predecessor of a bucket is running or done
  if (tid == 0) {
    const bool stop = *reinterpret_cast<volatile int32_t*>(fallback) != 0;   // the pipeline already gave up
    s_misc[1] = stop ? 1u : 0u;
    if (!stop) s_misc[0] = atomicAdd(counters + 1, 1u);
  }
  __shared__ FsEdge s_E;
  if (tid < 8) reinterpret_cast<uint32_t*>(&s_E)[tid] = reinterpret_cast<const uint32_t*>(edge)[tid];
  const int S = 1 << log_s;
  for (int i = tid; i <= S; i += B_THREADS) s_cnt[i] = 0;
  __syncthreads();
  if (s_misc[1] != 0u) return;
  const uint32_t lo_bin = s_E.lo_bin, hi_bin = s_E.hi_bin;
  const int b = int(s_misc[0]);
  const int cnt = int(min(__ldg(cursor + b), uint32_t(FS_CAP)));
  uint32_t base;
  {
    uint32_t part = 0;
    for (int q = tid; q < b; q += B_THREADS) part += min(__ldg(cursor + q), uint32_t(FS_CAP));
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_w[warp] = part;
    __syncthreads();
    base = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) base += s_w[w];
  }

  // ---- counting sort over the sub-buckets of the rank estimate
  uint32_t key[B_ITEMS], val[B_ITEMS], dr[B_ITEMS];   // dr = sub-bucket << 16 | arrival rank inside it
  const uint2* src = pairs + size_t(b) * FS_CAP;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    const int i = j * B_THREADS + tid;
    const uint2 pr = (i < cnt) ? __ldg(src + i) : make_uint2(0xffffffffu, 0xffffffffu);
    key[j] = pr.x;
    val[j] = pr.y;
  }
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    const int i = j * B_THREADS + tid;
    if (i < cnt) {
      const uint32_t d = __umulhi(fs_rank(key[j], lut, lo_bin, hi_bin, &s_E), mult2) & uint32_t(S - 1);
      dr[j] = (d << 16) | atomicAdd(&s_cnt[d], 1u);
    } else {
      dr[j] = 0xffffffffu;
    }
  }
  __syncthreads();
  uint32_t kmax;   // largest sub-bucket of the block
  {
    const int per = (S + B_THREADS - 1) / B_THREADS;
    uint32_t sum = 0, big = 0;
    for (int k = 0; k < per; ++k) {
      const int d = tid * per + k;
      const uint32_t c = (d < S) ? s_cnt[d] : 0u;
      sum += c;
      big = max(big, c);
    }
    big = __reduce_max_sync(0xffffffffu, big);
    if (lane == 0) s_w2[warp] = big;
    const uint2 sc = block_scan_u32<NW>(sum, s_w, lane, warp);   // (its barriers also publish s_w2)
    uint32_t run = sc.x - sum;
    for (int k = 0; k < per; ++k) {
      const int d = tid * per + k;
      if (d < S) {
        const uint32_t c = s_cnt[d];
        s_cnt[d] = run;
        run += c;
      }
    }
    if (tid == 0) s_cnt[S] = uint32_t(cnt);
    kmax = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) kmax = max(kmax, s_w2[w]);
    // heavy ties / a density jump inside a bin: the finish below would be quadratic.  Give the whole input up (the LSD
    // pipeline redoes it) - but publish, so that successors waiting on this bucket's sum are released.
    if (kmax > uint32_t(FS_CMAX)) {
      if (tid == 0) {
        atomicExch(fallback, 1);
        st_relaxed_f64(agg_val + b, -0.0);
      }
      return;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    if (dr[j] != 0xffffffffu) {
      const uint32_t pos = s_cnt[dr[j] >> 16] + (dr[j] & 0xffffu);
      s_a[pos] = key[j];
      s_b[pos] = val[j];
    }
  }
  __syncthreads();
  // ---- final rank = sub-bucket start + members that precede in the total order (key, index).  The trip count is the
  // block's largest sub-bucket (~8 at 1.25 samples per sub-bucket), the 16 samples of a thread advance together.
  // Per sample: dr = sub-bucket << 16 | arrival rank << 8 | members found to precede; lohi = start | end << 13.
  {
    uint32_t lohi[B_ITEMS];
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const bool ok = dr[j] != 0xffffffffu;
      const uint32_t d = ok ? (dr[j] >> 16) : 0u;
      const uint32_t lo = s_cnt[d], hi = s_cnt[d + 1];
      lohi[j] = ok ? (lo | (hi << 13)) : 0u;                         // no sample: empty range
      dr[j] = ok ? ((d << 16) | ((dr[j] & 0xffffu) << 8)) : 0u;      // arrival rank < FS_CMAX = 128
    }
    for (uint32_t k = 0; k < kmax; ++k) {
#pragma unroll
      for (int j = 0; j < B_ITEMS; ++j) {
        const uint32_t q = (lohi[j] & 0x1fffu) + k;
        const bool act = q < (lohi[j] >> 13);
        const uint32_t k2 = s_a[act ? q : 0u];
        if (act && k2 <= key[j]) {
          if (k2 < key[j]) {
            ++dr[j];
          } else {   // equal keys: ascending original index (the stable order)
            const uint32_t mine = s_b[(lohi[j] & 0x1fffu) + ((dr[j] >> 8) & 0xffu)];
            if ((s_b[q] & 0x7fffffffu) < (mine & 0x7fffffffu)) ++dr[j];
          }
        }
      }
    }
    // final position and payload (re-read from the sub-bucket order; s_b is not overwritten before the next barrier)
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const bool ok = (lohi[j] >> 13) != 0u;
      const uint32_t lo = lohi[j] & 0x1fffu;
      val[j] = ok ? s_b[lo + ((dr[j] >> 8) & 0xffu)] : 0u;
      dr[j] = ok ? (lo + (dr[j] & 0xffu)) : 0xffffffffu;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j)
    if (dr[j] != 0xffffffffu) s_a[skew(int(dr[j]))] = val[j];
  __syncthreads();

  if (scores == nullptr) {   // mmbs_risk_order: the permutation only
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int i = j * B_THREADS + tid;
      if (i < cnt) perm_out[base + i] = int32_t(s_a[skew(i)]);
    }
    return;
  }

  // ---- gather s~ = scores[index] - max through the sorted payloads (scores L2-resident: evict_last since the histogram)
  const float smax = float_order_dec(__ldg(max_enc));
  const uint64_t pol = make_evict_last_policy();
  {
    uint32_t v[B_ITEMS];
    float st[B_ITEMS];
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int i = j * B_THREADS + tid;
      v[j] = (i < cnt) ? s_a[skew(i)] : 0u;
    }
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int i = j * B_THREADS + tid;
      st[j] = (i < cnt) ? ld_f32_hint(scores + (v[j] & 0x7fffffffu), pol) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int i = j * B_THREADS + tid;
      uint32_t enc = 0x7f800000u;   // no sample: s~ = -inf, no event
      if (i < cnt) {
        const float s = st[j] - smax;
        perm_out[base + i] = int32_t(v[j]);
        saved_s[base + i] = s;
        enc = (__float_as_uint(s) & 0x7fffffffu) | (v[j] & 0x80000000u);
        if (s == 0.f) {   // an argmax position (gradient through max(scores) in the backward pass)
          const int pos = atomicAdd(max_count, 1);
          if (pos < FS_MAX_LIST) max_list[pos] = int32_t(v[j] & 0x7fffffffu);
        }
      }
      s_b[skew(i)] = enc;
    }
  }
  __syncthreads();

  // ---- blocked arrangement: thread t owns sorted positions 16 t .. 16 t + 15
  float st[B_ITEMS], c[B_ITEMS];
  uint32_t evm = 0;
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    const uint32_t enc = s_b[tid * 17 + j];
    st[j] = __uint_as_float(enc | 0x80000000u);   // s~ <= 0
    evm |= (enc >> 31) << j;
    run += fs_exp(st[j]);                         // exp(-inf) = 0 beyond the bucket
    c[j] = run;
  }
  double total;
  const double incl = block_scan_f64<NW, false>(double(run), s_red, lane, warp, &total);
  const double off_in = incl - double(run);
  // publish the bucket's sum, then fetch the sum over all earlier buckets (every one of them holds an earlier ticket).
  // One 8-byte word is value and flag at once: the workspace is zeroed, a published sum is never +0.0 (an empty sum is
  // stored as -0.0), so there is nothing to order and no fence / cache invalidation on the spinning side.
  if (tid == 0) st_relaxed_f64(agg_val + b, total == 0.0 ? -0.0 : total);
  double pre = 0.0;
  for (int q = tid; q < b; q += B_THREADS) {
    unsigned long long bits;
    while ((bits = ld_relaxed_u64(reinterpret_cast<const unsigned long long*>(agg_val + q))) == 0ull) __nanosleep(40);
    pre += __longlong_as_double((long long)bits);
  }
  const double P = block_sum_f64<NW>(pre, s_red, lane, warp);
  const double off = P + off_in;
  // C = off + c[j], rounded to fp32 once: off as a (hi, lo) float pair (2 FADD per sample instead of fp64 converts)
  const float off_hi = float(off), off_lo = float(off - double(off_hi));

  const bool gather_status = (*nonbinary_flag != 0);
  float lsum = 0.f, ws = 0.f;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    const int p = tid * B_ITEMS + j;
    if (p < cnt) {
      float dl = float((evm >> j) & 1u);
      if (gather_status) dl = __ldg(status + (s_a[tid * 17 + j] & 0x7fffffffu));
      const float den = ((off_hi + c[j]) + off_lo) + FS_EPS;   // cumsum + eps (models.py:104)
      const float term = -(st[j] - fs_log(den)) * dl;          // models.py:104-105
      lsum += term;
      ws += __fdividef(dl, den);
      bad |= (term != term);
    }
  }
  const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0 && any_bad) atomicOr(nan_flag, 1);
  const double tl = block_sum_f64<NW>(double(lsum), s_red, lane, warp);
  const double tw = block_sum_f64<NW>(double(ws), s_red, lane, warp);
  if (tid == 0) {
    exp_prefix[b] = P;
    wsum[b] = tw;
    loss_part[b] = tl;
    bucket_base[b] = base;
    bucket_cnt[b] = uint32_t(cnt);
    __threadfence();
    s_misc[0] = (atomicAdd(counters + 2, 1u) == uint32_t(nb) - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_misc[0] == 0u) return;
  // the last block: loss = sum of the bucket partials / n   (.mean() over N, models.py:111)
  __threadfence();
  double t = 0.0;
  for (int q = tid; q < nb; q += B_THREADS) t += ld_volatile_f64(loss_part + q);
  t = block_sum_f64<NW>(t, s_red, lane, warp);
  if (tid == 0) {
    const int f = *reinterpret_cast<const volatile int32_t*>(nan_flag);
    if (*reinterpret_cast<volatile int32_t*>(fallback) == 0) {   // otherwise the LSD pipeline writes the result
      loss_out[0] = f ? __int_as_float(0x7fc00000) : float(t / double(n));
      if (flags_out) flags_out[0] = f;
    }
  }
}

// ------------------------------------------------------------------------------------------ per-bucket backward
constexpr int G_DYN_SMEM = 2 * B_SKEW_WORDS * 4;

__global__ void __launch_bounds__(B_THREADS, 2) fs_bucket_backward_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ saved_s, const float* __restrict__ status,
    const float* __restrict__ grad_loss, int64_t n, int nb, const uint32_t* __restrict__ bucket_base,
    const uint32_t* __restrict__ bucket_cnt, const double* __restrict__ exp_prefix, const double* __restrict__ wsum,
    double* gsum_part, uint32_t* counters, const int32_t* __restrict__ nonbinary_flag, const int32_t* __restrict__ max_count,
    const int32_t* __restrict__ max_list, double* __restrict__ gsum_total, float* grad_scores,
    const int32_t* __restrict__ fallback) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ double s_red[B_THREADS / 32];
  __shared__ uint32_t s_misc[2];
  uint32_t* s_v = s_dyn;                   // perm words (skewed)
  uint32_t* s_s = s_dyn + B_SKEW_WORDS;    // s~ bits (skewed)
  constexpr int NW = B_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*fallback != 0) return;
  const int b = blockIdx.x;
  const uint32_t base = __ldg(bucket_base + b);
  const int cnt = int(__ldg(bucket_cnt + b));
  const double P = __ldg(exp_prefix + b);
  double later = 0.0;
  for (int q = b + 1 + tid; q < nb; q += B_THREADS) later += __ldg(wsum + q);
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    const int i = j * B_THREADS + tid;
    uint32_t v = 0u, sb = 0xff800000u;   // no sample: s~ = -inf
    if (i < cnt) {
      v = uint32_t(__ldg(perm + base + i));
      sb = __float_as_uint(__ldg(saved_s + base + i));
    }
    s_v[skew(i)] = v;
    s_s[skew(i)] = sb;
  }
  const double S_later = block_sum_f64<NW>(later, s_red, lane, warp);   // (its barriers also publish s_v / s_s)
  const bool gather_status = (*nonbinary_flag != 0);
  const float scale = float(double(grad_loss[0]) / double(n));
  float e[B_ITEMS], cw[B_ITEMS];
  uint32_t evm = 0;
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    evm |= (s_v[tid * 17 + j] >> 31) << j;
    e[j] = fs_exp(__uint_as_float(s_s[tid * 17 + j]));
    run += e[j];
    cw[j] = run;
  }
  double total;
  const double incl = block_scan_f64<NW, false>(double(run), s_red, lane, warp, &total);
  const double off = P + (incl - double(run));
  const float off_hi = float(off), off_lo = float(off - double(off_hi));   // the same C as the forward pass
  float wrun = 0.f;
#pragma unroll
  for (int j = B_ITEMS - 1; j >= 0; --j) {
    float w = 0.f;
    if (tid * B_ITEMS + j < cnt) {
      float dl = float((evm >> j) & 1u);
      if (gather_status) dl = __ldg(status + (s_v[tid * 17 + j] & 0x7fffffffu));
      w = __fdividef(dl, ((off_hi + cw[j]) + off_lo) + FS_EPS);
    }
    wrun += w;
    cw[j] = wrun;   // inclusive suffix of w inside the thread
  }
  const double sincl = block_scan_f64<NW, true>(double(wrun), s_red, lane, warp, &total);
  const double soff = S_later + (sincl - double(wrun));
  const float soff_hi = float(soff), soff_lo = float(soff - double(soff_hi));
  float gs = 0.f;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    if (tid * B_ITEMS + j < cnt) {
      const uint32_t idx = s_v[tid * 17 + j] & 0x7fffffffu;
      float dl = float((evm >> j) & 1u);
      if (gather_status) dl = __ldg(status + idx);
      const float W = (soff_hi + cw[j]) + soff_lo;   // sum of w over this and all later positions
      const float g = -(dl - e[j] * W) * scale;
      grad_scores[idx] = g;                          // un-permute
      gs += g;
    }
  }
  const double tg = block_sum_f64<NW>(double(gs), s_red, lane, warp);
  if (tid == 0) {
    gsum_part[b] = tg;
    __threadfence();
    s_misc[0] = (atomicAdd(counters + 3, 1u) == uint32_t(nb) - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_misc[0] == 0u) return;
  // the last block: gradient through "- max(scores)": every argmax position receives -(sum_k g~_k) / count
  __threadfence();
  double t = 0.0;
  for (int q = tid; q < nb; q += B_THREADS) t += ld_volatile_f64(gsum_part + q);
  t = block_sum_f64<NW>(t, s_red, lane, warp);
  if (tid == 0) gsum_total[0] = t;
  const int mc = *max_count;
  if (mc <= FS_MAX_LIST) {
    const float fix = float(t / double(mc));
    for (int i = tid; i < mc; i += B_THREADS) {
      float* p = grad_scores + max_list[i];
      *reinterpret_cast<volatile float*>(p) = *reinterpret_cast<volatile float*>(p) - fix;
    }
  }
  if (tid == 0) counters[3] = 0u;   // a second backward over the same forward (retain_graph) counts from zero again
}

// ------------------------------------------------------------------------------------------ host side
static int fs_configure() {
  static PerDeviceOnce configured;
  if (configured.first()) {
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (3 * FS_MAX_BUCKETS + 2 * P_TILE) * 4 + P_TILE * 2));
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_bucket_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_DYN_SMEM));
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_bucket_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_DYN_SMEM));
  }
  return MMBS_OK;
}

int fs_forward_enqueue(const float* times, const float* status, const float* scores, int64_t n, const FastSortWs& w,
                       uint32_t* max_enc, int32_t* nan_flag, int32_t* nonbinary_flag, int32_t* perm_out, float* saved_s,
                       int32_t* max_count, int32_t* max_list, float* loss_out, int32_t* flags_out, cudaStream_t stream) {
  MMBS_REQUIRE(n > 0 && n <= FS_MAX_N, "bucketed Cox pipeline: n=%lld out of range", (long long)n);
  if (int rc = fs_configure()) return rc;
  const FsPlan p = fs_plan(n);
  const int sms = sm_count();
  const int hist_grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 1024 * 2), int64_t(sms) * 8)));
  fs_hist_kernel<<<hist_grid, FS_HIST_THREADS, 0, stream>>>(times, n, scores, w.hist12, w.kext, max_enc, nan_flag,
                                                            w.counters + 0, w.lut, w.edge);
  MMBS_LAUNCH_CHECK();
  const int64_t tiles = ceil_div(n, P_TILE);
  const int nbp = int(ceil_div(p.nb, P_THREADS)) * P_THREADS;
  const int p_smem = (3 * nbp + 2 * P_TILE) * 4 + P_TILE * 2;
  fs_partition_kernel<<<unsigned(tiles), P_THREADS, p_smem, stream>>>(times, status, n, w.lut, w.edge, p.mult2, p.log_s,
                                                                     p.nb, w.cursor, w.pairs, nonbinary_flag, w.fallback);
  MMBS_LAUNCH_CHECK();
  fs_bucket_forward_kernel<<<p.nb, B_THREADS, B_DYN_SMEM, stream>>>(
      w.pairs, w.cursor, w.lut, w.edge, p.mult2, p.log_s, p.nb, n, w.counters, scores, status, max_enc, nan_flag,
      nonbinary_flag, perm_out, saved_s, max_count, max_list, w.agg_val, w.exp_prefix, w.wsum, w.loss_part,
      w.bucket_base, w.bucket_cnt, loss_out, flags_out, w.fallback);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

int fs_backward_enqueue(const float* status, const int32_t* perm, const float* saved_s, const float* grad_loss, int64_t n,
                        const FastSortWs& w, const int32_t* nonbinary_flag, const int32_t* max_count,
                        const int32_t* max_list, double* gsum_total, float* grad_scores, cudaStream_t stream) {
  if (int rc = fs_configure()) return rc;
  const FsPlan p = fs_plan(n);
  fs_bucket_backward_kernel<<<p.nb, B_THREADS, G_DYN_SMEM, stream>>>(
      perm, saved_s, status, grad_loss, n, p.nb, w.bucket_base, w.bucket_cnt, w.exp_prefix, w.wsum, w.gsum_part, w.counters,
      nonbinary_flag, max_count, max_list, gsum_total, grad_scores, w.fallback);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

}  // namespace mmbs
