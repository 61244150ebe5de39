// Bucketed Cox pipeline (sm_100a) for risk sets of 2049 .. FS_MAX_N samples: the forward pass is
//   fs_hist_kernel            one read of `times` (a 1-in-8 sample of it for n >= 2 M): 4096-bin histogram of the top 12 key
//                             bits, smallest / largest key (+ max(scores), NaN flag when everything is read); its last block
//                             turns the histogram into a piecewise-linear CDF table  lut[bin] = (P, c)  and the two edge-bin
//                             corrections.
//   fs_partition_kernel       every sample goes to bucket  floor(nb * R(key) / n),  R(key) = P[bin] + frac * c[bin]
//                             (monotone in the key, so buckets are contiguous key ranges of ~6 K samples for any smooth
//                             distribution of survival times).  Ranking inside a tile is ONE shared-memory atomic per
//                             sample and a tile claims its slots in a bucket region with one global atomic per bucket:
//                             the partition is NOT stable and does not have to be (below).  The sample's score travels to
//                             the same slot of a parallel array.  12 B read + 12 B written per sample.
//   fs_bucket_sort_kernel     one block per bucket: counting sort over 4096 sub-buckets (~1.5 samples each, one
//                             shared-memory atomic per sample) of 64-bit (key, index, event) composites, then every sample
//                             finds its final rank by comparing its composite with the few slots from its sub-bucket's
//                             start - a total order, so the result is the stable order whatever the atomics did.  Scores
//                             move to their sorted slot through shared memory; the block writes perm, |s~| | event << 31,
//                             its sum of exp(s~) and one partial sum per 32 sorted positions.
//   fs_prefix_kernel          exclusive prefix of the buckets' sums (one block).
//   fs_row_loss_kernel        one WARP per 512 sorted positions: offset from the bucket prefix + the partial sums in front
//                             of the row, scan with shuffles, per-row partials of the loss terms and of w.
//   fs_loss_finalize_block    (cox_sort.cuh; runs inside cox.cu's one-block dispatch kernel) loss, suffix sums of w.
// and the backward pass is fs_row_backward_kernel (same warp rows: recompute C and w = status / (C + eps) from s~, suffix
// sums, gradient scatter) + fs_backward_finalize_block (the gradient through max(scores)).
//
// Replaces  _, idx = torch.sort(-times); scores[idx]; status[idx]; exp; cumsum; log; mask; mean  of cox_loss()
//   /root/reference/1_HistoPathology/models.py:99-111 (and its three textual copies).
//
// Inputs this map cannot balance (a bucket over FS_CAP samples, sub-buckets whose squared sizes add up to more than
// FS_SQ_BUDGET: heavy ties, densities with jumps inside a 12-bit bin) raise the device flag `fallback`; the LSD-sort
// pipeline of radix_sort.cu / cox.cu is enqueued behind and only runs when the flag is set - the decision never costs a
// host synchronisation.
#include <algorithm>
#include <cstdlib>

#include "cox_sort.cuh"

namespace mmbs {

constexpr int FS_HIST_THREADS = 256;
constexpr int P_THREADS = 1024;
constexpr int P_ITEMS = 16;
constexpr int P_TILE = P_THREADS * P_ITEMS;        // 16384: ~10 samples per (tile, bucket) run (8192-sample tiles at two
                                                   // blocks per SM: 119 us at 10 M, 4096 at three: 165 us, this: 111 us)
constexpr int B_THREADS = 512;
constexpr int B_ITEMS = FS_CAP / B_THREADS;        // 16 (the skew below assumes 16)
constexpr float FS_EPS = 1e-5f;
static_assert(B_ITEMS % 4 == 0, "blocked 16-byte loads");

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
// exp / log / reciprocal of the scan: bare MUFU operations (2^-21 relative; the loss and its gradient are specified to
// 1e-5).  No denormal fix-ups: the arguments of log / reciprocal are >= eps = 1e-5, exp results below 2^-126 count as 0.
__device__ __forceinline__ float fs_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fs_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fs_log(float x) { return fs_lg2(x) * 0.6931471805599453f; }
__device__ __forceinline__ float fs_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rank estimate of a key in [0, n): piecewise-linear CDF over the 12-bit bins.  E: the edge-bin record in SHARED memory
// (read only by the few keys that fall into the two edge bins; lo_bin / hi_bin are register copies)
__device__ __forceinline__ uint32_t fs_rank(uint32_t key, const uint2* __restrict__ lut, uint32_t lo_bin, uint32_t hi_bin,
                                            const FsEdge* E) {
  const uint32_t bin = key >> 20;
  const uint2 e = __ldg(lut + bin);
  uint32_t in = uint32_t((uint64_t(key & 0xfffffu) * e.y) >> 20);
  if (bin == lo_bin)   // (a sampled histogram may have missed smaller keys: they rank first)
    in = key < E->lo_base ? 0u : min(e.y - 1u, __float2uint_rz(__uint2float_rz(key - E->lo_base) * E->lo_scale));
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

// ------------------------------------------------------------------------------------------ histogram + CDF table
__global__ void __launch_bounds__(FS_HIST_THREADS, 6) fs_hist_kernel(
    const float* __restrict__ times, int64_t n, int sample_shift, int nb, const float* __restrict__ scores,
    uint32_t* __restrict__ hist, uint32_t* __restrict__ kext, uint32_t* __restrict__ max_enc, int32_t* __restrict__ nan_flag,
    uint32_t* done_counter, uint2* __restrict__ lut, FsEdge* __restrict__ edge) {
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
  // chunks of 1024 samples; with sample_shift > 0 one pseudo-randomly placed chunk out of every 2^sample_shift
  const int64_t nchunks = (n + 1023) / 1024;
  const int64_t ngroups = (nchunks + (int64_t(1) << sample_shift) - 1) >> sample_shift;
  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {   // warp-uniform trip count
    const int64_t blk = (grp << sample_shift) +
                        (sample_shift ? int64_t((uint32_t(grp) * 2654435761u) >> (32 - sample_shift)) : 0);
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
  if (tid == 0) {   // sc.y = m, the samples counted
    edge->mult = (sc.y <= uint32_t(nb)) ? 0xffffffffu : uint32_t((uint64_t(nb) << 32) / uint64_t(sc.y));
    edge->pad = 0u;
  }
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
// dynamic shared memory: s_cnt[nbp] (re-used for the slot deltas) | s_start[nbp] | staged pairs[P_TILE] | staged scores[P_TILE]
// (nbp = nb rounded up to a multiple of the block size).  A staged pair is (key, e | event << 14 | bucket << 15) with e the
// sample's position inside the tile: the write-out pass needs no separate bucket-id array.
constexpr uint32_t P_E_MASK = P_TILE - 1;
constexpr int P_E_BITS = 14;
static_assert(P_TILE == (1 << P_E_BITS) && FS_MAX_BUCKETS <= 4096, "staged payload packs 14 + 1 + 12 bits");

template <bool FULL>   // FULL: the tile holds P_TILE samples and `times` / `status` / `scores` are 16-byte aligned
__device__ __forceinline__ void partition_tile(const float* __restrict__ times, const float* __restrict__ status,
                                               const float* __restrict__ scores, bool do_max, uint32_t* __restrict__ max_enc,
                                               int32_t* __restrict__ nan_flag, int64_t tile_base, int n_valid,
                                               const uint2* __restrict__ lut, int nb, int per,
                                               uint32_t* __restrict__ cursor, uint2* __restrict__ pairs_out,
                                               float* __restrict__ sc_out, int32_t* __restrict__ nonbinary_flag,
                                               int32_t* __restrict__ fallback, uint32_t* s_cnt, uint32_t* s_start,
                                               uint2* s_pairs, float* s_sc, const FsEdge* s_E, uint32_t* s_w) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lo_bin = s_E->lo_bin, hi_bin = s_E->hi_bin, mult = s_E->mult;
  uint32_t* s_delta = s_cnt;            // a bucket's count is dead once its start / slot delta are known
  uint32_t key[P_ITEMS], br[P_ITEMS];   // br = bucket << 16 | rank inside (tile, bucket)
  uint32_t ev = 0;                      // event bits of the thread's samples
  bool nonbinary = false;
  // sample (q, k) of the thread is tile element 4 * (q * P_THREADS + tid) + k
#pragma unroll
  for (int q = 0; q < P_ITEMS / 4; ++q) {
    const int e0 = 4 * (q * P_THREADS + tid);
    float t4[4], s4[4] = {0.f, 0.f, 0.f, 0.f};
    if (FULL) {
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
    if (FULL || e < n_valid) {
      const uint32_t b = min(__umulhi(fs_rank(key[j], lut, lo_bin, hi_bin, s_E), mult), uint32_t(nb - 1));
      br[j] = (b << 16) | atomicAdd(&s_cnt[b], 1u);
    } else {
      br[j] = 0xffffffffu;   // no sample
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
      if (c != 0u) s_delta[b] = atomicAdd(cursor + b, c) - run;   // slot in the bucket region = tile position + delta
      run += c;
    }
  }
  __syncthreads();
  // stage the tile in bucket order; the scores ride along (second read of the tile's scores: this pass also owns
  // max(scores) / the NaN flag when the histogram only sampled)
  {
    const uint64_t pol = make_evict_last_policy();
    float vmax = -INFINITY;
    bool has_nan = false;
#pragma unroll
    for (int q = 0; q < P_ITEMS / 4; ++q) {
      const int e0 = 4 * (q * P_THREADS + tid);
      float s4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (scores != nullptr) {
        if (FULL) {
          const float4 v = ld_f32x4_hint(scores + tile_base + e0, pol);
          s4[0] = v.x; s4[1] = v.y; s4[2] = v.z; s4[3] = v.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (e0 + k < n_valid) s4[k] = ld_f32_hint(scores + tile_base + e0 + k, pol);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = q * 4 + k;
        has_nan |= (s4[k] != s4[k]);
        vmax = fmaxf(vmax, s4[k]);
        if (FULL || br[j] != 0xffffffffu) {
          const uint32_t b = br[j] >> 16;
          const uint32_t pos = s_start[b] + (br[j] & 0xffffu);
          s_pairs[pos] = make_uint2(key[j], uint32_t(e0 + k) | (((ev >> j) & 1u) << P_E_BITS) | (b << (P_E_BITS + 1)));
          s_sc[pos] = s4[k];
        }
      }
    }
    if (do_max) {
      vmax = warp_max(vmax);
      const unsigned any_nan = __ballot_sync(0xffffffffu, has_nan);
      if (lane == 0) {
        atomicMax(max_enc, float_order_enc(vmax));
        if (any_nan) atomicOr(nan_flag, 1);
      }
    }
  }
  __syncthreads();
  bool overflow = false;
#pragma unroll
  for (int j = 0; j < P_ITEMS; ++j) {
    const int i = j * P_THREADS + tid;
    if (FULL || i < n_valid) {
      const uint2 pr = s_pairs[i];
      const uint32_t b = pr.y >> (P_E_BITS + 1);
      const uint32_t dst = uint32_t(i) + s_delta[b];
      // (3 slots stay free: the blocked 16-byte loads of the loss / backward kernels start at base & ~3)
      if (dst < uint32_t(FS_CAP - 3)) {
        const size_t o = size_t(b) * FS_CAP + dst;
        pairs_out[o] = make_uint2(pr.x, uint32_t(tile_base + (pr.y & P_E_MASK)) | (((pr.y >> P_E_BITS) & 1u) << 31));
        if (scores != nullptr) sc_out[o] = s_sc[i];
      } else {
        overflow = true;
      }
    }
  }
  if (overflow) atomicOr(fallback, 1);
}

__global__ void __launch_bounds__(P_THREADS, 1) fs_partition_kernel(
    const float* __restrict__ times, const float* __restrict__ status, const float* __restrict__ scores, int do_max,
    uint32_t* __restrict__ max_enc, int32_t* __restrict__ nan_flag, int64_t n, const uint2* __restrict__ lut,
    const FsEdge* __restrict__ edge, int nb, uint32_t* __restrict__ cursor, uint2* __restrict__ pairs_out,
    float* __restrict__ sc_out, int32_t* __restrict__ nonbinary_flag, int32_t* __restrict__ fallback) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ uint32_t s_w[P_THREADS / 32];
  __shared__ FsEdge s_E;
  const int tid = threadIdx.x;
  const int per = (nb + P_THREADS - 1) / P_THREADS;
  const int nbp = per * P_THREADS;
  uint32_t* s_cnt = s_dyn;
  uint32_t* s_start = s_dyn + nbp;
  uint2* s_pairs = reinterpret_cast<uint2*>(s_dyn + 2 * nbp);
  float* s_sc = reinterpret_cast<float*>(s_dyn + 2 * nbp + 2 * P_TILE);
  if (tid < 8) reinterpret_cast<uint32_t*>(&s_E)[tid] = reinterpret_cast<const uint32_t*>(edge)[tid];
  if (tid == 8) s_w[0] = (*reinterpret_cast<volatile int32_t*>(fallback) != 0) ? 1u : 0u;
  for (int i = tid; i < nbp; i += P_THREADS) s_cnt[i] = 0;
  __syncthreads();
  // an earlier tile already gave the input up (a bucket overflowed: heavy ties): the LSD pipeline will redo everything,
  // the remaining tiles have nothing to add (with 40 K copies per key the flag is up after a fifth of the tiles)
  if (s_w[0] != 0u) return;
  __syncthreads();   // (s_w is the scan scratch of partition_tile)
  const int64_t tile_base = int64_t(blockIdx.x) * P_TILE;
  const int n_valid = int(min((long long)P_TILE, (long long)(n - tile_base)));
  const bool full = n_valid == P_TILE && (reinterpret_cast<uintptr_t>(times) & 15) == 0 &&
                    (status == nullptr || (reinterpret_cast<uintptr_t>(status) & 15) == 0) &&
                    (scores == nullptr || (reinterpret_cast<uintptr_t>(scores) & 15) == 0);
  if (full)
    partition_tile<true>(times, status, scores, do_max != 0, max_enc, nan_flag, tile_base, n_valid, lut, nb, per, cursor,
                         pairs_out, sc_out, nonbinary_flag, fallback, s_cnt, s_start, s_pairs, s_sc, &s_E, s_w);
  else
    partition_tile<false>(times, status, scores, do_max != 0, max_enc, nan_flag, tile_base, n_valid, lut, nb, per, cursor,
                          pairs_out, sc_out, nonbinary_flag, fallback, s_cnt, s_start, s_pairs, s_sc, &s_E, s_w);
}

// ------------------------------------------------------------------------------------------ per-bucket sort
// Row j of a block holds the samples j * B_THREADS + tid.  f(j, ok) runs for every row of the 4-row groups that hold
// samples: groups of full rows with ok == true as a compile-time constant (no per-sample predicate), the one partial
// group with ok = (sample exists); empty groups are skipped by a block-uniform branch.
template <typename F>
__device__ __forceinline__ void for_rows(int cnt, int tid, F&& f) {
  const int nfull = cnt / B_THREADS;
#pragma unroll
  for (int g = 0; g < B_ITEMS / 4; ++g) {
    if (4 * g + 4 <= nfull) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) f(4 * g + jj, true);
    } else if (4 * g * B_THREADS < cnt) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) f(4 * g + jj, (4 * g + jj) * B_THREADS + tid < cnt);
    }
  }
}

// One block per bucket.  dynamic shared memory: s_c[FS_CAP + S_PAD] (64-bit composites; after the finish its words hold
// the sorted payloads [FS_CAP] | sorted s~ words [FS_CAP]) | s_cnt[S_SUB + 1]
constexpr int S_SUB = 1 << FS_LOG_S;   // sub-buckets of the counting sort
constexpr int S_PAD = 160;             // >= S_KU sentinels behind the bucket (the finish reads S_KU slots from a sub-bucket's start)
#ifndef MMBS_S_KU
#define MMBS_S_KU 8
#endif
constexpr int S_KU = MMBS_S_KU;        // slots every sample compares unconditionally
constexpr int S_DYN_SMEM = (2 * (FS_CAP + S_PAD) + S_SUB + 8) * 4;
static_assert(S_PAD >= S_KU, "sentinel padding must cover the unconditional compares");

__global__ void __launch_bounds__(B_THREADS, 2) fs_bucket_sort_kernel(
    const uint2* __restrict__ pairs, const float* __restrict__ sc_part, const uint32_t* __restrict__ cursor, int nb,
    int has_scores, const uint32_t* __restrict__ max_enc, int32_t* __restrict__ perm_out, float* __restrict__ saved_s,
    int32_t* max_count, int32_t* __restrict__ max_list, double* __restrict__ agg_val, float* __restrict__ part32,
    uint32_t* __restrict__ bucket_base, uint32_t* __restrict__ bucket_cnt, int32_t* fallback) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ double s_red[B_THREADS / 32];
  __shared__ uint32_t s_w[B_THREADS / 32];
  __shared__ uint32_t s_w2[B_THREADS / 32];
  __shared__ uint32_t s_w3[B_THREADS / 32];
  __shared__ uint32_t s_misc[4];   // 0: stop flag, 1: smallest key, 2: largest key
  // (key << 32 | index << 1 | event) in sub-bucket order: one 64-bit compare orders by (key, original index)
  unsigned long long* s_c = reinterpret_cast<unsigned long long*>(s_dyn);
  uint32_t* s_p = s_dyn;                           // after the finish: payloads in sorted order
  uint32_t* s_q = s_dyn + FS_CAP;                  // after the finish: |s~| | event << 31 in sorted order
  uint32_t* s_cnt = s_dyn + 2 * (FS_CAP + S_PAD);  // sub-bucket counters / starts, one sentinel
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = B_THREADS / 32;
  const int b = blockIdx.x;
  if (tid == 0) {   // one thread decides for the block: the flag may be raised by a running block at any time
    s_misc[0] = (*reinterpret_cast<volatile int32_t*>(fallback) != 0) ? 1u : 0u;
    s_misc[1] = 0xffffffffu;
    s_misc[2] = 0u;
  }
  for (int i = tid; i <= S_SUB; i += B_THREADS) s_cnt[i] = 0;
  const int cnt = int(min(__ldg(cursor + b), uint32_t(FS_CAP)));
  uint32_t base;
  {
    uint32_t part = 0;
    for (int q = tid; q < b; q += B_THREADS) part += min(__ldg(cursor + q), uint32_t(FS_CAP));
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_w[warp] = part;
  }
  // ---- load the bucket, its smallest and largest key
  uint32_t key[B_ITEMS], val[B_ITEMS], dr[B_ITEMS];   // dr = sub-bucket << 16 | arrival rank inside it
  const uint2* src = pairs + size_t(b) * FS_CAP;
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  for_rows(cnt, tid, [&](int j, bool ok) {
    const uint2 pr = ok ? __ldg(src + j * B_THREADS + tid) : make_uint2(0xffffffffu, 0u);
    key[j] = pr.x;
    val[j] = pr.y;
    kmin = min(kmin, pr.x);
    if (ok) kmax = max(kmax, pr.x);
  });
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  __syncthreads();   // s_misc initialised, s_cnt cleared, s_w written
  if (s_misc[0] != 0u) return;   // the pipeline already gave up
  if (lane == 0) {
    atomicMin(&s_misc[1], kmin);
    atomicMax(&s_misc[2], kmax);
  }
  base = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) base += s_w[w];
  __syncthreads();
  kmin = s_misc[1];
  // sub-bucket = floor(S_SUB * (key - kmin) / (range + 1)): monotone in the key; keys of a bucket are near-uniform
  const uint32_t range = s_misc[2] - kmin;
  const uint64_t sc64 = (uint64_t(S_SUB) << 32) / (uint64_t(range) + 1u);
  const uint32_t scale = sc64 > 0xffffffffull ? 0xffffffffu : uint32_t(sc64);
  for_rows(cnt, tid, [&](int j, bool ok) {
    if (ok) {
      const uint32_t d = __umulhi(key[j] - kmin, scale);
      dr[j] = (d << 16) | atomicAdd(&s_cnt[d], 1u);
    }
  });
  if (tid < S_PAD) s_c[cnt + tid] = ~0ull;   // sentinels: larger than every composite of the bucket
  __syncthreads();
  uint32_t cmax;   // largest sub-bucket of the block
  {
    constexpr int PER = S_SUB / B_THREADS;
    uint32_t sum = 0, big = 0, sq = 0;   // (two passes over the thread's counters instead of a register array: the
#pragma unroll                           //  48 sample registers stay live through this phase)
    for (int k = 0; k < PER; ++k) {
      const uint32_t c = s_cnt[tid * PER + k];
      sum += c;
      big = max(big, c);
      sq += c > uint32_t(S_KU) ? c * c : 0u;   // (sizes add up to <= FS_CAP: the squares to < 2^27)
    }
    big = __reduce_max_sync(0xffffffffu, big);
    sq = __reduce_add_sync(0xffffffffu, sq);
    if (lane == 0) {
      s_w2[warp] = big;
      s_w3[warp] = sq;
    }
    const uint2 sc = block_scan_u32<NW>(sum, s_w, lane, warp);   // (its barriers also publish s_w2)
    uint32_t run = sc.x - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const uint32_t c = s_cnt[tid * PER + k];
      s_cnt[tid * PER + k] = run;
      run += c;
    }
    if (tid == 0) s_cnt[S_SUB] = uint32_t(cnt);
    cmax = 0;
    uint32_t sqsum = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      cmax = max(cmax, s_w2[w]);
      sqsum += s_w3[w];
    }
    // heavy ties / a density jump inside the bucket: the finish below is quadratic in the size of a sub-bucket.  Past the
    // budget give the whole input up: the LSD pipeline redoes it.
    if (sqsum > uint32_t(FS_SQ_BUDGET)) {
      if (tid == 0) atomicOr(fallback, 2);
      return;
    }
  }
  __syncthreads();
  for_rows(cnt, tid, [&](int j, bool ok) {
    if (ok) {
      const uint32_t lo = s_cnt[dr[j] >> 16];
      const uint32_t pos = lo + (dr[j] & 0xffffu);
      s_c[pos] = (static_cast<unsigned long long>(key[j]) << 32) | __funnelshift_l(val[j], val[j], 1);
      dr[j] = lo | (dr[j] & 0xffff0000u);   // sub-bucket << 16 | its first slot (< 8192)
    }
  });
  __syncthreads();
  // ---- final rank = first slot of the sub-bucket + members that precede in the total order (key, index).  Slots behind
  // the sub-bucket hold larger composites (sub-buckets are ordered, the bucket ends in sentinels), so S_KU slots are
  // compared unconditionally; tied survival times order by original index inside the same 64-bit compare - the stable
  // order, whatever order the atomics produced.
  for_rows(cnt, tid, [&](int j, bool ok) {
    if (ok) {
      const uint32_t lo = dr[j] & 0xffffu;
      const unsigned long long me = (static_cast<unsigned long long>(key[j]) << 32) | __funnelshift_l(val[j], val[j], 1);
      const unsigned long long* row = s_c + lo;
      unsigned long long c2[S_KU];
#pragma unroll
      for (int k = 0; k < S_KU; ++k) c2[k] = row[k];
      uint32_t rank = lo;
#pragma unroll
      for (int k = 0; k < S_KU; ++k) rank += (c2[k] < me) ? 1u : 0u;
      if (cmax > uint32_t(S_KU)) {   // (block-uniform, rare) a sub-bucket over S_KU samples: the rest of its members
        const uint32_t hi = s_cnt[(dr[j] >> 16) + 1];
#pragma unroll 4
        for (uint32_t q = lo + S_KU; q < hi; ++q) rank += (s_c[q] < me) ? 1u : 0u;
      }
      dr[j] = rank;
    }
  });
  uint32_t* po = reinterpret_cast<uint32_t*>(perm_out) + base + tid;
  if (!has_scores) {   // mmbs_risk_order: the permutation only
    __syncthreads();
    for_rows(cnt, tid, [&](int j, bool ok) {
      if (ok) s_p[dr[j]] = val[j];
    });
    __syncthreads();
    for_rows(cnt, tid, [&](int j, bool ok) {
      if (ok) po[j * B_THREADS] = s_p[j * B_THREADS + tid];
    });
    return;
  }
  // ---- the scores came along with the partition (same slot as the pair): s~ = score - max, kept as |s~| with the event
  // bit in the sign position (s~ <= 0 always); both words move to their sorted slot
  const float* scp = sc_part + size_t(b) * FS_CAP + tid;
  for_rows(cnt, tid, [&](int j, bool ok) {   // (key is dead: its registers carry the scores)
    key[j] = ok ? __float_as_uint(__ldg(scp + j * B_THREADS)) : 0u;
  });
  const float smax = float_order_dec(__ldg(max_enc));
  __syncthreads();   // every thread is done with the composites (and with the sub-bucket starts)
  for_rows(cnt, tid, [&](int j, bool ok) {
    if (ok) {
      const float sv = __uint_as_float(key[j]) - smax;
      s_p[dr[j]] = val[j];
      s_q[dr[j]] = (__float_as_uint(sv) & 0x7fffffffu) | (val[j] & 0x80000000u);
      if (sv == 0.f) {   // an argmax position (gradient through max(scores) in the backward pass)
        const int pos = atomicAdd(max_count, 1);
        if (pos < FS_MAX_LIST) max_list[pos] = int32_t(val[j] & 0x7fffffffu);
      }
    }
  });
  __syncthreads();
  uint32_t* so = reinterpret_cast<uint32_t*>(saved_s) + base + tid;
  for_rows(cnt, tid, [&](int j, bool ok) {
    if (ok) {
      po[j * B_THREADS] = s_p[j * B_THREADS + tid];
      so[j * B_THREADS] = s_q[j * B_THREADS + tid];
    }
  });
  // sums of exp(s~): thread t adds up the 16 consecutive sorted positions 16 t .. 16 t + 15 of the bucket, two adjacent
  // threads make one of the 32-position partial sums the row kernels build their offsets from
  float esum = 0.f;
  if (16 * tid < cnt) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 x = *reinterpret_cast<const uint4*>(s_q + 16 * tid + 4 * q);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        esum += (16 * tid + 4 * q + k < cnt) ? fs_exp(__uint_as_float(xs[k] | 0x80000000u)) : 0.f;   // s~ = -|s~|
    }
  }
  const float pair = esum + __shfl_xor_sync(0xffffffffu, esum, 1);
  if ((tid & 1) == 0) part32[size_t(b) * FS_PART + (tid >> 1)] = pair;
  const double tot = block_sum_f64<NW>(double(esum), s_red, lane, warp);
  if (tid == 0) {
    agg_val[b] = tot;
    bucket_base[b] = base;
    bucket_cnt[b] = uint32_t(cnt);
  }
}

// ------------------------------------------------------------------------------------------ per-row loss / gradient
// Both kernels work in WARP ROWS: warp `wg` (0..15) of bucket b owns the 512 sorted positions a0 + 512 wg .. +512, with
// a0 = the bucket's first position rounded down to a multiple of 4; lane l owns 16 consecutive positions, fetched as four
// 16-byte loads (positions in front of the bucket / behind it are masked).  A warp finds the sum of exp(s~) in front of its
// row from the bucket prefix and the 32-position partial sums the sort kernel left (part32), scans inside the warp with
// shuffles, and leaves per-row partial sums: no block-wide scan, no barrier between the phases of different warps.  (One
// block per bucket with block-wide scans spent most of its time in barriers: 50 / 134 us against ... for this layout.)
constexpr int R_THREADS = 128;
constexpr int R_WARPS = R_THREADS / 32;
constexpr int R_SPAN = 512;                    // positions per warp row
constexpr int R_ROWS = FS_CAP / R_SPAN;        // 16 warp rows per bucket
constexpr int R_BLOCKS = R_ROWS / R_WARPS;     // 4 blocks per bucket
constexpr int L_ITEMS = R_SPAN / 32;           // 16
static_assert(L_ITEMS % 4 == 0 && R_ROWS == 16 && FS_PART == FS_CAP / 32, "row geometry");
struct BucketSlice {
  uint32_t a0;     // first position covered by the bucket's rows: bucket base rounded down to a multiple of 4
  int lead;        // masked positions in front of the bucket (0..3)
  int end;         // lead + samples of the bucket: row-local positions [lead, end) are real
};
__device__ __forceinline__ BucketSlice bucket_slice(const uint32_t* __restrict__ bucket_base,
                                                    const uint32_t* __restrict__ bucket_cnt, int b) {
  const uint32_t base = __ldg(bucket_base + b);
  BucketSlice s;
  s.a0 = base & ~3u;
  s.lead = int(base - s.a0);
  s.end = s.lead + int(__ldg(bucket_cnt + b));
  return s;
}
// L_ITEMS consecutive words of the lane (zeros where no 16-byte group of the bucket lies)
__device__ __forceinline__ void load_items(const uint32_t* __restrict__ p, int first, int end, uint32_t (&v)[L_ITEMS]) {
#pragma unroll
  for (int q = 0; q < L_ITEMS / 4; ++q) {
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (first + 4 * q < end) x = __ldg(reinterpret_cast<const uint4*>(p + first + 4 * q));
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}

// A lane is INTERIOR when all L_ITEMS of its positions are samples of the bucket (every lane but the one or two at the
// bucket's ends and the idle ones behind it): its loops carry no per-sample predicate.  f(j, ok) is instantiated once with
// ok == true as a constant and once with the real test.
template <typename F>
__device__ __forceinline__ void for_items(bool interior, int first, int lead, int end, F&& f) {
  if (interior) {
#pragma unroll
    for (int j = 0; j < L_ITEMS; ++j) f(j, true);
  } else if (first < end) {
#pragma unroll
    for (int j = 0; j < L_ITEMS; ++j) f(j, first + j >= lead && first + j < end);
  }
}
template <typename F>
__device__ __forceinline__ void for_items_rev(bool interior, int first, int lead, int end, F&& f) {
  if (interior) {
#pragma unroll
    for (int j = L_ITEMS - 1; j >= 0; --j) f(j, true);
  } else if (first < end) {
#pragma unroll
    for (int j = L_ITEMS - 1; j >= 0; --j) f(j, first + j >= lead && first + j < end);
  }
}
// exclusive scans of one double per lane: sum over the lower lanes / over the higher lanes
__device__ __forceinline__ double warp_excl_fwd(double v, int lane) {
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  return incl - v;
}
__device__ __forceinline__ double warp_excl_rev(double v, int lane) {
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  return incl - v;
}
// sum of a per-thread double over the block's R_WARPS warps (one barrier; every thread of the block must call)
// sum of the lane's first `lead` running values = c[lead - 1] (lane 0 of a row: the samples in front of position 512 wg)
__device__ __forceinline__ float lead_sum(const float (&c)[L_ITEMS], int lead) {
  return lead == 1 ? c[0] : (lead == 2 ? c[1] : (lead == 3 ? c[2] : 0.f));
}
__device__ __forceinline__ double row_block_sum(double v, double* s_red, int lane, int warp) {
  v = warp_sum(v);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < R_WARPS; ++w) t += s_red[w];
  return t;
}
// sum of exp(s~) over the first 512 wg positions of the bucket: 16 wg partial sums of 32 positions each.  (The row itself
// starts `lead` positions earlier, at the 16-byte aligned address: the caller takes those samples' terms off again.)
__device__ __forceinline__ double row_exp_offset(const float* __restrict__ part32, int b, int wg, int lane) {
  const float* p = part32 + size_t(b) * FS_PART;
  double s = 0.0;
  for (int c = lane; c < wg * (R_SPAN / 32); c += 32) s += double(__ldg(p + c));
  return warp_sum(s);
}

__global__ void __launch_bounds__(R_THREADS, 8) fs_row_loss_kernel(
    const float* __restrict__ saved_s, const int32_t* __restrict__ perm, const float* __restrict__ status,
    const uint32_t* __restrict__ bucket_base, const uint32_t* __restrict__ bucket_cnt, const double* __restrict__ exp_prefix,
    const float* __restrict__ part32, double* __restrict__ row_loss, double* __restrict__ row_w, int32_t* nan_flag,
    const int32_t* __restrict__ nonbinary_flag, const int32_t* __restrict__ fallback) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*fallback != 0) return;   // final by now: the kernels that raise it have finished
  const int b = blockIdx.x / R_BLOCKS, wg = (blockIdx.x % R_BLOCKS) * R_WARPS + warp;
  const BucketSlice sl = bucket_slice(bucket_base, bucket_cnt, b);
  const int first = wg * R_SPAN + lane * L_ITEMS;
  const bool active = wg * R_SPAN < sl.end;   // warp-uniform: the row holds samples
  uint32_t enc[L_ITEMS];
  if (active) load_items(reinterpret_cast<const uint32_t*>(saved_s) + sl.a0, first, sl.end, enc);
  const double P = __ldg(exp_prefix + b);   // sum of exp(s~) over all earlier buckets (fs_prefix_kernel)
  double tl = 0.0, tw = 0.0;
  if (active) {
    const bool interior = first >= sl.lead && first + L_ITEMS <= sl.end;
    const double rowoff = P + row_exp_offset(part32, b, wg, lane);
    float c[L_ITEMS];
    float run = 0.f;
#pragma unroll
    for (int j = 0; j < L_ITEMS; ++j) c[j] = 0.f;
    for_items(interior, first, sl.lead, sl.end, [&](int j, bool ok) {
      run += ok ? fs_exp(__uint_as_float(enc[j] | 0x80000000u)) : 0.f;   // s~ = -|s~|
      c[j] = run;
    });
    const float head = __shfl_sync(0xffffffffu, lead_sum(c, sl.lead), 0);
    const double off = (rowoff - double(head)) + warp_excl_fwd(double(run), lane);
    // C + eps = off + c[j] + eps, rounded to fp32 once: off as a (hi, lo) float pair, eps folded into lo (2 FADD per sample
    // instead of fp64 converts)
    const float off_hi = float(off), off_lo = float(off - double(off_hi)) + FS_EPS;
    float lsum = 0.f, ws = 0.f;
    if (*nonbinary_flag == 0) {
      for_items(interior, first, sl.lead, sl.end, [&](int j, bool ok) {
        const bool event = ok && (enc[j] >> 31);
        const float den = (off_hi + c[j]) + off_lo;                                             // cumsum + eps (models.py:104)
        const float term = fmaf(fs_lg2(den), 0.6931471805599453f, __uint_as_float(enc[j] & 0x7fffffffu));   // log(.) - s~
        lsum += event ? term : 0.f;                                                             // (models.py:104-105)
        ws += event ? fs_rcp(den) : 0.f;
      });
    } else if (first < sl.end) {   // general status weights: gathered through the permutation
      uint32_t pv[L_ITEMS];
      load_items(reinterpret_cast<const uint32_t*>(perm) + sl.a0, first, sl.end, pv);
#pragma unroll
      for (int j = 0; j < L_ITEMS; ++j) {
        const bool ok = first + j >= sl.lead && first + j < sl.end;
        if (ok) {
          const float dl = __ldg(status + (pv[j] & 0x7fffffffu));
          const float den = (off_hi + c[j]) + off_lo;
          const float term = -(__uint_as_float(enc[j] | 0x80000000u) - fs_log(den)) * dl;
          lsum += term;
          ws += dl * fs_rcp(den);
        }
      }
    }
    const unsigned any_bad = __ballot_sync(0xffffffffu, lsum != lsum);   // a NaN term makes the lane's sum NaN
    if (lane == 0 && any_bad) atomicOr(nan_flag, 1);
    tl = warp_sum(double(lsum));
    tw = warp_sum(double(ws));
  }
  if (lane == 0) {   // fs_loss_finalize_kernel adds the rows up
    row_loss[b * R_ROWS + wg] = tl;
    row_w[b * R_ROWS + wg] = tw;
  }
}

// the per-sample part of the backward pass; GATHER: general (non-binary) status weights, read through the permutation
template <bool GATHER>
__device__ __forceinline__ double row_backward_body(const BucketSlice sl, int first, const int32_t* __restrict__ perm,
                                                    const float* __restrict__ saved_s, const float* __restrict__ status,
                                                    double rowoff, double S_later, float scale, float* grad_scores,
                                                    int lane) {
  const bool interior = first >= sl.lead && first + L_ITEMS <= sl.end;
  float e[L_ITEMS], cw[L_ITEMS], dl[GATHER ? L_ITEMS : 1];
  uint32_t evm = 0;
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < L_ITEMS; ++j) e[j] = cw[j] = 0.f;
  {
    uint32_t enc[L_ITEMS];
    load_items(reinterpret_cast<const uint32_t*>(saved_s) + sl.a0, first, sl.end, enc);
    for_items(interior, first, sl.lead, sl.end, [&](int j, bool ok) {
      e[j] = ok ? fs_exp(__uint_as_float(enc[j] | 0x80000000u)) : 0.f;   // s~ = -|s~|; no sample: 0
      evm |= (ok ? (enc[j] >> 31) : 0u) << j;
      run += e[j];
      cw[j] = run;
    });
  }
  if (GATHER) {
    uint32_t pv[L_ITEMS];
    load_items(reinterpret_cast<const uint32_t*>(perm) + sl.a0, first, sl.end, pv);
#pragma unroll
    for (int j = 0; j < L_ITEMS; ++j) {
      const bool ok = first + j >= sl.lead && first + j < sl.end;
      dl[GATHER ? j : 0] = ok ? __ldg(status + (pv[j] & 0x7fffffffu)) : 0.f;
    }
  }
  const float head = __shfl_sync(0xffffffffu, lead_sum(cw, sl.lead), 0);
  const double off = (rowoff - double(head)) + warp_excl_fwd(double(run), lane);
  const float off_hi = float(off), off_lo = float(off - double(off_hi)) + FS_EPS;   // the same C + eps as the forward pass
  float wrun = 0.f;
  for_items_rev(interior, first, sl.lead, sl.end, [&](int j, bool ok) {
    const float den = (off_hi + cw[j]) + off_lo;
    float w;
    if (GATHER) w = dl[GATHER ? j : 0] * fs_rcp(den);               // 0 where there is no sample
    else w = ((evm >> j) & 1u) ? fs_rcp(den) : 0.f;
    wrun += w;
    cw[j] = wrun;   // inclusive suffix of w inside the lane
  });
  const double soff = S_later + warp_excl_rev(double(wrun), lane);
  const float soff_hi = float(soff), soff_lo = float(soff - double(soff_hi));
  uint32_t pv[L_ITEMS];
  load_items(reinterpret_cast<const uint32_t*>(perm) + sl.a0, first, sl.end, pv);
  float gs = 0.f;
  for_items(interior, first, sl.lead, sl.end, [&](int j, bool ok) {
    if (ok) {
      const float d = GATHER ? dl[GATHER ? j : 0] : float((evm >> j) & 1u);
      const float W = (soff_hi + cw[j]) + soff_lo;   // sum of w over this and all later positions
      const float g = -(d - e[j] * W) * scale;
      grad_scores[pv[j] & 0x7fffffffu] = g;          // un-permute
      gs += g;
    }
  });
  return warp_sum(double(gs));
}

__device__ __noinline__ double row_backward_gather(const BucketSlice sl, int first, const int32_t* __restrict__ perm,
                                                   const float* __restrict__ saved_s, const float* __restrict__ status,
                                                   double rowoff, double S_later, float scale, float* grad_scores, int lane) {
  return row_backward_body<true>(sl, first, perm, saved_s, status, rowoff, S_later, scale, grad_scores, lane);
}

__global__ void __launch_bounds__(R_THREADS, 8) fs_row_backward_kernel(
    const int32_t* __restrict__ perm, const float* __restrict__ saved_s, const float* __restrict__ status,
    const float* __restrict__ grad_loss, int64_t n, const uint32_t* __restrict__ bucket_base,
    const uint32_t* __restrict__ bucket_cnt, const double* __restrict__ exp_prefix, const float* __restrict__ part32,
    const double* __restrict__ wsuffix, const double* __restrict__ row_w, double* __restrict__ row_g,
    const int32_t* __restrict__ nonbinary_flag, float* grad_scores, const int32_t* __restrict__ fallback) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*fallback != 0) return;
  const int b = blockIdx.x / R_BLOCKS, wg = (blockIdx.x % R_BLOCKS) * R_WARPS + warp;
  const BucketSlice sl = bucket_slice(bucket_base, bucket_cnt, b);
  const int first = wg * R_SPAN + lane * L_ITEMS;
  double tg = 0.0;
  if (wg * R_SPAN < sl.end) {   // warp-uniform: the row holds samples
    // sum of w over all later buckets (fs_loss_finalize_kernel) and over the later rows of this bucket
    const double S_later = __ldg(wsuffix + b) +
                           warp_sum((lane > wg && lane < R_ROWS) ? __ldg(row_w + b * R_ROWS + lane) : 0.0);
    const double rowoff = __ldg(exp_prefix + b) + row_exp_offset(part32, b, wg, lane);
    const float scale = float(double(grad_loss[0]) / double(n));
    if (*nonbinary_flag != 0)   // (rare, and out of line: its extra live array must not cost the common path registers)
      tg = row_backward_gather(sl, first, perm, saved_s, status, rowoff, S_later, scale, grad_scores, lane);
    else
      tg = row_backward_body<false>(sl, first, perm, saved_s, status, rowoff, S_later, scale, grad_scores, lane);
  }
  if (lane == 0) row_g[b * R_ROWS + wg] = tg;   // fs_backward_finalize_kernel adds the rows up
}

// ------------------------------------------------------------------------------------------ one-block scans between the passes
// (A grid-wide "last block finishes the job" needs a fence + counter atomic in every warp row; with sixteen scattered
// stores in flight per lane those fences were a quarter of the backward kernel's stall samples.  Kernel boundaries order
// the passes instead: the prefix kernel below, and fs_loss_finalize_block / fs_backward_finalize_block (cox_sort.cuh),
// which run inside the one-block dispatch kernels of cox.cu that follow either pass anyway.)
constexpr int F_THREADS = FS_FINAL_THREADS;
constexpr int F_PER = FS_MAX_BUCKETS / F_THREADS;   // 4 consecutive buckets per thread

// exclusive prefix of the buckets' sums of exp(s~)
__global__ void __launch_bounds__(F_THREADS, 1) fs_prefix_kernel(const double* __restrict__ agg_val, int nb,
                                                                 double* __restrict__ exp_prefix,
                                                                 const int32_t* __restrict__ fallback) {
  __shared__ double s_red[F_THREADS / 32];
  if (*fallback != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[F_PER], sum = 0.0;
#pragma unroll
  for (int k = 0; k < F_PER; ++k) {
    const int q = tid * F_PER + k;
    v[k] = q < nb ? __ldg(agg_val + q) : 0.0;
    sum += v[k];
  }
  double total;
  double run = block_scan_f64<F_THREADS / 32, false>(sum, s_red, lane, warp, &total) - sum;
#pragma unroll
  for (int k = 0; k < F_PER; ++k) {
    const int q = tid * F_PER + k;
    if (q < nb) exp_prefix[q] = run;
    run += v[k];
  }
}

// ------------------------------------------------------------------------------------------ host side
int fs_sample_shift(int64_t n) {
  static const int forced = []() {
    const char* e = getenv("MMBS_COX_HIST_SAMPLE");
    return (e && e[0] >= '0' && e[0] <= '5') ? int(e[0] - '0') : -1;
  }();
  if (forced >= 0) return forced;
  return n >= (int64_t(1) << 21) ? 3 : 0;   // >= 2 M samples: one chunk in 8 (>= 256 K samples counted)
}

int fs_target() {
  static const int t = []() {
    const char* e = getenv("MMBS_COX_TARGET");
    const int v = e ? atoi(e) : 0;
    return (v >= 1024 && v <= FS_TARGET) ? v : FS_TARGET;   // never above FS_TARGET: FS_MAX_N assumes it
  }();
  return t;
}

static int fs_configure() {
  static PerDeviceOnce configured;
  if (configured.first()) {
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (2 * FS_MAX_BUCKETS + 3 * P_TILE) * 4));
    MMBS_CUDA_TRY(cudaFuncSetAttribute(fs_bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_DYN_SMEM));
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
  // a sampling histogram leaves max(scores) / the NaN flag to the partition pass
  const bool sampled = p.sample_shift > 0;
  const int64_t groups = ceil_div(ceil_div(n, 1024), int64_t(1) << p.sample_shift);
  const int hist_grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(groups, 2), int64_t(sms) * 6)));
  fs_hist_kernel<<<hist_grid, FS_HIST_THREADS, 0, stream>>>(times, n, p.sample_shift, p.nb, sampled ? nullptr : scores,
                                                            w.hist12, w.kext, max_enc, nan_flag, w.counters + 0, w.lut,
                                                            w.edge);
  MMBS_LAUNCH_CHECK();
  const int64_t tiles = ceil_div(n, P_TILE);
  const int nbp = int(ceil_div(p.nb, P_THREADS)) * P_THREADS;
  const int p_smem = (2 * nbp + 3 * P_TILE) * 4;
  fs_partition_kernel<<<unsigned(tiles), P_THREADS, p_smem, stream>>>(times, status, scores, sampled ? 1 : 0, max_enc,
                                                                     nan_flag, n, w.lut, w.edge, p.nb, w.cursor, w.pairs,
                                                                     w.sc_part, nonbinary_flag, w.fallback);
  MMBS_LAUNCH_CHECK();
  fs_bucket_sort_kernel<<<p.nb, B_THREADS, S_DYN_SMEM, stream>>>(w.pairs, w.sc_part, w.cursor, p.nb, scores != nullptr,
                                                                max_enc, perm_out, saved_s, max_count, max_list, w.agg_val,
                                                                w.part32, w.bucket_base, w.bucket_cnt, w.fallback);
  MMBS_LAUNCH_CHECK();
  if (scores == nullptr) return MMBS_OK;   // mmbs_risk_order
  fs_prefix_kernel<<<1, F_THREADS, 0, stream>>>(w.agg_val, p.nb, w.exp_prefix, w.fallback);
  MMBS_LAUNCH_CHECK();
  fs_row_loss_kernel<<<p.nb * R_BLOCKS, R_THREADS, 0, stream>>>(saved_s, perm_out, status, w.bucket_base, w.bucket_cnt,
                                                               w.exp_prefix, w.part32, w.row_loss, w.row_w, nan_flag,
                                                               nonbinary_flag, w.fallback);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;   // the caller's dispatch kernel (cox.cu) finishes the pass: fs_loss_finalize_block
}

int fs_backward_enqueue(const float* status, const int32_t* perm, const float* saved_s, const float* grad_loss, int64_t n,
                        const FastSortWs& w, const int32_t* nonbinary_flag, const int32_t* max_count,
                        const int32_t* max_list, double* gsum_total, float* grad_scores, cudaStream_t stream) {
  const FsPlan p = fs_plan(n);
  fs_row_backward_kernel<<<p.nb * R_BLOCKS, R_THREADS, 0, stream>>>(perm, saved_s, status, grad_loss, n, w.bucket_base,
                                                                   w.bucket_cnt, w.exp_prefix, w.part32, w.wsum, w.row_w,
                                                                   w.row_g, nonbinary_flag, grad_scores, w.fallback);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;   // the caller's dispatch kernel (cox.cu) finishes the pass: fs_backward_finalize_block
}

}  // namespace mmbs
