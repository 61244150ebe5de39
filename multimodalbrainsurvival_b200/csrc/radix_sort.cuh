// Stable LSD radix sort of (u32 key, u32 index) pairs, Onesweep style: one global
// histogram pass, then one read+write pass per 8-bit digit in which every tile
// resolves its global digit offsets by decoupled look-back over its predecessors
// (no separate scan kernel between passes, 16 B/pair/pass of HBM traffic).
//
// Used by the Cox loss (key = order-preserving image of -time, reference
// torch.sort(-times) at /root/reference/1_HistoPathology/models.py:99) and by the
// per-patient aggregation (key = segment id).
#pragma once

#include "common.cuh"

namespace mmbs {

// 8192 pairs per tile: with 256 digits a tile leaves ~32 keys = one full 128-byte line per digit run
// (4096-pair tiles wrote half lines and ran ~1.6x slower per pass)
#ifndef RS_THREADS_DEF
#define RS_THREADS_DEF 512
#endif
constexpr int RS_THREADS = RS_THREADS_DEF;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_BLOCKS_PER_SM = 1024 / RS_THREADS;   // 64 registers per thread
constexpr int RS_DYN_SMEM = 2 * RS_TILE * 4;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;
constexpr int RS_LOOKBACK = 8;
#ifndef RS_PERSISTENT
#define RS_PERSISTENT 0
#endif
#ifndef RS_UNROLL_OUT
#define RS_UNROLL_OUT 4
#endif
#ifndef RS_SPIN_NS
#define RS_SPIN_NS 200
#endif  // predecessor tiles inspected per look-back round trip
constexpr uint32_t RS_FLAG_AGG = 1u << 30;
constexpr uint32_t RS_FLAG_INCL = 2u << 30;
constexpr uint32_t RS_FLAG_MASK = 3u << 30;
constexpr uint32_t RS_VALUE_MASK = (1u << 30) - 1;
constexpr int64_t RS_MAX_N = (int64_t(1) << 30) - 1;

enum KeyKind : int { KEY_NEG_TIME_F32 = 0, KEY_U32 = 1 };

struct SortWorkspace {
  uint32_t* keys_a;
  uint32_t* keys_b;
  uint32_t* vals_a;
  uint32_t* vals_b;
  uint32_t* hist;        // [4][256]    (zeroed by caller)
  uint32_t* digit_base;  // [4][256]
  uint32_t* counters;    // [4]         (zeroed by caller)
  uint32_t* lookback;    // [4][tiles][256] (zeroed by caller)
};

static inline int64_t rs_tiles(int64_t n) { return n > 0 ? (n + RS_TILE - 1) / RS_TILE : 1; }

// Enqueue the sort.  `src` holds n floats (times) or n u32 keys.  The histogram
// must already be in ws.hist (see cox.cu / segmean.cu: the histogram kernel is
// fused with other per-element work there).  Writes the sorted original indices
// to perm_out.  num_passes in [1,4]: number of low bytes that can differ.
// When `status` is given, bit 31 of every output word carries (status[idx] != 0) and
// *nonbinary_flag is set if some status value is neither 0 nor 1.
// `enable` (device flag, optional): every kernel returns at once when *enable == 0 - the sort is then enqueued
// unconditionally behind the two-pass sort of cox_sort.cu and only runs when that one gave up (no host sync).
int rs_sort_enqueue(const void* src, KeyKind kind, int64_t n, int num_passes,
                    const SortWorkspace& ws, int32_t* perm_out, cudaStream_t stream,
                    const float* status = nullptr, int32_t* nonbinary_flag = nullptr,
                    const int32_t* enable = nullptr);

// Plain histogram (+ optional fused max / NaN flag over `scores`).
int rs_histogram_enqueue(const void* src, KeyKind kind, int64_t n, int num_passes, uint32_t* hist,
                         uint32_t* digit_base, const float* scores, uint32_t* max_enc,
                         int32_t* nan_flag, cudaStream_t stream, const int32_t* enable = nullptr);

// launch geometry / one-time kernel configuration, for callers that enqueue the sort from the device (below)
int rs_configure();
int rs_hist_grid(int64_t n);
unsigned rs_sort_grid(int64_t n);

#ifdef __CUDACC__
// Device-side twin of rs_histogram_enqueue + rs_sort_enqueue (tail launches behind the calling grid); the caller has
// cleared ws.hist / ws.counters / ws.lookback and run rs_configure() on the host.  Needs -rdc=true.
__device__ void rs_sort_tail_launch(const void* src, int kind, int64_t n, int num_passes, SortWorkspace ws,
                                    int32_t* perm_out, const float* status, int32_t* nonbinary_flag, int hist_grid,
                                    unsigned sort_grid, int64_t tiles);

__device__ __forceinline__ uint32_t rs_load_key(const void* src, int kind, int64_t i) {
  if (kind == KEY_NEG_TIME_F32) return time_key(__ldg(static_cast<const float*>(src) + i));
  return __ldg(static_cast<const uint32_t*>(src) + i);
}
// order-preserving u32 image of a float (for atomicMax over floats)
__device__ __forceinline__ uint32_t float_order_enc(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_order_dec(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// L2 evict_last loads: keep a gathered-from array resident while streams pass through L2
__device__ __forceinline__ uint64_t make_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float ld_f32_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ld_f32x4_hint(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

#endif

}  // namespace mmbs
