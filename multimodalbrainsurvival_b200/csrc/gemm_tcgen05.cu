// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a).
//
// One warp-specialised kernel serves every conv of ResNet.forward_extract
// (/root/reference/5_JointFusion/resnet.py:151-165, Bottleneck.forward :70-90) in eval and training mode, its
// layer4 backward, and every nn.Linear of the RNA / fusion MLPs
// (2_GeneExpression/1_GeneExpress_train.py:247-257 ...), forward and backward:
//
//   out[m, n] = act( scale[n] * sum_k A[m, k] * W[n, k] + shift[n] (+ residual[m, n]) )
//
// * A rows are output pixels.  An M tile is a BOX of 128 output pixels (tw x th x tn
//   over W, H, N), so the A operand of any filter tap is a plain 4-D TMA box of the
//   NHWC input shifted by the tap offset (out-of-range rows/cols zero-filled by TMA =
//   the conv padding).  Stride-2 convs read one of four "parity" views of the input
//   (base pointer offset + doubled strides), so they are shifted boxes as well.
// * W is [Cout][kh][kw][Cin] bf16: K-major, one 2-D TMA box per (tap, 64-channel chunk); small weight matrices
//   stay resident in shared memory (RES_BYTES > 0), the stem and 64-channel 3x3 convs use halo boxes.
// * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (accumulators ping-pong in TMEM), warps 2-9 = epilogue
//   (tcgen05.ld -> scale/shift (+residual) (+ReLU) -> bf16 staging -> TMA store, or fp32 direct stores),
//   warp 10 = output-DMA warp of the 256-wide path (per-block TMA stores, residual loads, staging recycling).
// * smem ring of STAGES x (A 16 KB + B N_TILE*128 B), 128-byte swizzle end to end.
// * Operand majors (ConvParams::mn_major): K-major A and B (forward convs / linear layers); MN-major B = the data
//   gradient reading the FORWARD weights (flags bit 1, mmbs_linear_nn_plan_create); MN-major A and B = weight
//   gradients straight from the row-major activations / gradients (mmbs_linear_tn_plan_create), with 4-D boxes of
//   64 images at one output pixel as the K chunk for 3x3 / strided convs (mmbs_conv_wgrad_plan_create).
// * Training-mode BatchNorm: the epilogue also accumulates per-channel sum / sum of squares (ConvParams::stats).
// * Split-K (fp32 reductions into a cleared output) for weight-gradient shaped problems.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "common.cuh"
#include "tc_common.cuh"

namespace mmbs {

constexpr int GM_TILE_M = 128;
constexpr int GM_CHUNK_K = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int GM_A_BYTES = GM_TILE_M * GM_CHUNK_K * 2;  // 16 KB
constexpr int GM_EPI_WARPS = 8;
constexpr int GM_EPI_THREADS = GM_EPI_WARPS * 32;    // 256
constexpr int GM_DMA_WARP = 2 + GM_EPI_WARPS;        // warp 10: output stores / residual loads of the 256-wide path
constexpr int GM_THREADS = 64 + GM_EPI_THREADS + 32;  // producer warp, MMA warp, 8 epilogue warps, output-DMA warp
constexpr int GM_MAX_TAPS = 16;
constexpr int GM_OUT_BLK_BYTES = GM_TILE_M * 128;    // one 64-channel column block of a bf16 tile
constexpr int GM_RES64_A_STAGE = 24 * 1024;          // halo box: up to (th + 3) * tw = 192 rows of 128 B
constexpr int GM_RES64_BYTES = 72 * 1024;            // 64 x 576 weights (layer1 3x3) resident
constexpr int GM_RES128_BYTES = 64 * 1024;           // 256 x 128 / 128 x 256 weights resident

// Division by a runtime constant as multiply-high + shift (tile index decomposition runs per tile in
// every epilogue thread; hardware integer division costs ~25 instructions each).
struct FastDiv {
  uint32_t d, mul, shr;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1u << l) < d) ++l;
  f.shr = l;
  f.mul = uint32_t(((uint64_t(1) << 32) * ((uint64_t(1) << l) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) {
  if (f.d == 1) return x;
  const uint32_t t = __umulhi(x, f.mul);
  return (t + ((x - t) >> 1)) >> (f.shr - 1);
}

struct ConvParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap out_map;   // bf16 output, box (64, tw, th, tn)      (TMA-store path)
  CUtensorMap res_map;   // bf16 residual, same geometry           (TMA-load into the staging tile)
  int32_t tw, th, tn;
  int32_t tiles_w, tiles_h, tiles_n;
  int32_t out_w, out_h, batch;
  int32_t c_out;
  int32_t num_taps, k_chunks;
  int32_t relu, out_f32;
  int32_t total_tiles;
  FastDiv fd_ntiles, fd_tw, fd_twh;   // n_tiles, tiles_w, tiles_w * tiles_h
  int32_t tw_shift, twh_shift;        // log2(tw), log2(tw * th)  (the pixel box sides are powers of two)
  int32_t taps_per_stage;   // MMA tap groups fed by one A stage (halo mode: rows shifted by tap_row_bytes)
  int32_t tap_row_bytes;    // tw * 128: shared-memory distance between the A views of consecutive row taps
  int32_t a_tx_bytes;       // bytes of one A stage load
  int32_t res_boxes;        // resident-weights mode: number of (N_TILE x 64) weight boxes loaded once
  int32_t kb_per_tile;      // weight boxes per n-tile
  // split-K (weight-gradient GEMMs: few output tiles, very long K): tile = ks * mn_tiles + (m, n) tile; split ks
  // covers K chunks [ks * chunks_per_split, ...) and adds its partial sums to the pre-zeroed fp32 output
  int32_t k_split, chunks_per_split, mn_tiles;
  int32_t cluster2;   // CTA pairs: the two CTAs of a cluster work on adjacent pixel tiles of the same column tile and
                      // each fetches half of every weight box for both (TMA multicast)
  int32_t mn_major;   // bit0: A is MN-major (row-major [K, M]); bit1: B is MN-major (row-major [K, N]).  3 = TN GEMM (weight
                      // gradients from NHWC tensors); 2 = data gradient reading the FORWARD conv's packed weights
  // implicit weight-gradient GEMM of a convolution (mn_major == 3, wg_conv): K chunk = 64 images at one output pixel
  int32_t wg_conv, wg_nb, wg_ow, wg_stride, wg_ksize, wg_cin, wg_cin_tiles;
  int32_t out_col_stride;   // fp32 direct-store epilogue: distance between consecutive output columns (OIHW: k*k)
  int32_t b_tap_stride;   // mode 2: columns per tap of the forward weight matrix (= forward c_in = this conv's c_out)
  FastDiv fd_mn;
  uint32_t idesc;
  int8_t tap_map[GM_MAX_TAPS];
  int8_t tap_dw[GM_MAX_TAPS];
  int8_t tap_dh[GM_MAX_TAPS];
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
  float* stats;   // training-mode BatchNorm: [2][c_out] per-channel sum / sum of squares of the bf16 output
  // dropout fused into the epilogue of a linear layer (after bias / ReLU; nn.Dropout in front of the NEXT Linear of the
  // reference MLPs, 1_GeneExpress_train.py:247-257): keep-mask of common.cuh dropout_keep8 at (row, column / 8)
  uint32_t drop_thresh;   // 0: no dropout; otherwise round(p * 65536)
  uint32_t drop_tag;
  float drop_scale;       // 1 / (1 - p)
  uint64_t drop_seed;
};

// Shared-memory carve-up (offsets from a 1024-B aligned base):
//   [ring: STAGES x (A 16 KB | B N_TILE*128 B)] [staging: NSTG x N_TILE/64 x 16 KB] [barriers] [tmem slot] [scale|shift]
// RES_BYTES > 0 selects the weights-resident variant: the whole weight matrix is loaded into smem once
// per CTA and the ring holds A stages only (A_STAGE bytes each, enough for a halo box).
template <int N_TILE, int STAGES, int NSTG, int A_STAGE, int RES_BYTES>
struct GemmSmem {
  static constexpr int B_BYTES = N_TILE * GM_CHUNK_K * 2;
  static constexpr int STAGE_BYTES = RES_BYTES > 0 ? A_STAGE : (GM_A_BYTES + B_BYTES);
  static constexpr int OUT_BLKS = (N_TILE + 63) / 64;
  static constexpr int RES_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFFSET = RES_OFFSET + RES_BYTES;
  static constexpr int STAGING_BYTES = OUT_BLKS * GM_OUT_BLK_BYTES;    // one staging tile
  static constexpr int BAR_OFFSET = STAGING_OFFSET + NSTG * STAGING_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 15;  // full[S] empty[S] tfull[2] tempty[2] res[2] wres blk_in[4] blk_out[4]
  static constexpr int TMEM_SLOT_OFFSET = BAR_OFFSET + NUM_BARS * 8;
  static constexpr int SCALE_OFFSET = (TMEM_SLOT_OFFSET + 4 + 15) / 16 * 16;
  static constexpr int TOTAL = SCALE_OFFSET + NSTG * 2 * N_TILE * 4;   // scale|shift per epilogue group
  static constexpr int DYNAMIC = TOTAL + 1024;  // slack for the 1024-B alignment of the ring
  static constexpr int TMEM_COLS = 2 * N_TILE;  // double-buffered fp32 accumulator
  static_assert(STAGE_BYTES % 1024 == 0 && RES_BYTES % 1024 == 0, "swizzle atoms need 1024-B alignment");
  static_assert(DYNAMIC <= 227 * 1024, "shared memory budget exceeded");
};

struct TileCoord {
  int nt, w0, h0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int tile, int n_tiles) {
  TileCoord c;
  // n-tile fastest: CTAs that share an A tile run back to back (L2 reuse)
  uint32_t mt = fdiv(uint32_t(tile), p.fd_ntiles);
  c.nt = tile - int(mt) * n_tiles;
  if (p.cluster2) {   // tiles 2j, 2j+1 (the two CTAs of a pair): pixel tiles 2i, 2i+1 of the same column tile
    const uint32_t pair = uint32_t(tile) >> 1;
    const uint32_t mtp = fdiv(pair, p.fd_ntiles);
    c.nt = int(pair) - int(mtp) * n_tiles;
    mt = 2u * mtp + (uint32_t(tile) & 1u);
  }
  const uint32_t in_ = fdiv(mt, p.fd_twh);                 // image-box index
  const uint32_t rem = mt - in_ * p.fd_twh.d;
  const uint32_t ih = fdiv(rem, p.fd_tw);
  c.w0 = int(rem - ih * p.fd_tw.d) * p.tw;
  c.h0 = int(ih) * p.th;
  c.n0 = int(in_) * p.tn;
  return c;
}

// split-K decomposition of a tile index: (m, n) tile and the K-chunk range [c_begin, c_end)
struct KRange {
  int mn, c_begin, c_end;
};
__device__ __forceinline__ KRange k_range(const ConvParams& p, int tile) {
  KRange r;
  r.mn = tile; r.c_begin = 0; r.c_end = p.k_chunks;
  if (p.k_split > 1) {
    const int ks = int(fdiv(uint32_t(tile), p.fd_mn));
    r.mn = tile - ks * p.mn_tiles;
    r.c_begin = ks * p.chunks_per_split;
    r.c_end = min(p.k_chunks, r.c_begin + p.chunks_per_split);
  }
  return r;
}

// Epilogue arithmetic for 32 columns starting at c0 of tile row r: accumulators -> scale/shift (+ residual)
// (+ ReLU) -> bf16 into the 128B-swizzled staging tile (TMA-store path) or straight to global memory.
template <int N_TILE>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t (&accr)[32], int c0,
                                               const float* s_scale, const float* s_shift, uint32_t stg, int r,
                                               bool use_tma_store, bool tma_res, bool row_ok, int64_t row_off,
                                               int grow, int gcol0) {
  {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + j);
      const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j);
      v[j] = __uint_as_float(accr[j]) * sc.x + sh.x;
      v[j + 1] = __uint_as_float(accr[j + 1]) * sc.y + sh.y;
      v[j + 2] = __uint_as_float(accr[j + 2]) * sc.z + sh.z;
      v[j + 3] = __uint_as_float(accr[j + 3]) * sc.w + sh.w;
    }
    if (p.drop_thresh != 0u) {   // linear layers only (no residual): bias, ReLU, then the dropout of the next layer's input
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t keep = dropout_keep8(p.drop_seed, p.drop_tag, uint32_t(grow), uint32_t(gcol0 + c0) / 8u + g,
                                            p.drop_thresh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = v[g * 8 + j];
          if (p.relu) x = fmaxf(x, 0.0f);
          v[g * 8 + j] = ((keep >> j) & 1u) ? x * p.drop_scale : 0.0f;
        }
      }
    }
    if (use_tma_store) {
      // staging tile: column block (c0/64), row r, 16-B chunk index XOR-swizzled by (r & 7)
      const uint32_t blk = stg + uint32_t(c0 >> 6) * GM_OUT_BLK_BYTES + uint32_t(r) * 128u;
      const uint32_t ch0 = uint32_t((c0 & 63) >> 3);  // first 16-B chunk of these 32 columns (0 or 4)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t addr = blk + (((ch0 + g) ^ uint32_t(r & 7)) << 4);
        if (tma_res) {
          uint32_t rw[4];
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(addr));
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[h]);
            v[g * 8 + h * 2] += __bfloat162float(b2.x);
            v[g * 8 + h * 2 + 1] += __bfloat162float(b2.y);
          }
        }
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float a = v[g * 8 + h * 2], b = v[g * 8 + h * 2 + 1];
          if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
          w[h] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]),
                     "r"(w[2]), "r"(w[3]) : "memory");
      }
    } else if (row_ok) {
      // fp32 output (final block / MLP head) or narrow tiles: direct, row-predicated 128-bit stores
      if (p.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 rv = __ldg(rp + g);
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[h]);
            v[g * 8 + h * 2] += __bfloat162float(b2.x);
            v[g * 8 + h * 2 + 1] += __bfloat162float(b2.y);
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
      if (p.out_f32 && p.out_col_stride != 1) {   // strided columns (OIHW weight gradient): scalar stores / reductions
        float* op = static_cast<float*>(p.out) + row_off + int64_t(c0) * p.out_col_stride;
        if (p.k_split > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(op + j * p.out_col_stride, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) op[j * p.out_col_stride] = v[j];
        }
      } else if (p.out_f32 && p.k_split > 1) {
        float* op = static_cast<float*>(p.out) + row_off + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(op + j, v[j]);   // fire-and-forget reductions (RED.ADD.F32)
      } else if (p.out_f32) {
        float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row_off + c0);
#pragma unroll
        for (int g = 0; g < 8; ++g) op[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      } else {
        uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row_off + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[g * 8 + h * 2], v[g * 8 + h * 2 + 1]);
            w[h] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          op[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
}

// The 32-column chunks of [c_begin, c_end): the TMEM load of chunk i+1 is issued before chunk i is processed
// (two register buffers), so its latency hides behind the arithmetic / shared-memory traffic of chunk i - with
// two epilogue warps per scheduler nothing else would cover it.
template <int N_TILE>
__device__ __forceinline__ void epilogue_columns(const ConvParams& p, uint32_t t_addr, int c_begin, int c_end,
                                                 const float* s_scale, const float* s_shift, uint32_t stg, int r,
                                                 bool use_tma_store, bool tma_res, bool row_ok, int64_t row_off,
                                                 int grow, int gcol0) {
  uint32_t acc_a[32], acc_b[32];
  tc::tmem_ld_32x32(t_addr + uint32_t(c_begin), acc_a);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 64) {
    tc::tmem_ld_wait();
    const bool more1 = c0 + 32 < c_end;
    if (more1) tc::tmem_ld_32x32(t_addr + uint32_t(c0 + 32), acc_b);
    epilogue_chunk<N_TILE>(p, acc_a, c0, s_scale, s_shift, stg, r, use_tma_store, tma_res, row_ok, row_off, grow, gcol0);
    if (more1) {
      tc::tmem_ld_wait();
      if (c0 + 64 < c_end) tc::tmem_ld_32x32(t_addr + uint32_t(c0 + 64), acc_a);
      epilogue_chunk<N_TILE>(p, acc_b, c0 + 32, s_scale, s_shift, stg, r, use_tma_store, tma_res, row_ok, row_off, grow,
                             gcol0);
    }
  }
}

// Training-mode BatchNorm statistics: per-channel sum and sum of squares of the bf16 values just written
// to the staging tile (128 rows x N_TILE columns, 64-column blocks, 128-byte swizzle).  A warp owns one 64-column
// block and a band of rows, a lane owns two adjacent columns: every LDS.32 of the warp reads one complete row
// (one wavefront, conflict-free).  The sums stay in registers across the tiles of the persistent CTA and are
// flushed with fire-and-forget global reductions only when the CTA moves to another column tile (never, when the
// number of column tiles divides the grid) - per-tile reductions from 148 SMs onto the same few sectors
// serialise in the L2 atomic units (measured: 16x more sector reductions = +8 ms per step).
// Tile rows of images beyond the batch hold conv(0) = 0 and add nothing (the plan rejects boxes that overhang
// the image spatially).
struct StatsAcc {
  float s0, s1, q0, q1;
  int nt;   // column tile the sums belong to (-1: empty)
};
template <int N_TILE>
__device__ __forceinline__ void stats_flush(StatsAcc& a, int t, int nthreads, float* __restrict__ stats, int c_out) {
  if (a.nt < 0) return;
  constexpr int BLKS = N_TILE / 64;
  const int warp = t >> 5, lane = t & 31, wpb = (nthreads >> 5) / BLKS;
  float* dst = stats + a.nt * N_TILE + (warp / wpb) * 64 + 2 * lane;
  atomicAdd(dst, a.s0);
  atomicAdd(dst + 1, a.s1);
  atomicAdd(dst + c_out, a.q0);
  atomicAdd(dst + c_out + 1, a.q1);
  a.s0 = a.s1 = a.q0 = a.q1 = 0.f;
  a.nt = -1;
}
template <int N_TILE>
__device__ __forceinline__ void staging_column_stats(StatsAcc& a, uint32_t stg, int t, int nthreads,
                                                     float* __restrict__ stats, int c_out, int nt) {
  if (a.nt != nt) {
    stats_flush<N_TILE>(a, t, nthreads, stats, c_out);
    a.nt = nt;
  }
  constexpr int BLKS = N_TILE / 64;
  const int warp = t >> 5, lane = t & 31, wpb = (nthreads >> 5) / BLKS;
  const int blk = warp / wpb, band = warp - blk * wpb;
  const int rows = GM_TILE_M / wpb;
  const uint32_t chunk = uint32_t(lane >> 2), word = uint32_t(lane & 3) * 4u;
  const uint32_t base = stg + uint32_t(blk) * GM_OUT_BLK_BYTES + word;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
  for (int i = 0; i < rows; ++i) {
    const int r = band * rows + i;
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(base + uint32_t(r) * 128u + ((chunk ^ uint32_t(r & 7)) << 4)));
    const float x = __uint_as_float(w << 16), y = __uint_as_float(w & 0xffff0000u);
    s0 += x;
    s1 += y;
    q0 = fmaf(x, x, q0);
    q1 = fmaf(y, y, q1);
  }
  a.s0 += s0; a.s1 += s1; a.q0 += q0; a.q1 += q1;
}

template <int N_TILE, int STAGES, int NSTG, int A_STAGE, int RES_BYTES, int CL = 1>
__global__ void __launch_bounds__(GM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(CL == 1 || (CL == 2 && RES_BYTES == 0 && N_TILE == 256), "CTA pairs: streamed 256-wide tiles only");
  const uint32_t cta_rank = (CL == 2) ? tc::cluster_ctarank() : 0u;
  using L = GemmSmem<N_TILE, STAGES, NSTG, A_STAGE, RES_BYTES>;
  constexpr bool kResident = RES_BYTES > 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = tc::smem_u32(smem_raw);
  const uint32_t base_u32 = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base_u32 - raw_u32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full_bar = base_u32 + L::BAR_OFFSET;
  const uint32_t empty_bar = full_bar + STAGES * 8;
  const uint32_t tfull_bar = empty_bar + STAGES * 8;   // [2] accumulator ready   (MMA -> epilogue)
  const uint32_t tempty_bar = tfull_bar + 16;          // [2] accumulator drained (epilogue -> MMA)
  const uint32_t res_bar = tempty_bar + 16;            // [2] residual tile landed in staging[i]
  const uint32_t wres_bar = res_bar + 16;              // resident weights landed
  const uint32_t blk_in_bar = wres_bar + 8;            // [4] staging block b usable by the epilogue (residual landed / free)
  const uint32_t blk_out_bar = blk_in_bar + 32;        // [4] staging block b written by the epilogue: store it
  const uint32_t wres_u32 = base_u32 + L::RES_OFFSET;
  const uint32_t tmem_slot = base_u32 + L::TMEM_SLOT_OFFSET;
  const uint32_t staging_u32 = base_u32 + L::STAGING_OFFSET;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFFSET);
  float* s_scale = reinterpret_cast<float*>(base_ptr + L::SCALE_OFFSET);
  float* s_shift = s_scale + N_TILE;

  const int n_tiles = p.c_out / N_TILE;
  // epilogue organisation: two independent 4-warp groups (double-staged variants without a residual),
  // else all 8 warps on one tile (with a residual the spare staging tile prefetches the next residual)
  const bool group_mode = (NSTG == 2) && !((N_TILE >= 64) && !p.out_f32 && p.residual != nullptr);

  // 256-wide bf16 tiles (no statistics): the epilogue hands every finished 64-column staging block to the output-DMA
  // warp, which stores it and, once the store has read the block out, re-arms it for the next tile - with that
  // tile's residual block loaded into it (in-place add) or simply as free.  No epilogue thread ever waits for a
  // store, and only 16 KB (not the 64 KB tile) sit between a block's last write and its reuse.
  const bool dma_mode = (N_TILE == 256) && !p.out_f32 && p.stats == nullptr;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tc::tma_prefetch_desc(&p.a_map[i]);
    tc::tma_prefetch_desc(&p.b_map);
    if (N_TILE >= 64 && !p.out_f32) tc::tma_prefetch_desc(&p.out_map);
    if (N_TILE >= 64 && !p.out_f32 && p.residual != nullptr) tc::tma_prefetch_desc(&p.res_map);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, CL);   // CTA pairs: both consumers release a stage (the peer refills half of it)
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar + 8 * a, 1);
      tc::mbar_init(tempty_bar + 8 * a, group_mode ? GM_EPI_WARPS / 2 : GM_EPI_WARPS);  // arrives per accumulator
      tc::mbar_init(res_bar + 8 * a, 1);
    }
    tc::mbar_init(wres_bar, 1);
    for (int b = 0; b < 4; ++b) {
      tc::mbar_init(blk_in_bar + 8 * b, 1);
      tc::mbar_init(blk_out_bar + 8 * b, 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, L::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync();   // the peer's barriers exist before anything is multicast onto them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) overlapped the tail of the previous kernel in the stream; its results are visible after
  // the wait.  The next kernel may start its own prologue as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer (whole warp stays converged; one elected lane issues): runs ahead across tiles =====
    if (kResident) {  // the whole weight matrix, once per CTA
      if (tc::elect_one()) {
        tc::mbar_expect_tx(wres_bar, uint32_t(p.res_boxes) * L::B_BYTES);
        for (int b = 0; b < p.res_boxes; ++b)
          tc::tma_load_2d(&p.b_map, wres_bar, wres_u32 + b * L::B_BYTES, (b % p.kb_per_tile) * GM_CHUNK_K,
                          (b / p.kb_per_tile) * N_TILE);
      }
      __syncwarp();
    }
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const KRange kr = k_range(p, tile);
      const TileCoord tcd = tile_coord(p, kr.mn, n_tiles);
      for (int t = 0; t < p.num_taps; ++t) {
        const CUtensorMap* amap = &p.a_map[p.tap_map[t]];
        const int cw = tcd.w0 + p.tap_dw[t], ch = tcd.h0 + p.tap_dh[t];
        for (int c = kr.c_begin; c < kr.c_end; ++c, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          tc::mbar_wait(empty_bar + 8 * s, ph ^ 1u);
          if (tc::elect_one()) {
            tc::mbar_expect_tx(full_bar + 8 * s, kResident ? uint32_t(p.a_tx_bytes) : uint32_t(L::STAGE_BYTES));
            const uint32_t a_dst = base_u32 + s * L::STAGE_BYTES;
            if (!kResident && p.mn_major) {
              // MN-major operands: boxes of 64 (M|N) x 64 (K rows) out of a row-major [K, M] / [K, N] matrix
              if (p.wg_conv) {
                // dW[co, ci, kh, kw] = sum over (pixel, image) of dY[n, ho, wo, co] * X[n, ho*s+kh-pad, wo*s+kw-pad, ci]:
                // K chunk c = 64 images at output pixel (ho, wo); the N tile fixes the tap and a block of ci
                const int pos = c / p.wg_nb, nb = c - pos * p.wg_nb;
                const int ho = pos / p.wg_ow, wo = pos - ho * p.wg_ow;
                const int tap = tcd.nt / p.wg_cin_tiles, ci0 = (tcd.nt - tap * p.wg_cin_tiles) * N_TILE;
                const int pad = p.wg_ksize >> 1;
                const int iy = ho * p.wg_stride + tap / p.wg_ksize - pad, ix = wo * p.wg_stride + tap % p.wg_ksize - pad;
                for (int j = 0; j < 2; ++j)
                  tc::tma_load_4d(amap, full_bar + 8 * s, a_dst + j * 8192, tcd.w0 + 64 * j, wo, ho, nb * 64);
                for (int j = 0; j < N_TILE / 64; ++j)
                  tc::tma_load_4d(&p.b_map, full_bar + 8 * s, a_dst + GM_A_BYTES + j * 8192, ci0 + 64 * j, ix, iy,
                                  nb * 64);
              } else if (p.mn_major & 1) {
                for (int j = 0; j < 2; ++j)
                  tc::tma_load_2d(amap, full_bar + 8 * s, a_dst + j * 8192, tcd.w0 + 64 * j, c * GM_CHUNK_K);
              } else {
                tc::tma_load_4d(amap, full_bar + 8 * s, a_dst, c * GM_CHUNK_K, cw, ch, tcd.n0);
              }
              // mode 2 (data gradient): K rows = forward output channels, columns = (flipped tap, forward input channel)
              const int b_col0 = ((p.mn_major & 1) ? 0 : (p.num_taps - 1 - t) * p.b_tap_stride) + tcd.nt * N_TILE;
              if (!p.wg_conv)
                for (int j = 0; j < N_TILE / 64; ++j)
                  tc::tma_load_2d(&p.b_map, full_bar + 8 * s, a_dst + GM_A_BYTES + j * 8192, b_col0 + 64 * j,
                                  c * GM_CHUNK_K);
            } else {
              tc::tma_load_4d(amap, full_bar + 8 * s, a_dst, c * GM_CHUNK_K, cw, ch, tcd.n0);
              if (CL == 2)   // this CTA's half of the weight box, for both CTAs of the pair
                tc::tma_load_2d_multicast(&p.b_map, full_bar + 8 * s, a_dst + GM_A_BYTES + cta_rank * (N_TILE / 2) * 128u,
                                          (t * p.k_chunks + c) * GM_CHUNK_K, tcd.nt * N_TILE + int(cta_rank) * (N_TILE / 2),
                                          uint16_t(3));
              else if (!kResident)
                tc::tma_load_2d(&p.b_map, full_bar + 8 * s, a_dst + GM_A_BYTES,
                                (t * p.k_chunks + c) * GM_CHUNK_K, tcd.nt * N_TILE);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp converged, one elected lane issues), accumulators ping-pong in TMEM =====
    uint32_t it = 0, tl = 0;
    if (kResident) tc::mbar_wait(wres_bar, 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      const KRange kr = k_range(p, tile);
      const uint32_t nt = uint32_t(kr.mn) - fdiv(uint32_t(kr.mn), p.fd_ntiles) * uint32_t(n_tiles);
      const uint32_t w_tile = wres_u32 + nt * uint32_t(p.kb_per_tile) * L::B_BYTES;
      const int k_iters = p.num_taps * (kr.c_end - kr.c_begin);
      tc::mbar_wait(tempty_bar + 8 * acc, aph ^ 1u);  // epilogue has drained this accumulator
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N_TILE;
      for (int ki = 0; ki < k_iters; ++ki, ++it) {
        const uint32_t s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        tc::mbar_wait(full_bar + 8 * s, ph);
        tc::tc_fence_after();
        const uint32_t a_addr = base_u32 + s * L::STAGE_BYTES;
        if (tc::elect_one()) {
          if (kResident) {
            // one A stage (possibly a halo box) feeds taps_per_stage row taps: tap u reads the same
            // box shifted by u rows of tw pixels (a multiple of the 1024-B swizzle atom)
            for (int u = 0; u < p.taps_per_stage; ++u) {
              const uint64_t da = tc::make_sw128_desc(a_addr + uint32_t(u * p.tap_row_bytes));
              const uint64_t db = tc::make_sw128_desc(w_tile + uint32_t(ki * p.taps_per_stage + u) * L::B_BYTES);
#pragma unroll
              for (int k = 0; k < GM_CHUNK_K / 16; ++k)
                tc::umma_bf16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), p.idesc,
                              (ki | u | k) != 0 ? 1u : 0u);
            }
          } else if (p.mn_major) {
            const uint64_t da_k = tc::make_sw128_desc(a_addr);
#pragma unroll
            for (int k = 0; k < GM_CHUNK_K / 16; ++k) {   // MN-major: 16 K rows = 2048 bytes further down every box
              const uint64_t da = (p.mn_major & 1) ? tc::make_sw128_mn_desc(a_addr + uint32_t(k) * 2048u)
                                                   : da_k + uint64_t(2 * k);
              tc::umma_bf16(d_tmem, da, tc::make_sw128_mn_desc(a_addr + GM_A_BYTES + uint32_t(k) * 2048u), p.idesc,
                            (ki | k) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t da = tc::make_sw128_desc(a_addr);
            const uint64_t db = tc::make_sw128_desc(a_addr + GM_A_BYTES);
#pragma unroll
            for (int k = 0; k < GM_CHUNK_K / 16; ++k) {
              // +32 B per K=16 step inside the 128-B swizzle row (start-address field is >>4)
              tc::umma_bf16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), p.idesc,
                            (ki | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem stage when these MMAs retire (CTA pairs: in both CTAs)
          if (CL == 2) tc::umma_commit_multicast(empty_bar + 8 * s, uint16_t(3));
          else tc::umma_commit(empty_bar + 8 * s);
          if (ki == k_iters - 1) tc::umma_commit(tfull_bar + 8 * acc);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp == GM_DMA_WARP) {
    // ===== output-DMA warp (256-wide path): TMA stores per 64-column block + residual loads / block recycling =====
    if (dma_mode && tc::elect_one()) {
      const bool has_res = p.residual != nullptr;
      auto provide = [&](int tile, int b) {   // make staging block b usable for `tile`
        if (has_res) {
          const TileCoord t2 = tile_coord(p, k_range(p, tile).mn, n_tiles);
          tc::mbar_expect_tx(blk_in_bar + 8 * b, GM_OUT_BLK_BYTES);
          tc::tma_load_4d(&p.res_map, blk_in_bar + 8 * b, staging_u32 + b * GM_OUT_BLK_BYTES, t2.nt * N_TILE + b * 64,
                          t2.w0, t2.h0, t2.n0);
        } else {
          tc::mbar_arrive(blk_in_bar + 8 * b);
        }
      };
      if (int(blockIdx.x) < p.total_tiles)
        for (int b = 0; b < 4; ++b) provide(blockIdx.x, b);
      uint32_t tl = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
        const TileCoord tcd = tile_coord(p, k_range(p, tile).mn, n_tiles);
        const int next = tile + int(gridDim.x);
        // the halves finish blocks (0, 2) first, then (1, 3)
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int b = ((i & 1) << 1) | (i >> 1);
          tc::mbar_wait(blk_out_bar + 8 * b, tl & 1u);
          asm volatile(
              "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                  reinterpret_cast<uint64_t>(&p.out_map)),
              "r"(staging_u32 + b * GM_OUT_BLK_BYTES), "r"(tcd.nt * N_TILE + b * 64), "r"(tcd.w0), "r"(tcd.h0),
              "r"(tcd.n0)
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (i >= 1) {   // the previous block's store has been read out: recycle that block
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            const int pb = (((i - 1) & 1) << 1) | ((i - 1) >> 1);
            if (next < p.total_tiles) provide(next, pb);
          }
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (next < p.total_tiles) provide(next, 3);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ===== 8 epilogue warps.  TMEM lane quarter = warp % 4 (tile rows). =====
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    StatsAcc st_acc = {0.f, 0.f, 0.f, 0.f, -1};   // training-mode BatchNorm sums of this thread's two columns
    const bool use_tma_store = (N_TILE >= 64) && !p.out_f32;
    const bool tma_res = use_tma_store && (p.residual != nullptr);
    // position of tile row r inside the (tw x th x tn) pixel box (the sides are powers of two)
    const int r_w = r & (p.tw - 1);
    const int r_h = (r >> p.tw_shift) & (p.th - 1);
    const int r_n = r >> p.twh_shift;

    auto issue_residual = [&](int tile, uint32_t sb) {
      const TileCoord t2 = tile_coord(p, tile, n_tiles);
      tc::mbar_expect_tx(res_bar + 8 * sb, L::OUT_BLKS * GM_OUT_BLK_BYTES);
      for (int b = 0; b < L::OUT_BLKS; ++b)
        tc::tma_load_4d(&p.res_map, res_bar + 8 * sb, staging_u32 + sb * L::STAGING_BYTES + b * GM_OUT_BLK_BYTES,
                        t2.nt * N_TILE + b * 64, t2.w0, t2.h0, t2.n0);
    };
    auto issue_store = [&](const TileCoord& tcd, uint32_t stg) {
      for (int b = 0; b < L::OUT_BLKS; ++b) {
        asm volatile(
            "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                reinterpret_cast<uint64_t>(&p.out_map)),
            "r"(stg + b * GM_OUT_BLK_BYTES), "r"(tcd.nt * N_TILE + b * 64), "r"(tcd.w0), "r"(tcd.h0), "r"(tcd.n0)
            : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    };
    auto load_scale_shift = [&](int nt, float* sc, float* sh, int t0, int nthreads) {
      for (int i = t0; i < N_TILE; i += nthreads) {
        const int n = nt * N_TILE + i;
        sc[i] = p.scale ? __ldg(p.scale + n) : 1.0f;
        sh[i] = p.shift ? __ldg(p.shift + n) : 0.0f;
      }
    };

    if (group_mode) {
      // Two independent groups of 4 warps; tiles (and with them the TMEM accumulators and the
      // staging tiles) alternate between the groups, so one group's TMEM/store/residual latencies
      // overlap the other group's arithmetic.
      const int gt = threadIdx.x - 64 - grp * 128;  // 0..127 inside the group
      const uint32_t stg = staging_u32 + grp * L::STAGING_BYTES;
      float* g_scale = s_scale + grp * 2 * N_TILE;
      float* g_shift = g_scale + N_TILE;
      const bool fixed_nt = (n_tiles == 1);
      const int tstride = 2 * int(gridDim.x);
      int tile = int(blockIdx.x) + grp * int(gridDim.x);
      if (fixed_nt) {
        load_scale_shift(0, g_scale, g_shift, gt, 128);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      if (tma_res && gt == 0 && tile < p.total_tiles) issue_residual(tile, grp);
      for (uint32_t use = 0; tile < p.total_tiles; tile += tstride, ++use) {
        const TileCoord tcd = tile_coord(p, k_range(p, tile).mn, n_tiles);
        if (!fixed_nt) load_scale_shift(tcd.nt, g_scale, g_shift, gt, 128);
        // staging[grp] is free: with a residual, its arrival implies the previous store was read out;
        // otherwise the leader waited for the read-out right after committing it
        if (!fixed_nt || (use_tma_store && !tma_res)) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        bool row_ok = false;
        int64_t row_off = 0;
        if (!use_tma_store) {   // only the direct-store path needs per-row addresses
          const int pw = tcd.w0 + r_w, phh = tcd.h0 + r_h, pn = tcd.n0 + r_n;
          row_ok = (pw < p.out_w) && (phh < p.out_h) && (pn < p.batch);
          row_off = ((int64_t(pn) * p.out_h + phh) * p.out_w + pw) * p.c_out + int64_t(tcd.nt) * N_TILE;
          if (p.wg_conv) {   // OIHW: ((co * Cin + ci) * k*k + tap), co = pw
            const int tap = tcd.nt / p.wg_cin_tiles, ci0 = (tcd.nt - tap * p.wg_cin_tiles) * N_TILE;
            row_off = (int64_t(pw) * p.wg_cin + ci0) * p.out_col_stride + tap;
          }
        }
        tc::mbar_wait(tfull_bar + 8 * grp, use & 1u);
        tc::tc_fence_after();
        if (tma_res) tc::mbar_wait(res_bar + 8 * grp, use & 1u);
        const uint32_t t_addr = tmem_base + uint32_t(grp) * N_TILE + (uint32_t(q * 32) << 16);
        epilogue_columns<N_TILE>(p, t_addr, 0, N_TILE, g_scale, g_shift, stg, r, use_tma_store, tma_res, row_ok,
                                 row_off, tcd.w0 + r_w, tcd.nt * N_TILE);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty_bar + 8 * grp);
        if (use_tma_store) {
          tc::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (gt == 0) issue_store(tcd, stg);
          if constexpr (N_TILE >= 64) {
            if (p.stats != nullptr) staging_column_stats<N_TILE>(st_acc, stg, gt, 128, p.stats, p.c_out, tcd.nt);
          }
          if (gt == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (tma_res && tile + tstride < p.total_tiles) issue_residual(tile + tstride, grp);
          }
        } else if (!fixed_nt) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // scale/shift reloaded next tile
        }
      }
      if (use_tma_store && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if constexpr (N_TILE >= 64) {
        if (p.stats != nullptr) stats_flush<N_TILE>(st_acc, gt, 128, p.stats, p.c_out);
      }
    } else {
      // All 8 warps work on the same tile, the column range split in two halves.  With two staging
      // tiles (NSTG == 2) the TMA store of tile i drains while tile i+1 is computed and the residual of
      // tile i+1 is prefetched into the spare staging tile during tile i.
      constexpr int COLS_PER_HALF = (N_TILE / 2 >= 32) ? N_TILE / 2 : 32;
      constexpr int ACTIVE_HALVES = N_TILE / COLS_PER_HALF;  // 2, or 1 when N_TILE == 32
      const int et = threadIdx.x - 64;  // 0..255
      const bool active = grp < ACTIVE_HALVES;
      const int col_lo = grp * COLS_PER_HALF;
      const bool tma_res_old = tma_res && !dma_mode;   // residual through res_bar / issue_residual (non-DMA variants)
      if (NSTG == 2 && tma_res_old && et == 0 && int(blockIdx.x) < p.total_tiles) issue_residual(blockIdx.x, 0);
      uint32_t tl = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
        const TileCoord tcd = tile_coord(p, k_range(p, tile).mn, n_tiles);
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        const uint32_t sb = (NSTG == 2) ? (tl & 1u) : 0u;
        const uint32_t res_parity = (NSTG == 2) ? ((tl >> 1) & 1u) : (tl & 1u);
        const uint32_t stg = staging_u32 + sb * L::STAGING_BYTES;
        // staging[sb] was last read by the TMA store of tile (tl - NSTG): it must be done reading
        if (use_tma_store && !dma_mode && et == 0) {
          if (NSTG == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        load_scale_shift(tcd.nt, s_scale, s_shift, et, GM_EPI_THREADS);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tma_res_old && et == 0) {
          if (NSTG == 2) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (tile + int(gridDim.x) < p.total_tiles) issue_residual(tile + gridDim.x, sb ^ 1u);
          } else {
            issue_residual(tile, 0);
          }
        }
        bool row_ok = false;
        int64_t row_off = 0;
        if (!use_tma_store) {   // only the direct-store path needs per-row addresses
          const int pw = tcd.w0 + r_w, phh = tcd.h0 + r_h, pn = tcd.n0 + r_n;
          row_ok = (pw < p.out_w) && (phh < p.out_h) && (pn < p.batch);
          row_off = ((int64_t(pn) * p.out_h + phh) * p.out_w + pw) * p.c_out + int64_t(tcd.nt) * N_TILE;
          if (p.wg_conv) {   // OIHW: ((co * Cin + ci) * k*k + tap), co = pw
            const int tap = tcd.nt / p.wg_cin_tiles, ci0 = (tcd.nt - tap * p.wg_cin_tiles) * N_TILE;
            row_off = (int64_t(pw) * p.wg_cin + ci0) * p.out_col_stride + tap;
          }
        }
        tc::mbar_wait(tfull_bar + 8 * acc, aph);
        tc::tc_fence_after();
        if (tma_res_old) tc::mbar_wait(res_bar + 8 * sb, res_parity);
        const uint32_t t_addr = tmem_base + acc * N_TILE + (uint32_t(q * 32) << 16);
        if (dma_mode) {
#pragma unroll 1
          for (int b = 0; b < 2; ++b) {
            const int blk = grp * 2 + b;
            tc::mbar_wait(blk_in_bar + 8 * blk, tl & 1u);   // previous store read out (+ this tile's residual landed)
            epilogue_columns<N_TILE>(p, t_addr, blk * 64, blk * 64 + 64, s_scale, s_shift, stg, r, true, tma_res, false,
                                     0, tcd.w0 + r_w, tcd.nt * N_TILE);
            if (b == 1) {   // accumulator fully read: hand it back to the MMA warp
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(tempty_bar + 8 * acc);
            }
            tc::fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA engine
            asm volatile("bar.sync %0, 128;" ::"r"(3 + grp) : "memory");
            if ((et & 127) == 0) tc::mbar_arrive(blk_out_bar + 8 * blk);
          }
          continue;
        }
        if (active)
          epilogue_columns<N_TILE>(p, t_addr, col_lo, col_lo + COLS_PER_HALF, s_scale, s_shift, stg, r,
                                   use_tma_store, tma_res, row_ok, row_off, tcd.w0 + r_w, tcd.nt * N_TILE);
        // accumulator fully read: hand it back to the MMA warp before the stores drain
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty_bar + 8 * acc);
        if (use_tma_store) tc::fence_proxy_async();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (use_tma_store && et == 0) issue_store(tcd, stg);
        if constexpr (N_TILE >= 64) {
          if (use_tma_store && p.stats != nullptr)
            staging_column_stats<N_TILE>(st_acc, stg, et, GM_EPI_THREADS, p.stats, p.c_out, tcd.nt);
        }
      }
      if (use_tma_store && !dma_mode && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if constexpr (N_TILE >= 64) {
        if (p.stats != nullptr) stats_flush<N_TILE>(st_acc, et, GM_EPI_THREADS, p.stats, p.c_out);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync();   // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) tc::tmem_dealloc(tmem_base, L::TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MMBS_ERR_DEVICE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(base), gdim,
                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
              int(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return MMBS_ERR_CUDA;
  }
  return MMBS_OK;
}

}  // namespace mmbs

using namespace mmbs;

struct mmbs_conv_plan {
  ConvParams p;
  int variant;   // 0 = streamed weights, 1 = weights resident in smem
  int cluster;   // 1, or 2: CTA pairs sharing the weight boxes through TMA multicast
  int n_tile;
  int stages;
  unsigned grid;
  size_t zero_bytes;   // split-K: the fp32 output is cleared before every run (the splits accumulate into it)
};

template <int N_TILE, int STAGES, int NSTG, int A_STAGE = 0, int RES_BYTES = 0, int CL = 1>
static int launch_conv(const mmbs_conv_plan* plan, cudaStream_t stream) {
  using L = GemmSmem<N_TILE, STAGES, NSTG, A_STAGE, RES_BYTES>;
  static PerDeviceOnce configured;
  if (configured.first())
    MMBS_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, STAGES, NSTG, A_STAGE, RES_BYTES, CL>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYNAMIC));
  static const bool use_pdl = []() {
    const char* e = getenv("MMBS_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = L::DYNAMIC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl && plan->zero_bytes == 0) {   // a memset precedes split-K launches: plain ordering
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  MMBS_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<N_TILE, STAGES, NSTG, A_STAGE, RES_BYTES, CL>, plan->p));
  count_launch();
  return MMBS_OK;
}

extern "C" int mmbs_conv_run(const mmbs_conv_plan* plan, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(plan != nullptr, "mmbs_conv_run: null plan");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (plan->zero_bytes) MMBS_CUDA_TRY(cudaMemsetAsync(plan->p.out, 0, plan->zero_bytes, stream));
  if (plan->variant == 1) {
    if (plan->n_tile == 64) return launch_conv<64, 5, 2, GM_RES64_A_STAGE, GM_RES64_BYTES>(plan, stream);
    if (plan->n_tile == 128) return launch_conv<128, 5, 2, GM_A_BYTES, GM_RES128_BYTES>(plan, stream);
    set_error("mmbs_conv_run: no resident variant for n_tile %d", plan->n_tile);
    return MMBS_ERR_ARG;
  }
  switch (plan->n_tile) {
    case 256:
      if (plan->cluster == 2) return launch_conv<256, 3, 1, 0, 0, 2>(plan, stream);
      return launch_conv<256, 3, 1>(plan, stream);
    case 128: return launch_conv<128, 4, 2>(plan, stream);
    case 64: return launch_conv<64, 6, 2>(plan, stream);
    case 32: return launch_conv<32, 6, 1>(plan, stream);
    default: set_error("mmbs_conv_run: bad n_tile %d", plan->n_tile); return MMBS_ERR_ARG;
  }
}

extern "C" void mmbs_conv_plan_destroy(mmbs_conv_plan* plan) { delete plan; }

// Dropout of the NEXT layer's input, applied by this linear layer's epilogue (after bias / ReLU).  p = 0 switches it
// off.  Takes effect from the next mmbs_conv_run; the mask is common.cuh dropout_keep8(seed, tag, row, column / 8).
extern "C" int mmbs_plan_set_dropout(mmbs_conv_plan* plan, float p, uint64_t seed, uint32_t tag) {
  MMBS_REQUIRE(plan != nullptr, "mmbs_plan_set_dropout: null plan");
  MMBS_REQUIRE(p >= 0.f && p < 1.f, "mmbs_plan_set_dropout: p=%f", double(p));
  ConvParams& q = plan->p;
  MMBS_REQUIRE(p == 0.f || (q.out_h == 1 && q.batch == 1 && q.th == 1 && q.tn == 1 && q.residual == nullptr &&
                            q.stats == nullptr && q.k_split <= 1 && !q.wg_conv),
               "mmbs_plan_set_dropout: only plain linear layers (rows x features) take a fused dropout");
  const float t = p * 65536.0f + 0.5f;
  q.drop_thresh = p > 0.f ? (t >= 65535.0f ? 65535u : uint32_t(t)) : 0u;
  q.drop_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  q.drop_seed = seed;
  q.drop_tag = tag;
  return MMBS_OK;
}

// choose the (tw, th, tn) output-pixel box (product 128) that wastes the fewest rows
// (exact_spatial: only boxes that tile the image exactly - the batch-statistics epilogue counts every tile row)
static void choose_box(int ow, int oh, int b, bool exact_spatial, int* tw, int* th, int* tn) {
  double best = -1.0;
  for (int w = 1; w <= 128; w *= 2)
    for (int h = 1; w * h <= 128; h *= 2) {
      const int n = 128 / (w * h);
      if (w > 256 || h > 256 || n > 256) continue;
      if (exact_spatial && (ow % w != 0 || oh % h != 0)) continue;
      const double tiles = double(ceil_div(ow, w)) * double(ceil_div(oh, h)) * double(ceil_div(b, n));
      const double util = double(ow) * oh * b / (tiles * 128.0);
      const double score = util + 1e-6 * w + 1e-7 * h;  // ties: wider boxes (longer contiguous runs)
      if (score > best) {
        best = score;
        *tw = w; *th = h; *tn = n;
      }
    }
}

static int pick_n_tile(int c_out, int64_t m_tiles, int min_tile) {
  // N tile that minimises (waves over the SMs) x (cost of one tile ~ N + 32): a narrower tile only pays when it
  // removes a wave.  Measured at B = 128: 7x7 512->512 3x3 with 98 tiles of 256 (one wave) 43 us, with 196 tiles of
  // 128 (two waves) 52 us; 14x14 256->256 3x3 with 196 tiles of 256 (two waves) 43 us vs 392 tiles of 128 45 us.
  const int sms = sm_count();
  int best = 0;
  int64_t best_cost = 0;
  for (int nt = 256; nt >= 32; nt >>= 1) {
    if (c_out % nt != 0) continue;
    if (nt < min_tile && best != 0) break;   // below the minimum only when nothing wider divides c_out
    const int64_t tiles = m_tiles * (c_out / nt);
    const int64_t cost = ((tiles + sms - 1) / sms) * (nt + 32);
    if (best == 0 || cost < best_cost) {
      best = nt;
      best_cost = cost;
    }
  }
  return best ? best : 32;
}

static int build_plan(const mmbs_conv_desc* d, int stem_mode, int linear_mode, mmbs_conv_plan** out) {
  MMBS_REQUIRE(d && out, "conv plan: null argument");
  MMBS_REQUIRE(d->in && d->weight && d->out, "conv plan: null tensor pointer");
  MMBS_REQUIRE(d->c_out > 0 && d->c_out % 32 == 0, "conv plan: c_out=%d must be a multiple of 32", d->c_out);
  MMBS_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0, "conv plan: bad shape");
  MMBS_REQUIRE(reinterpret_cast<uintptr_t>(d->in) % 16 == 0 && reinterpret_cast<uintptr_t>(d->weight) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(d->out) % 16 == 0,
               "conv plan: pointers must be 16-byte aligned");
  std::unique_ptr<mmbs_conv_plan> guard(new (std::nothrow) mmbs_conv_plan());   // freed on every error return
  mmbs_conv_plan* plan = guard.get();
  MMBS_REQUIRE(plan, "conv plan: out of host memory");
  ConvParams& p = plan->p;
  std::memset(&p, 0, sizeof(p));
  int rc = MMBS_OK;
  const int s = d->stride, k = d->ksize;
  int out_h, out_w, k_total;
  if (stem_mode) {
    // input = space-to-depth buffer [B,116,116,16]; 4 row taps of 64 contiguous elements
    out_h = 112; out_w = 112;
    p.num_taps = 4; p.k_chunks = 1; k_total = 256;
    for (int a = 0; a < 4; ++a) { p.tap_map[a] = 0; p.tap_dw[a] = 0; p.tap_dh[a] = int8_t(a); }
  } else {
    MMBS_REQUIRE(d->c_in > 0 && (linear_mode == 2 || d->c_in % 64 == 0), "conv plan: c_in=%d must be a multiple of 64",
                 d->c_in);
    MMBS_REQUIRE((k == 1 || k == 3) && (s == 1 || s == 2), "conv plan: ksize=%d stride=%d unsupported", k, s);
    MMBS_REQUIRE(s == 1 || (d->in_h % 2 == 0 && d->in_w % 2 == 0), "conv plan: stride 2 needs even H, W");
    const int pad = k / 2;
    out_h = (d->in_h + 2 * pad - k) / s + 1;
    out_w = (d->in_w + 2 * pad - k) / s + 1;
    p.num_taps = k * k; p.k_chunks = (d->c_in + 63) / 64; k_total = k * k * d->c_in;   // (TN GEMM: K rows, OOB zero)
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int t = kh * k + kw;
        const int rh = kh - pad, rw = kw - pad;  // input offset relative to stride*out
        if (s == 1) {
          p.tap_map[t] = 0; p.tap_dh[t] = int8_t(rh); p.tap_dw[t] = int8_t(rw);
        } else {
          const int ph = rh & 1, pw = rw & 1;           // parity view
          p.tap_map[t] = int8_t(ph * 2 + pw);
          p.tap_dh[t] = int8_t((rh - ph) / 2);          // floor(rh / 2)
          p.tap_dw[t] = int8_t((rw - pw) / 2);
        }
      }
  }
  p.out_w = out_w; p.out_h = out_h; p.batch = d->batch; p.c_out = d->c_out;
  p.relu = d->relu; p.out_f32 = d->out_f32;
  p.scale = d->scale; p.shift = d->shift;
  p.residual = static_cast<const __nv_bfloat16*>(d->residual);
  p.out = d->out;
  p.stats = d->stats;
  MMBS_REQUIRE(!d->stats || (d->c_out % 64 == 0 && !d->out_f32 && !d->residual && !d->scale && !d->shift && !d->relu),
               "conv plan: batch statistics need a raw bf16 output with c_out %% 64 == 0 (c_out=%d)", d->c_out);
  // ---- variant selection
  // resident: the whole weight matrix fits in shared memory -> loaded once per CTA, the ring holds A only
  // halo    : (resident, 64-wide) one A box with extra rows feeds several row taps (stem: 4, 3x3: 3 per kw)
  const int64_t w_bytes = int64_t(k_total) * d->c_out * 2;
  bool resident = false, halo = false;
  int forced_n_tile = 0;
  if (!linear_mode && !d->out_f32 && !(d->flags & 2)) {
    if (d->c_out == 64 && w_bytes <= GM_RES64_BYTES) { resident = true; forced_n_tile = 64; }
    // (256-output convs with tiny K - layer1 conv3 / downsample - are faster on the streamed 256-wide path with the
    //  output-DMA warp: 310 vs 332 us with the residual, 190 vs 215 us without; MMBS_RES256=1 restores the old choice)
    else if ((d->c_out == 128 || (d->c_out == 256 && getenv("MMBS_RES256") != nullptr)) && w_bytes <= GM_RES128_BYTES) { resident = true; forced_n_tile = 128; }
  }
  const bool want_halo = (d->flags & 1) != 0;
  if (resident && forced_n_tile == 64 && (stem_mode || (want_halo && k == 3 && s == 1))) halo = true;
  if (want_halo && !halo) {
    set_error("conv plan: halo weight order requested for an ineligible conv (k=%d s=%d c_out=%d)", k, s, d->c_out);
    return MMBS_ERR_ARG;
  }
  p.taps_per_stage = 1; p.tap_row_bytes = 0; p.a_tx_bytes = GM_A_BYTES;
  int halo_rows = 0;
  if (linear_mode) { p.tw = 128; p.th = 1; p.tn = 1; }
  else if (halo && stem_mode) {
    p.tw = 16; p.th = 8; p.tn = 1; halo_rows = 3;
    p.num_taps = 1; p.taps_per_stage = 4; p.tap_dw[0] = 0; p.tap_dh[0] = 0; p.tap_map[0] = 0;
  } else if (halo) {
    p.tw = 8; p.th = 16; p.tn = 1; halo_rows = 2;
    p.num_taps = 3; p.taps_per_stage = 3;   // stage t = kw; row taps kh = 0..2 inside the box
    for (int t = 0; t < 3; ++t) { p.tap_map[t] = 0; p.tap_dw[t] = int8_t(t - 1); p.tap_dh[t] = -1; }
  } else choose_box(out_w, out_h, d->batch, d->stats != nullptr, &p.tw, &p.th, &p.tn);
  if (halo) {
    p.tap_row_bytes = p.tw * 128;
    p.a_tx_bytes = (p.th + halo_rows) * p.tw * 128;
    MMBS_REQUIRE(p.a_tx_bytes <= GM_RES64_A_STAGE, "conv plan: halo box too large");
  }
  p.tiles_w = int(ceil_div(out_w, p.tw)); p.tiles_h = int(ceil_div(out_h, p.th)); p.tiles_n = int(ceil_div(d->batch, p.tn));
  const int64_t m_tiles = int64_t(p.tiles_w) * p.tiles_h * p.tiles_n;
  // weight-gradient shaped GEMMs (fp32 output, no epilogue arithmetic): keep the widest N tile and split K
  const bool can_split = linear_mode && d->out_f32 && !d->scale && !d->shift && !d->relu && !d->residual && !d->stats &&
                         getenv("MMBS_NO_SPLITK") == nullptr;
  plan->n_tile = pick_n_tile(d->c_out, m_tiles, can_split ? 256 : ((d->stats || (d->flags & 2)) ? 64 : 32));   // statistics: TMA-store path
  MMBS_REQUIRE(!d->stats || (out_w % p.tw == 0 && out_h % p.th == 0),
               "conv plan: batch statistics need pixel boxes that tile the %dx%d output exactly", out_h, out_w);
  // epilogue-bound residual layers (short K loop): 128-wide tile = double-staged epilogue with the
  // residual prefetched one tile ahead; long K loops keep the 256-wide tile (operand-feed bound)
  // (the 256-wide path recycles its staging blocks through the output-DMA warp, residual included; the 128-wide
  //  double-staged variant is kept selectable for comparison: MMBS_RES_NTILE=128)
  if (const char* e = getenv("MMBS_RES_NTILE"))
    if (atoi(e) == 128 && d->residual && !d->out_f32 && plan->n_tile > 128 && p.num_taps * p.k_chunks <= 8) plan->n_tile = 128;
  if (const char* e = getenv("MMBS_FORCE_NTILE")) {   // experiments only
    const int f = atoi(e);
    if ((f == 64 || f == 128 || f == 256) && d->c_out % f == 0 && !can_split) plan->n_tile = f;
  }
  plan->variant = 0;
  if (resident) {
    plan->variant = 1;
    plan->n_tile = forced_n_tile;
    p.kb_per_tile = p.num_taps * p.k_chunks * p.taps_per_stage;
    p.res_boxes = (d->c_out / plan->n_tile) * p.kb_per_tile;
  }
  plan->stages = 0;
  MMBS_REQUIRE(m_tiles * (d->c_out / plan->n_tile) < (int64_t(1) << 29), "conv plan: grid too large");
  p.mn_tiles = int32_t(m_tiles * (d->c_out / plan->n_tile));
  p.k_split = 1; p.chunks_per_split = p.k_chunks;
  plan->zero_bytes = 0;
  if (can_split && p.mn_tiles < sm_count()) {
    int ks = std::min(sm_count() / p.mn_tiles, p.k_chunks / 4);   // fill one wave, >= 4 K chunks per split
    if (ks >= 2) {
      p.chunks_per_split = (p.k_chunks + ks - 1) / ks;
      p.k_split = (p.k_chunks + p.chunks_per_split - 1) / p.chunks_per_split;   // no empty split
      plan->zero_bytes = size_t(d->in_w) * size_t(d->c_out) * sizeof(float);
    }
  }
  p.fd_mn = make_fastdiv(uint32_t(p.mn_tiles));
  p.total_tiles = p.mn_tiles * p.k_split;
  plan->grid = unsigned(std::min<int64_t>(p.total_tiles, sm_count()));  // persistent: <= one CTA per SM
  // CTA pairs (MMBS_CLUSTER=1): forward convs on the streamed 256-wide path with an even number of pixel tiles
  plan->cluster = 1;
  {
    static const bool want = []() {
      const char* e = getenv("MMBS_CLUSTER");
      return e && e[0] == '1';
    }();
    if (want && !linear_mode && !stem_mode && plan->variant == 0 && plan->n_tile == 256 && p.k_split == 1 &&
        !(d->flags & 2) && m_tiles % 2 == 0 && plan->grid >= 2) {
      plan->cluster = 2;
      p.cluster2 = 1;
      plan->grid &= ~1u;
    }
  }
  const bool fwd_weights = linear_mode != 2 && !stem_mode && (d->flags & 2) != 0;   // dgrad on the forward weights
  p.mn_major = (linear_mode == 2) ? 3 : (fwd_weights ? 2 : 0);
  p.out_col_stride = 1;
  p.b_tap_stride = d->c_out;
  MMBS_REQUIRE(!fwd_weights || (!resident && plan->n_tile >= 64 && s == 1),
               "conv plan: forward-weight data gradients need stride 1, c_out %% 64 == 0 and streamed weights");
  p.idesc = make_idesc_bf16(GM_TILE_M, plan->n_tile, linear_mode == 2, linear_mode == 2 || fwd_weights);
  p.fd_ntiles = make_fastdiv(uint32_t(d->c_out / plan->n_tile));
  p.fd_tw = make_fastdiv(uint32_t(p.tiles_w));
  p.fd_twh = make_fastdiv(uint32_t(p.tiles_w) * uint32_t(p.tiles_h));
  p.tw_shift = 0; while ((1 << p.tw_shift) < p.tw) ++p.tw_shift;
  p.twh_shift = 0; while ((1 << p.twh_shift) < p.tw * p.th) ++p.twh_shift;
  MMBS_REQUIRE((1 << p.tw_shift) == p.tw && (1 << p.twh_shift) == p.tw * p.th, "conv plan: pixel box sides must be powers of two");

  const uint32_t box_a[4] = {64u, uint32_t(p.tw), uint32_t(p.th + halo_rows), uint32_t(p.tn)};
  const uint32_t box_out[4] = {64u, uint32_t(p.tw), uint32_t(p.th), uint32_t(p.tn)};
  const char* in = static_cast<const char*>(d->in);
  if (linear_mode == 2) {
    // TN GEMM out[M, N] = A[K, M]^T B[K, N]: both operands row-major with K outermost, 64 x 64 boxes
    MMBS_REQUIRE(plan->n_tile >= 64 && d->in_w % 8 == 0 && d->c_out % 64 == 0,
                 "TN linear plan: need M %% 8 == 0 and N %% 64 == 0 (M=%d N=%d)", d->in_w, d->c_out);
    const uint32_t box[2] = {64u, 64u};
    const uint64_t adims[2] = {uint64_t(d->in_w), uint64_t(d->c_in)};
    const uint64_t astr[1] = {uint64_t(d->in_w) * 2};
    rc = encode_map(&p.a_map[0], in, 2, adims, astr, box);
    for (int i = 1; i < 4 && !rc; ++i) p.a_map[i] = p.a_map[0];
    if (!rc) {
      const uint64_t bdims[2] = {uint64_t(d->c_out), uint64_t(d->c_in)};
      const uint64_t bstr[1] = {uint64_t(d->c_out) * 2};
      rc = encode_map(&p.b_map, d->weight, 2, bdims, bstr, box);
    }
    if (rc) return rc;
    *out = guard.release();
    return MMBS_OK;
  }
  if (stem_mode) {
    const uint64_t dims[4] = {64, 113, 116, uint64_t(d->batch)};
    const uint64_t str[3] = {32, 116ull * 32, 116ull * 116 * 32};
    rc = encode_map(&p.a_map[0], in, 4, dims, str, box_a);
    for (int i = 1; i < 4 && !rc; ++i) p.a_map[i] = p.a_map[0];
  } else if (s == 1) {
    const uint64_t C = uint64_t(d->c_in), W = uint64_t(d->in_w), H = uint64_t(d->in_h);
    const uint64_t dims[4] = {C, W, H, uint64_t(d->batch)};
    const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    rc = encode_map(&p.a_map[0], in, 4, dims, str, box_a);
    for (int i = 1; i < 4 && !rc; ++i) p.a_map[i] = p.a_map[0];
  } else {
    const uint64_t C = uint64_t(d->c_in), W = uint64_t(d->in_w), H = uint64_t(d->in_h);
    for (int ph = 0; ph < 2 && !rc; ++ph)
      for (int pw = 0; pw < 2 && !rc; ++pw) {
        const uint64_t dims[4] = {C, W / 2, H / 2, uint64_t(d->batch)};
        const uint64_t str[3] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2};
        rc = encode_map(&p.a_map[ph * 2 + pw], in + (uint64_t(ph) * W + pw) * C * 2, 4, dims, str, box_a);
      }
  }
  if (!rc && fwd_weights) {
    // the forward conv's matrix [c_in (its c_out) rows][taps * c_out (its c_in) columns], 64 x 64 boxes (MN-major B)
    const uint64_t cols = uint64_t(k) * k * d->c_out;
    const uint64_t dims[2] = {cols, uint64_t(d->c_in)};
    const uint64_t str[1] = {cols * 2};
    const uint32_t box_b[2] = {64u, 64u};
    rc = encode_map(&p.b_map, d->weight, 2, dims, str, box_b);
  } else if (!rc) {
    const uint64_t dims[2] = {uint64_t(k_total), uint64_t(d->c_out)};
    const uint64_t str[1] = {uint64_t(k_total) * 2};
    const uint32_t box_b[2] = {64u, uint32_t(plan->n_tile / plan->cluster)};   // (CTA pairs: each CTA fetches half a box)
    rc = encode_map(&p.b_map, d->weight, 2, dims, str, box_b);
  }
  if (!rc && !d->out_f32 && plan->n_tile >= 64) {
    // TMA-store view of the bf16 output (and TMA-load view of the residual): same pixel box as A
    const uint64_t C = uint64_t(d->c_out), W = uint64_t(out_w), H = uint64_t(out_h);
    const uint64_t dims[4] = {C, W, H, uint64_t(d->batch)};
    const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    rc = encode_map(&p.out_map, d->out, 4, dims, str, box_out);
    if (!rc && d->residual) {
      MMBS_REQUIRE(reinterpret_cast<uintptr_t>(d->residual) % 16 == 0, "conv plan: residual must be 16-byte aligned");
      rc = encode_map(&p.res_map, d->residual, 4, dims, str, box_out);
    }
  }
  if (rc) return rc;
  *out = guard.release();
  return MMBS_OK;
}

extern "C" int mmbs_conv_plan_create(const mmbs_conv_desc* desc, mmbs_conv_plan** plan_out) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(desc != nullptr, "mmbs_conv_plan_create: null desc");
  return build_plan(desc, desc->ksize == 4 ? 1 : 0, 0, plan_out);
}

extern "C" int mmbs_linear_plan_create(const void* x_bf16, const void* w_bf16, const float* bias, void* y,
                                       int64_t m, int64_t n, int64_t k, int32_t relu, int32_t out_f32,
                                       mmbs_conv_plan** plan_out) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(m > 0 && n > 0 && k > 0 && m < (int64_t(1) << 31) && k % 64 == 0 && n % 32 == 0,
               "mmbs_linear_plan_create: need k %% 64 == 0 and n %% 32 == 0 (m=%lld n=%lld k=%lld)",
               (long long)m, (long long)n, (long long)k);
  mmbs_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.batch = 1; d.in_h = 1; d.in_w = int32_t(m); d.c_in = int32_t(k); d.c_out = int32_t(n);
  d.ksize = 1; d.stride = 1; d.relu = relu; d.out_f32 = out_f32;
  d.in = x_bf16; d.weight = w_bf16; d.scale = nullptr; d.shift = bias; d.residual = nullptr; d.out = y;
  return build_plan(&d, 0, 1, plan_out);
}

/* TN GEMM: y[M, N] (fp32) = a[K, M]^T b[K, N], a / b bf16 row-major with K outermost - the weight gradient
 * dW[co, ci] = sum_p dY[p, co] X[p, ci] straight from the NHWC tensors (no transposed copies). */
extern "C" int mmbs_linear_tn_plan_create(const void* a_km_bf16, const void* b_kn_bf16, float* y, int64_t m, int64_t n,
                                          int64_t k, mmbs_conv_plan** plan_out) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(m > 0 && n > 0 && k > 0 && m < (int64_t(1) << 31) && k < (int64_t(1) << 31) && m % 8 == 0 && n % 64 == 0,
               "mmbs_linear_tn_plan_create: need m %% 8 == 0 and n %% 64 == 0 (m=%lld n=%lld k=%lld)", (long long)m,
               (long long)n, (long long)k);
  mmbs_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.batch = 1; d.in_h = 1; d.in_w = int32_t(m); d.c_in = int32_t(k); d.c_out = int32_t(n);
  d.ksize = 1; d.stride = 1; d.relu = 0; d.out_f32 = 1;
  d.in = a_km_bf16; d.weight = b_kn_bf16; d.out = y;
  return build_plan(&d, 0, 2, plan_out);
}

/* NN GEMM: y[M, N] = x[M, K] w[K, N] (+ bias, ReLU): w is row-major with K outermost - the data gradient
 * dh = dz W of a linear layer straight from the forward weight matrix W[N_out, K_in] (MN-major B operand). */
extern "C" int mmbs_linear_nn_plan_create(const void* x_bf16, const void* w_kn_bf16, const float* bias, void* y,
                                          int64_t m, int64_t n, int64_t k, int32_t relu, int32_t out_f32,
                                          mmbs_conv_plan** plan_out) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(m > 0 && n > 0 && k > 0 && m < (int64_t(1) << 31) && k % 64 == 0 && n % 64 == 0,
               "mmbs_linear_nn_plan_create: need k %% 64 == 0 and n %% 64 == 0 (m=%lld n=%lld k=%lld)", (long long)m,
               (long long)n, (long long)k);
  mmbs_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.batch = 1; d.in_h = 1; d.in_w = int32_t(m); d.c_in = int32_t(k); d.c_out = int32_t(n);
  d.ksize = 1; d.stride = 1; d.relu = relu; d.out_f32 = out_f32; d.flags = 2;
  d.in = x_bf16; d.weight = w_kn_bf16; d.scale = nullptr; d.shift = bias; d.residual = nullptr; d.out = y;
  return build_plan(&d, 0, 1, plan_out);
}

/* Implicit weight-gradient GEMM of a convolution: dw[co, ci, kh, kw] (fp32, OIHW) = sum_{n, ho, wo}
 * dy[n, ho, wo, co] * x[n, ho*stride + kh - pad, wo*stride + kw - pad, ci]; dy / x bf16 NHWC.  Both operands are
 * fed MN-major straight from the NHWC tensors (K chunk = 64 images at one output pixel, zero fill = padding);
 * split-K with fp32 reductions into the cleared output.  c_in % 64 == 0 (256 | c_in when c_in > 128), c_out % 8 == 0. */
extern "C" int mmbs_conv_wgrad_plan_create(const void* dy_bf16, const void* x_bf16, float* dw_oihw, int64_t batch,
                                           int64_t in_h, int64_t in_w, int64_t c_in, int64_t c_out, int64_t ksize,
                                           int64_t stride, mmbs_conv_plan** plan_out) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(dy_bf16 && x_bf16 && dw_oihw && plan_out && batch > 0 && in_h > 0 && in_w > 0 && c_in > 0 && c_out > 0 &&
                   (ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) && c_in % 64 == 0 && c_out % 8 == 0,
               "mmbs_conv_wgrad_plan_create: bad argument");
  const int pad = int(ksize / 2);
  const int64_t oh = (in_h + 2 * pad - ksize) / stride + 1, ow = (in_w + 2 * pad - ksize) / stride + 1;
  const int64_t kk = ksize * ksize, nb = (batch + 63) / 64;
  mmbs_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.batch = 1; d.in_h = 1; d.in_w = int32_t(c_out);                 // M = c_out rows
  d.c_in = int32_t(oh * ow * nb * 64);                               // K = 64-image chunks over every output pixel
  d.c_out = int32_t(kk * c_in);                                      // N = (tap, ci)
  d.ksize = 1; d.stride = 1; d.out_f32 = 1;
  d.in = dy_bf16; d.weight = x_bf16; d.out = dw_oihw;
  mmbs_conv_plan* plan = nullptr;
  if (int rc = build_plan(&d, 0, 2, &plan)) return rc;
  std::unique_ptr<mmbs_conv_plan> guard(plan);
  ConvParams& p = plan->p;
  MMBS_REQUIRE(c_in % plan->n_tile == 0, "mmbs_conv_wgrad_plan_create: c_in=%lld must be a multiple of the %d-wide N tile",
               (long long)c_in, plan->n_tile);
  p.wg_conv = 1; p.wg_nb = int32_t(nb); p.wg_ow = int32_t(ow); p.wg_stride = int32_t(stride);
  p.wg_ksize = int32_t(ksize); p.wg_cin = int32_t(c_in); p.wg_cin_tiles = int32_t(c_in / plan->n_tile);
  p.out_col_stride = int32_t(kk);
  const uint32_t box[4] = {64u, 1u, 1u, 64u};
  {
    const uint64_t dims[4] = {uint64_t(c_out), uint64_t(ow), uint64_t(oh), uint64_t(batch)};
    const uint64_t str[3] = {uint64_t(c_out) * 2, uint64_t(ow) * c_out * 2, uint64_t(oh) * ow * c_out * 2};
    if (int rc = encode_map(&p.a_map[0], dy_bf16, 4, dims, str, box)) return rc;
    for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  }
  {
    const uint64_t dims[4] = {uint64_t(c_in), uint64_t(in_w), uint64_t(in_h), uint64_t(batch)};
    const uint64_t str[3] = {uint64_t(c_in) * 2, uint64_t(in_w) * c_in * 2, uint64_t(in_h) * in_w * c_in * 2};
    if (int rc = encode_map(&p.b_map, x_bf16, 4, dims, str, box)) return rc;
  }
  *plan_out = guard.release();
  return MMBS_OK;
}
