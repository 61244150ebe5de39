// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a).
//
// One warp-specialised kernel serves every conv of ResNet.forward_extract
// (/root/reference/5_JointFusion/resnet.py:151-165, Bottleneck.forward :70-90) and every
// nn.Linear of the RNA / fusion MLPs (2_GeneExpression/1_GeneExpress_train.py:247-257 ...):
//
//   out[m, n] = act( scale[n] * sum_k A[m, k] * W[n, k] + shift[n] (+ residual[m, n]) )
//
// * A rows are output pixels.  An M tile is a BOX of 128 output pixels (tw x th x tn
//   over W, H, N), so the A operand of any filter tap is a plain 4-D TMA box of the
//   NHWC input shifted by the tap offset (out-of-range rows/cols zero-filled by TMA =
//   the conv padding).  Stride-2 convs read one of four "parity" views of the input
//   (base pointer offset + doubled strides), so they are shifted boxes as well.
// * W is [Cout][kh][kw][Cin] bf16: K-major, one 2-D TMA box per (tap, 64-channel chunk).
// * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (accumulator in TMEM),
//   warps 2-5 = epilogue (tcgen05.ld -> scale/shift (+residual) (+ReLU) -> bf16/fp32 store).
// * smem ring of STAGES x (A 16 KB + B N_TILE*128 B), 128-byte swizzle end to end.
#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"
#include "tc_common.cuh"

namespace mmbs {

constexpr int GM_TILE_M = 128;
constexpr int GM_CHUNK_K = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int GM_A_BYTES = GM_TILE_M * GM_CHUNK_K * 2;  // 16 KB
constexpr int GM_THREADS = 192;
constexpr int GM_MAX_TAPS = 16;

struct ConvParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  int32_t tw, th, tn;
  int32_t tiles_w, tiles_h, tiles_n;
  int32_t out_w, out_h, batch;
  int32_t c_out;
  int32_t num_taps, k_chunks;
  int32_t relu, out_f32;
  uint32_t idesc;
  int8_t tap_map[GM_MAX_TAPS];
  int8_t tap_dw[GM_MAX_TAPS];
  int8_t tap_dh[GM_MAX_TAPS];
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
};

template <int N_TILE, int STAGES>
struct GemmSmem {
  static constexpr int B_BYTES = N_TILE * GM_CHUNK_K * 2;
  static constexpr int STAGE_BYTES = GM_A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;            // full[S], empty[S], accum
  static constexpr int TMEM_SLOT_OFFSET = BAR_OFFSET + (2 * STAGES + 1) * 8;
  static constexpr int SCALE_OFFSET = (TMEM_SLOT_OFFSET + 4 + 15) / 16 * 16;
  static constexpr int TOTAL = SCALE_OFFSET + 2 * N_TILE * 4;
  static constexpr int DYNAMIC = TOTAL + 1024;  // slack for 1024-B alignment of the ring
};

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(GM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  using L = GemmSmem<N_TILE, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = tc::smem_u32(smem_raw);
  const uint32_t base_u32 = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base_u32 - raw_u32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full_bar = base_u32 + L::BAR_OFFSET;
  const uint32_t empty_bar = full_bar + STAGES * 8;
  const uint32_t accum_bar = empty_bar + STAGES * 8;
  const uint32_t tmem_slot = base_u32 + L::TMEM_SLOT_OFFSET;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFFSET);
  float* s_scale = reinterpret_cast<float*>(base_ptr + L::SCALE_OFFSET);
  float* s_shift = s_scale + N_TILE;

  // tile coordinates: n-tile fastest so CTAs that share an A tile run together (L2 reuse)
  const int n_tiles = p.c_out / N_TILE;
  const int nt = blockIdx.x % n_tiles;
  const int mt = blockIdx.x / n_tiles;
  const int w0 = (mt % p.tiles_w) * p.tw;
  const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
  const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;
  const int total_iters = p.num_taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tc::tma_prefetch_desc(&p.a_map[i]);
    tc::tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, 1);
    }
    tc::mbar_init(accum_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, N_TILE);
    tc::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < N_TILE; i += GM_THREADS - 64) {
      const int n = nt * N_TILE + i;
      s_scale[i] = p.scale ? __ldg(p.scale + n) : 1.0f;
      s_shift[i] = p.shift ? __ldg(p.shift + n) : 0.0f;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int it = 0;
      for (int t = 0; t < p.num_taps; ++t) {
        const CUtensorMap* amap = &p.a_map[p.tap_map[t]];
        const int cw = w0 + p.tap_dw[t], ch = h0 + p.tap_dh[t];
        for (int c = 0; c < p.k_chunks; ++c, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          tc::mbar_wait(empty_bar + 8 * s, ph ^ 1u);
          tc::mbar_expect_tx(full_bar + 8 * s, L::STAGE_BYTES);
          const uint32_t a_dst = base_u32 + s * L::STAGE_BYTES;
          tc::tma_load_4d(amap, full_bar + 8 * s, a_dst, c * GM_CHUNK_K, cw, ch, n0);
          tc::tma_load_2d(&p.b_map, full_bar + 8 * s, a_dst + GM_A_BYTES,
                          (t * p.k_chunks + c) * GM_CHUNK_K, nt * N_TILE);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread) =====
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        tc::mbar_wait(full_bar + 8 * s, ph);
        tc::tc_fence_after();
        const uint32_t a_addr = base_u32 + s * L::STAGE_BYTES;
        const uint64_t da = tc::make_sw128_desc(a_addr);
        const uint64_t db = tc::make_sw128_desc(a_addr + GM_A_BYTES);
#pragma unroll
        for (int k = 0; k < GM_CHUNK_K / 16; ++k) {
          // +32 B per K=16 step inside the 128-B swizzle row (start-address field is >>4)
          tc::umma_bf16(tmem_base, da + uint64_t(2 * k), db + uint64_t(2 * k), p.idesc,
                        (it | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(empty_bar + 8 * s);  // frees the smem stage when these MMAs retire
      }
      tc::umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM lane quarter (warp % 4) <-> tile rows =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int pw = w0 + (r % p.tw);
    const int phh = h0 + ((r / p.tw) % p.th);
    const int pn = n0 + (r / (p.tw * p.th));
    const bool row_ok = (pw < p.out_w) && (phh < p.out_h) && (pn < p.batch);
    const int64_t row = (int64_t(pn) * p.out_h + phh) * p.out_w + pw;
    const int64_t row_off = row * p.c_out + int64_t(nt) * N_TILE;
    tc::mbar_wait(accum_bar, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      uint32_t acc[32];
      tc::tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0), acc);
      tc::tmem_ld_wait();
      if (row_ok) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * s_scale[c0 + j] + s_shift[c0 + j];
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
        if (p.out_f32) {
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
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, N_TILE);
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
  int n_tile;
  int stages;
  unsigned grid;
};

template <int N_TILE, int STAGES>
static int launch_conv(const mmbs_conv_plan* plan, cudaStream_t stream) {
  using L = GemmSmem<N_TILE, STAGES>;
  static bool configured = false;
  if (!configured) {
    MMBS_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, STAGES>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYNAMIC));
    configured = true;
  }
  conv_gemm_kernel<N_TILE, STAGES><<<plan->grid, GM_THREADS, L::DYNAMIC, stream>>>(plan->p);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_conv_run(const mmbs_conv_plan* plan, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(plan != nullptr, "mmbs_conv_run: null plan");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (plan->n_tile) {
    case 256: return launch_conv<256, 4>(plan, stream);
    case 128: return launch_conv<128, 6>(plan, stream);
    case 64: return launch_conv<64, 8>(plan, stream);
    case 32: return launch_conv<32, 8>(plan, stream);
    default: set_error("mmbs_conv_run: bad n_tile %d", plan->n_tile); return MMBS_ERR_ARG;
  }
}

extern "C" void mmbs_conv_plan_destroy(mmbs_conv_plan* plan) { delete plan; }

// choose the (tw, th, tn) output-pixel box (product 128) that wastes the fewest rows
static void choose_box(int ow, int oh, int b, int* tw, int* th, int* tn) {
  double best = -1.0;
  for (int w = 1; w <= 128; w *= 2)
    for (int h = 1; w * h <= 128; h *= 2) {
      const int n = 128 / (w * h);
      if (w > 256 || h > 256 || n > 256) continue;
      const double tiles = double(ceil_div(ow, w)) * double(ceil_div(oh, h)) * double(ceil_div(b, n));
      const double util = double(ow) * oh * b / (tiles * 128.0);
      const double score = util + 1e-6 * w + 1e-7 * h;  // ties: wider boxes (longer contiguous runs)
      if (score > best) {
        best = score;
        *tw = w; *th = h; *tn = n;
      }
    }
}

static int pick_n_tile(int c_out, int64_t m_tiles) {
  // widest tile that divides c_out, narrowed while the grid cannot fill the GPU
  int nt = 256;
  while (nt > 32 && (c_out % nt) != 0) nt >>= 1;
  const int sms = sm_count();
  while (nt > 32 && m_tiles * (c_out / nt) < sms) nt >>= 1;
  return nt;
}

static int build_plan(const mmbs_conv_desc* d, int stem_mode, int linear_mode, mmbs_conv_plan** out) {
  MMBS_REQUIRE(d && out, "conv plan: null argument");
  MMBS_REQUIRE(d->in && d->weight && d->out, "conv plan: null tensor pointer");
  MMBS_REQUIRE(d->c_out > 0 && d->c_out % 32 == 0, "conv plan: c_out=%d must be a multiple of 32", d->c_out);
  MMBS_REQUIRE(d->batch > 0 && d->in_h > 0 && d->in_w > 0, "conv plan: bad shape");
  MMBS_REQUIRE(reinterpret_cast<uintptr_t>(d->in) % 16 == 0 && reinterpret_cast<uintptr_t>(d->weight) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(d->out) % 16 == 0,
               "conv plan: pointers must be 16-byte aligned");
  mmbs_conv_plan* plan = new (std::nothrow) mmbs_conv_plan();
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
    MMBS_REQUIRE(d->c_in > 0 && d->c_in % 64 == 0, "conv plan: c_in=%d must be a multiple of 64", d->c_in);
    MMBS_REQUIRE((k == 1 || k == 3) && (s == 1 || s == 2), "conv plan: ksize=%d stride=%d unsupported", k, s);
    MMBS_REQUIRE(s == 1 || (d->in_h % 2 == 0 && d->in_w % 2 == 0), "conv plan: stride 2 needs even H, W");
    const int pad = k / 2;
    out_h = (d->in_h + 2 * pad - k) / s + 1;
    out_w = (d->in_w + 2 * pad - k) / s + 1;
    p.num_taps = k * k; p.k_chunks = d->c_in / 64; k_total = k * k * d->c_in;
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
  if (linear_mode) { p.tw = 128; p.th = 1; p.tn = 1; }
  else choose_box(out_w, out_h, d->batch, &p.tw, &p.th, &p.tn);
  p.tiles_w = int(ceil_div(out_w, p.tw)); p.tiles_h = int(ceil_div(out_h, p.th)); p.tiles_n = int(ceil_div(d->batch, p.tn));
  const int64_t m_tiles = int64_t(p.tiles_w) * p.tiles_h * p.tiles_n;
  plan->n_tile = pick_n_tile(d->c_out, m_tiles);
  plan->stages = 0;
  MMBS_REQUIRE(m_tiles * (d->c_out / plan->n_tile) < (int64_t(1) << 31), "conv plan: grid too large");
  plan->grid = unsigned(m_tiles * (d->c_out / plan->n_tile));
  p.idesc = make_idesc_bf16(GM_TILE_M, plan->n_tile);

  const uint32_t box_a[4] = {64u, uint32_t(p.tw), uint32_t(p.th), uint32_t(p.tn)};
  const char* in = static_cast<const char*>(d->in);
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
  if (!rc) {
    const uint64_t dims[2] = {uint64_t(k_total), uint64_t(d->c_out)};
    const uint64_t str[1] = {uint64_t(k_total) * 2};
    const uint32_t box_b[2] = {64u, uint32_t(plan->n_tile)};
    rc = encode_map(&p.b_map, d->weight, 2, dims, str, box_b);
  }
  if (rc) {
    delete plan;
    return rc;
  }
  *out = plan;
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
