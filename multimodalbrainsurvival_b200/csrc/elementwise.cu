// HBM-bound glue kernels around the tcgen05 conv/GEMM kernel (sm_100a):
// layout packing (NCHW fp32 -> NHWC bf16), BN folding, max/avg pooling, casts.
// Reference sites: ResNet.forward_extract /root/reference/5_JointFusion/resnet.py:151-165.
#include <algorithm>
#include <cuda_bf16.h>

#include "common.cuh"

namespace mmbs {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&u);
  return make_float2(__bfloat162float(v.x), __bfloat162float(v.y));
}

// ---- stem input: [B,3,224,224] fp32 -> space-to-depth, zero padded [B,116,116,16] bf16.
// buffer pixel (r, s), channel (p*2+q)*3 + c  =  x[c][2(r-2)+p][2(s-2)+q]   (12 of 16 used)
__global__ void __launch_bounds__(128) stem_pack_input_kernel(const float* __restrict__ x,
                                                              uint4* __restrict__ out, int64_t batch) {
  const int64_t pix = int64_t(blockIdx.x) * 128 + threadIdx.x;
  const int64_t total = batch * 116 * 116;
  if (pix >= total) return;
  const int s = int(pix % 116), r = int((pix / 116) % 116);
  const int64_t n = pix / (116 * 116);
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  const int col = 2 * (s - 2);
  if (col >= 0 && col < 224) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int row = 2 * (r - 2) + p;
      if (row >= 0 && row < 224) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(x + ((n * 3 + c) * 224 + row) * 224 + col));
          v[(p * 2 + 0) * 3 + c] = t.x;
          v[(p * 2 + 1) * 3 + c] = t.y;
        }
      }
    }
  }
  uint4 o0, o1;
  o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
  o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
  o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
  o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
  out[pix * 2] = o0;
  out[pix * 2 + 1] = o1;
}

// ---- stem input from raw pixels: uint8 [B,3,224,224] -> ((x/255 - mean[c]) / std[c]) -> the same
// space-to-depth bf16 buffer.  This is ToTensor()+Normalize() of the reference's transforms
// (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:133-137) done on the device, so
// the host ships 1 byte per value instead of 4.
__global__ void __launch_bounds__(128) stem_pack_input_u8_kernel(const uint8_t* __restrict__ x,
                                                                 uint4* __restrict__ out, int64_t batch,
                                                                 float m0, float m1, float m2, float i0, float i1,
                                                                 float i2) {
  const int64_t pix = int64_t(blockIdx.x) * 128 + threadIdx.x;
  const int64_t total = batch * 116 * 116;
  if (pix >= total) return;
  const int s = int(pix % 116), r = int((pix / 116) % 116);
  const int64_t n = pix / (116 * 116);
  const float mean[3] = {m0, m1, m2}, inv[3] = {i0, i1, i2};
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  const int col = 2 * (s - 2);
  if (col >= 0 && col < 224) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int row = 2 * (r - 2) + p;
      if (row >= 0 && row < 224) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uchar2 t = __ldg(reinterpret_cast<const uchar2*>(x + ((n * 3 + c) * 224 + row) * 224 + col));
          v[(p * 2 + 0) * 3 + c] = (float(t.x) * (1.0f / 255.0f) - mean[c]) * inv[c];
          v[(p * 2 + 1) * 3 + c] = (float(t.y) * (1.0f / 255.0f) - mean[c]) * inv[c];
        }
      }
    }
  }
  uint4 o0, o1;
  o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
  o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
  o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
  o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
  out[pix * 2] = o0;
  out[pix * 2 + 1] = o1;
}

// ---- 1- and 4-channel stems (RNone / RNfour, /root/reference/5_JointFusion/resnet.py:167-337): the same space-to-depth
// buffer with C channels per (row parity, column parity) slot: element (p * 2 + q) * C + c, the rest of the 16 zero.
template <int C>
__global__ void __launch_bounds__(128) stem_pack_input_c_kernel(const float* __restrict__ x, uint4* __restrict__ out,
                                                                int64_t batch) {
  static_assert(C >= 1 && C <= 4, "16 slots hold 4 pixel positions of up to 4 channels");
  const int64_t pix = int64_t(blockIdx.x) * 128 + threadIdx.x;
  const int64_t total = batch * 116 * 116;
  if (pix >= total) return;
  const int s = int(pix % 116), r = int((pix / 116) % 116);
  const int64_t n = pix / (116 * 116);
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  const int col = 2 * (s - 2);
  if (col >= 0 && col < 224) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int row = 2 * (r - 2) + p;
      if (row >= 0 && row < 224) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(x + ((n * C + c) * 224 + row) * 224 + col));
          v[(p * 2 + 0) * C + c] = t.x;
          v[(p * 2 + 1) * C + c] = t.y;
        }
      }
    }
  }
  uint4 o0, o1;
  o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
  o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
  o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
  o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
  out[pix * 2] = o0;
  out[pix * 2 + 1] = o1;
}

// stem weight [64,C,7,7] fp32 -> [64][a(4)][b(4)][(p*2+q)*C+c (16)] bf16 (same tap order as stem_pack_weight_kernel)
__global__ void stem_pack_weight_c_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 256) return;
  const int ch = i & 15, b = (i >> 4) & 3, a = (i >> 6) & 3, o = i >> 8;
  float v = 0.f;
  if (ch < 4 * C) {
    const int c = ch % C, pq = ch / C, p = pq >> 1, q = pq & 1;
    const int kh = 2 * a + p - 1, kw = 2 * b + q - 1;
    if (kh >= 0 && kw >= 0) v = w[((o * C + c) * 7 + kh) * 7 + kw];
  }
  out[i] = __float2bfloat16_rn(v);
}

// ---- stem weight: [64,3,7,7] fp32 -> [64][a(4)][b(4)][(p*2+q)*3+c (16)] bf16,
// kh = 2a+p-1, kw = 2b+q-1 (the 7x7 kernel zero-extended to 8x8 at the top/left).
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 256) return;
  const int ch = i & 15, b = (i >> 4) & 3, a = (i >> 6) & 3, o = i >> 8;
  float v = 0.f;
  if (ch < 12) {
    const int c = ch % 3, pq = ch / 3, p = pq >> 1, q = pq & 1;
    const int kh = 2 * a + p - 1, kw = 2 * b + q - 1;
    if (kh >= 0 && kw >= 0) v = w[((o * 3 + c) * 7 + kh) * 7 + kw];
  }
  out[i] = __float2bfloat16_rn(v);
}

// ---- conv weight OIHW fp32 -> [O][kh][kw][I] bf16
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                        int64_t c_out, int64_t c_in, int64_t k) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = c_out * k * k * c_in;
  if (i >= total) return;
  const int64_t ci = i % c_in, kw = (i / c_in) % k, kh = (i / (c_in * k)) % k, o = i / (c_in * k * k);
  out[i] = __float2bfloat16_rn(w[((o * c_in + ci) * k + kh) * k + kw]);
}

// k = 1: a plain cast, 4 elements per thread
__global__ void __launch_bounds__(256) cast_bf16_vec4_kernel(const float4* __restrict__ in, uint2* __restrict__ out,
                                                             int64_t n4) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(in + i);
  out[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
// k = 3: one thread per (o, ci) moves the 9 taps (contiguous 36-byte read, 9 writes coalesced over ci)
__global__ void __launch_bounds__(256) pack_conv_weight3_kernel(const float* __restrict__ w,
                                                                __nv_bfloat16* __restrict__ out, int64_t c_out,
                                                                int64_t c_in) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= c_out * c_in) return;
  const int64_t ci = i % c_in, o = i / c_in;
  const float* src = w + i * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) out[(o * 9 + t) * c_in + ci] = __float2bfloat16_rn(__ldg(src + t));
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               int64_t c, float* __restrict__ scale, float* __restrict__ shift) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float sc = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = sc;
  shift[i] = beta[i] - mean[i] * sc;
}

// ---- MaxPool2d(3, 2, 1), NHWC bf16; one thread = one output pixel x 8 channels
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                   *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// grid = (ceil(ow*c8 / 256), ceil(oh / MP_ROWS), batch): no integer divisions on the hot path.  A thread produces
// MP_ROWS vertically adjacent outputs of one (x, 8-channel group): the 2*MP_ROWS+1 input rows they cover are each
// reduced over their 3 columns once (6.75 loads per output instead of 9; an input row is fetched by 1.125 blocks
// instead of 1.5).
constexpr int MP_ROWS = 4;
__global__ void __launch_bounds__(256) maxpool_3x3s2_kernel(const uint4* __restrict__ in,
                                                            uint4* __restrict__ out, int64_t batch, int h,
                                                            int w, int c8, int c8_shift) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  const int xc = blockIdx.x * 256 + threadIdx.x;      // (x, channel group) flattened, c8 = 1 << c8_shift
  if (xc >= ow * c8) return;
  const int g = xc & (c8 - 1), x = xc >> c8_shift;
  const int y0 = blockIdx.y * MP_ROWS;
  const int64_t n = blockIdx.z;
  const uint32_t NEG = 0xff80ff80u;  // (-inf, -inf) in bf16
  const uint4* base = in + n * int64_t(h) * w * c8 + g;
  uint4 m[MP_ROWS];
#pragma unroll
  for (int k = 0; k < MP_ROWS; ++k) m[k] = make_uint4(NEG, NEG, NEG, NEG);
#pragma unroll
  for (int r = 0; r < 2 * MP_ROWS + 1; ++r) {
    const int iy = 2 * y0 - 1 + r;
    uint4 rm = make_uint4(NEG, NEG, NEG, NEG);
    if (iy >= 0 && iy < h) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = 2 * x - 1 + dx;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = __ldg(base + (int64_t(iy) * w + ix) * c8);
        rm.x = max_bf16x2(rm.x, v.x); rm.y = max_bf16x2(rm.y, v.y);
        rm.z = max_bf16x2(rm.z, v.z); rm.w = max_bf16x2(rm.w, v.w);
      }
    }
    // input row r belongs to outputs k with 2k <= r <= 2k + 2
#pragma unroll
    for (int k = 0; k < MP_ROWS; ++k) {
      if (r >= 2 * k && r <= 2 * k + 2) {
        m[k].x = max_bf16x2(m[k].x, rm.x); m[k].y = max_bf16x2(m[k].y, rm.y);
        m[k].z = max_bf16x2(m[k].z, rm.z); m[k].w = max_bf16x2(m[k].w, rm.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MP_ROWS; ++k)
    if (y0 + k < oh) out[((n * oh + y0 + k) * int64_t(ow) + x) * c8 + g] = m[k];
}

// ---- AvgPool2d(7) + flatten: [B, hw, C] (bf16 or fp32) -> fp32 [B, C]
__global__ void __launch_bounds__(256) avgpool_bf16_kernel(const uint4* __restrict__ in,
                                                           float* __restrict__ out, int64_t batch, int hw,
                                                           int c8) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= batch * c8) return;
  const int g = int(i % c8);
  const int64_t n = i / c8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < hw; ++p) {
    const uint4 v = __ldg(in + (n * hw + p) * c8 + g);
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
  }
  const float inv = 1.0f / float(hw);
  float4* o = reinterpret_cast<float4*>(out + (n * c8 + g) * 8);
  o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}
__global__ void __launch_bounds__(256) avgpool_f32_kernel(const float4* __restrict__ in,
                                                          float4* __restrict__ out, int64_t batch, int hw,
                                                          int c4) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= batch * c4) return;
  const int g = int(i % c4);
  const int64_t n = i / c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < hw; ++p) {
    const float4 v = __ldg(in + (n * hw + p) * c4 + g);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / float(hw);
  out[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// ---- fp32 [rows, cols] -> bf16 [rows, cols_padded] (zero pad); 8 outputs per thread
__global__ void __launch_bounds__(256) cast_pad_bf16_kernel(const float* __restrict__ in,
                                                            uint4* __restrict__ out, int64_t rows,
                                                            int64_t cols, int64_t cp8) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * cp8) return;
  const int64_t r = i / cp8, c0 = (i % cp8) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? __ldg(in + r * cols + c0 + j) : 0.f;
  out[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                      pack_bf16x2(v[6], v[7]));
}

}  // namespace mmbs

using namespace mmbs;

static inline unsigned blocks_for(int64_t total, int threads) {
  return unsigned(std::max<int64_t>(1, ceil_div(total, threads)));
}

extern "C" int mmbs_stem_pack_input(const float* x_nchw, void* out, int64_t batch, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x_nchw && out && batch > 0, "mmbs_stem_pack_input: bad argument");
  MMBS_REQUIRE(reinterpret_cast<uintptr_t>(x_nchw) % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "mmbs_stem_pack_input: misaligned pointer");
  stem_pack_input_kernel<<<blocks_for(batch * 116 * 116, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, static_cast<uint4*>(out), batch);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_stem_pack_input_u8(const uint8_t* x_nchw, void* out, int64_t batch, const float* mean_host,
                                       const float* std_host, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x_nchw && out && batch > 0 && mean_host && std_host, "mmbs_stem_pack_input_u8: bad argument");
  MMBS_REQUIRE(reinterpret_cast<uintptr_t>(x_nchw) % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "mmbs_stem_pack_input_u8: misaligned pointer");
  stem_pack_input_u8_kernel<<<blocks_for(batch * 116 * 116, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, static_cast<uint4*>(out), batch, mean_host[0], mean_host[1], mean_host[2], 1.0f / std_host[0],
      1.0f / std_host[1], 1.0f / std_host[2]);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_stem_pack_input_c(const float* x_nchw, void* out, int64_t batch, int channels, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x_nchw && out && batch > 0, "mmbs_stem_pack_input_c: bad argument");
  MMBS_REQUIRE(channels == 1 || channels == 3 || channels == 4, "mmbs_stem_pack_input_c: channels=%d (1, 3 or 4)", channels);
  MMBS_REQUIRE(reinterpret_cast<uintptr_t>(x_nchw) % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "mmbs_stem_pack_input_c: misaligned pointer");
  const unsigned grid = blocks_for(batch * 116 * 116, 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (channels == 1) stem_pack_input_c_kernel<1><<<grid, 128, 0, st>>>(x_nchw, static_cast<uint4*>(out), batch);
  else if (channels == 3) stem_pack_input_c_kernel<3><<<grid, 128, 0, st>>>(x_nchw, static_cast<uint4*>(out), batch);
  else stem_pack_input_c_kernel<4><<<grid, 128, 0, st>>>(x_nchw, static_cast<uint4*>(out), batch);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_stem_pack_weight_c(const float* w, void* out, int channels, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(w && out, "mmbs_stem_pack_weight_c: null pointer");
  MMBS_REQUIRE(channels == 1 || channels == 3 || channels == 4, "mmbs_stem_pack_weight_c: channels=%d (1, 3 or 4)", channels);
  stem_pack_weight_c_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<__nv_bfloat16*>(out), channels);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_stem_pack_weight(const float* w, void* out, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(w && out, "mmbs_stem_pack_weight: null pointer");
  stem_pack_weight_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<__nv_bfloat16*>(out));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_pack_conv_weight(const float* w, void* out, int64_t c_out, int64_t c_in, int64_t k,
                                     void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(w && out && c_out > 0 && c_in > 0 && k > 0, "mmbs_pack_conv_weight: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = c_out * c_in * k * k;
  if (k == 1 && total % 4 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0)
    cast_bf16_vec4_kernel<<<blocks_for(total / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(w),
                                                                     static_cast<uint2*>(out), total / 4);
  else if (k == 3)
    pack_conv_weight3_kernel<<<blocks_for(c_out * c_in, 256), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out), c_out,
                                                                           c_in);
  else
    pack_conv_weight_kernel<<<blocks_for(total, 256), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out), c_out, c_in, k);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                            float eps, int64_t c, float* scale, float* shift, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(gamma && beta && mean && var && scale && shift && c > 0, "mmbs_bn_fold: bad argument");
  bn_fold_kernel<<<blocks_for(c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(gamma, beta, mean, var, eps, c,
                                                                                    scale, shift);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_maxpool_3x3s2(const void* in, void* out, int64_t batch, int64_t h, int64_t w, int64_t c,
                                  void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "mmbs_maxpool_3x3s2: bad argument");
  const int64_t oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const int c8 = int(c / 8);
  int shift = 0;
  while ((1 << shift) < c8) ++shift;
  MMBS_REQUIRE((1 << shift) == c8 && batch <= 65535 && oh <= 65535, "mmbs_maxpool_3x3s2: c/8 must be a power of two");
  dim3 grid(unsigned(ceil_div(ow * c8, 256)), unsigned(ceil_div(oh, MP_ROWS)), unsigned(batch));
  maxpool_3x3s2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), static_cast<uint4*>(out), batch, int(h), int(w), c8, shift);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_avgpool_global(const void* in, float* out, int64_t batch, int64_t hw, int64_t c,
                                   void* stream) {
  // bf16 input; see mmbs_avgpool_global_f32 for fp32 input
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && batch > 0 && hw > 0 && c > 0 && c % 8 == 0, "mmbs_avgpool_global: bad argument");
  avgpool_bf16_kernel<<<blocks_for(batch * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), out, batch, int(hw), int(c / 8));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_avgpool_global_f32(const float* in, float* out, int64_t batch, int64_t hw, int64_t c,
                                       void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && batch > 0 && hw > 0 && c > 0 && c % 4 == 0, "mmbs_avgpool_global_f32: bad argument");
  avgpool_f32_kernel<<<blocks_for(batch * (c / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), batch, int(hw), int(c / 4));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_cast_pad_bf16(const float* in, void* out, int64_t rows, int64_t cols, int64_t cols_padded,
                                  void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && rows > 0 && cols > 0 && cols_padded >= cols && cols_padded % 8 == 0,
               "mmbs_cast_pad_bf16: bad argument");
  cast_pad_bf16_kernel<<<blocks_for(rows * (cols_padded / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<uint4*>(out), rows, cols, cols_padded / 8);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

// ====================================================================== MLP training glue
// Dropout masks come from Philox-4x32-10 keyed by (seed, tag) with the element's float4 index as
// counter, so the backward pass regenerates the forward mask instead of storing it
// (nn.Dropout of the reference MLPs: /root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257).
namespace mmbs {

// in (fp32 or bf16) [rows, cols] -> bf16 [rows, cols_padded]: dropout(p) then cast, zero padding.
__global__ void __launch_bounds__(256) dropout_cast_kernel(const void* __restrict__ in, int in_bf16,
                                                           int64_t in_stride, uint2* __restrict__ out,
                                                           int64_t rows, int64_t cols, int64_t cp4, float p,
                                                           uint64_t seed, uint32_t tag) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * cp4) return;
  const int64_t r = i / cp4, q = i % cp4, c0 = q * 4;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = 0.f;
    if (c0 + j < cols)
      x = in_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(in)[r * in_stride + c0 + j])
                  : __ldg(static_cast<const float*>(in) + r * in_stride + c0 + j);
    v[j] = x;
  }
  if (p > 0.f) {
    const uint32_t keep = dropout_keep4(seed, tag, uint32_t(r), uint32_t(q), drop_threshold(p));
    const float sc = 1.0f / (1.0f - p);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * sc : 0.f;
  }
  out[i] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// Same, 8 columns per thread with 16-byte accesses (cols, in_stride, cols_padded multiples of 8, aligned pointers).
// The mask is the one of dropout_cast_kernel: Philox counter = (column / 4, row).
__global__ void __launch_bounds__(256) dropout_cast_vec8_kernel(const void* __restrict__ in, int in_bf16,
                                                                int64_t in_stride, uint4* __restrict__ out,
                                                                int64_t rows, int64_t cols, int64_t cp8, float p,
                                                                uint64_t seed, uint32_t tag) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * cp8) return;
  const int64_t r = i / cp8, g = i % cp8, c0 = g * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (c0 < cols) {   // cols % 8 == 0: a group is entirely inside or entirely padding
    if (in_bf16) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(in) + r * in_stride + c0));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    } else {
      const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(in) + r * in_stride + c0);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (p > 0.f) {
      const uint32_t thresh = drop_threshold(p);
      const uint32_t k8 = dropout_keep8(seed, tag, uint32_t(r), uint32_t(g), thresh);
      const float sc = 1.0f / (1.0f - p);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((k8 >> j) & 1u) ? v[j] * sc : 0.f;
    }
  }
  out[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                      pack_bf16x2(v[6], v[7]));
}

// Backward elementwise step of one layer:
//   dz[m, n] = g[m, n] * dropmask(m, n)/(1-p) * (act[m, n] > 0 if relu)
// g: gradient wrt the layer's (dropped) output, fp32 [M, g_stride] or bf16; act: the layer's bf16 output.
// Writes dz bf16 [M, Np] (zero padded), its transpose dzT bf16 [Np, Mp] and accumulates the bias gradient
// db[n] += sum_m dz[m, n] (db pre-zeroed).  One block = 32 rows x 32 columns (transposed through smem).
__global__ void __launch_bounds__(256) mlp_bwd_elementwise_kernel(
    const void* __restrict__ g, int g_bf16, int64_t g_stride, const __nv_bfloat16* __restrict__ act,
    int64_t act_stride, int relu, float p, uint64_t seed, uint32_t tag, int64_t m, int64_t n, int64_t np,
    int64_t mp, __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dzt, float* __restrict__ db) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t col = int64_t(blockIdx.x) * 32 + tx;
  const uint32_t thresh = drop_threshold(p);
  const float sc = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  float colsum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t row = int64_t(blockIdx.y) * 32 + ty * 4 + k;
    float v = 0.f;
    if (row < m && col < n) {
      v = g_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(g)[row * g_stride + col])
                 : static_cast<const float*>(g)[row * g_stride + col];
      if (relu && !(__bfloat162float(act[row * act_stride + col]) > 0.f)) v = 0.f;
      if (p > 0.f) {
        const uint32_t keep = dropout_keep4(seed, tag, uint32_t(row), uint32_t(col >> 2), thresh);
        v = ((keep >> (col & 3)) & 1u) ? v * sc : 0.f;
      }
    }
    v = __bfloat162float(__float2bfloat16_rn(v));  // db sums exactly what the GEMMs will see
    if (row < m && col < np) dz[row * np + col] = __float2bfloat16_rn(v);
    tile[ty * 4 + k][tx] = v;
    colsum += v;
  }
  if (db != nullptr && col < n) atomicAdd(db + col, colsum);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t orow = int64_t(blockIdx.x) * 32 + ty * 4 + k;  // a column of dz
    const int64_t ocol = int64_t(blockIdx.y) * 32 + tx;          // a row of dz
    if (dzt != nullptr && orow < np && ocol < mp) dzt[orow * mp + ocol] = __float2bfloat16_rn(tile[tx][ty * 4 + k]);
  }
}

// Same step for bf16 gradients with 16-byte accesses: one block = 64 rows x 64 columns, transposed through shared
// memory (33-word rows: conflict-free writes, 2-way reads).  Needs g_stride, act_stride, np, mp multiples of 8.
constexpr int MBV_RT = 16;   // 64-row tiles per block: the bias-gradient atomics are issued once per block, not per tile
__global__ void __launch_bounds__(256) mlp_bwd_elementwise_vec_kernel(
    const __nv_bfloat16* __restrict__ g, int64_t g_stride, const __nv_bfloat16* __restrict__ act, int64_t act_stride,
    int relu, float p, uint64_t seed, uint32_t tag, int64_t m, int64_t n, int64_t np, int64_t mp,
    __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dzt, float* __restrict__ db) {
  __shared__ __align__(16) uint16_t tile[64][66];
  const int64_t c0 = int64_t(blockIdx.y) * 64;   // rows on x: up to 2^31 blocks
  const uint32_t thresh = drop_threshold(p);
  const float sc = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  float colsum[2] = {0.f, 0.f};   // this thread's (column, row group) partial of db, over all row tiles of the block
  for (int rt = 0; rt < MBV_RT; ++rt) {
    const int64_t r0 = (int64_t(blockIdx.x) * MBV_RT + rt) * 64;
    if (r0 >= mp) break;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = threadIdx.x + it * 256;   // 64 rows x 8 column groups
      const int rl = idx >> 3, cg = idx & 7;
      const int64_t row = r0 + rl, col = c0 + cg * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (row < m && col < np) {
        if (col < n) {   // (the last group of a row may straddle n: masked per element below)
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(g + row * g_stride + col));
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
          if (relu) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(act + row * act_stride + col));
            const float2 e = unpack_bf16x2(w.x), f = unpack_bf16x2(w.y), h = unpack_bf16x2(w.z), k = unpack_bf16x2(w.w);
            const float av[8] = {e.x, e.y, f.x, f.y, h.x, h.y, k.x, k.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = av[j] > 0.f ? v[j] : 0.f;
          }
          if (p > 0.f) {
            const uint32_t k8 = dropout_keep8(seed, tag, uint32_t(row), uint32_t(col >> 3), thresh);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ((k8 >> j) & 1u) ? v[j] * sc : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (col + j < n) ? v[j] : 0.f;
        }
        o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                       pack_bf16x2(v[6], v[7]));
        *reinterpret_cast<uint4*>(dz + row * np + col) = o;
      }
      uint32_t* trow = reinterpret_cast<uint32_t*>(&tile[rl][cg * 8]);
      trow[0] = o.x; trow[1] = o.y; trow[2] = o.z; trow[3] = o.w;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = threadIdx.x + it * 256;   // 64 columns x 8 row groups
      const int cl = idx >> 3, rg = idx & 7;
      const int64_t col = c0 + cl, row = r0 + rg * 8;
      uint16_t e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        e[j] = tile[rg * 8 + j][cl];
        colsum[it] += __uint_as_float(uint32_t(e[j]) << 16);   // db sums exactly what the GEMMs will see
      }
      if (dzt != nullptr && col < np && row < mp)
        *reinterpret_cast<uint4*>(dzt + col * mp + row) =
            make_uint4(uint32_t(e[0]) | (uint32_t(e[1]) << 16), uint32_t(e[2]) | (uint32_t(e[3]) << 16),
                       uint32_t(e[4]) | (uint32_t(e[5]) << 16), uint32_t(e[6]) | (uint32_t(e[7]) << 16));
    }
    __syncthreads();   // the tile is rewritten by the next row tile
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * 256;
    const int cl = idx >> 3, rg = idx & 7;
    const int64_t col = c0 + cl;
    float cs = colsum[it];   // the 8 row groups of a column sit in 8 consecutive lanes
    cs += __shfl_xor_sync(0xffffffffu, cs, 1);
    cs += __shfl_xor_sync(0xffffffffu, cs, 2);
    cs += __shfl_xor_sync(0xffffffffu, cs, 4);
    if (db != nullptr && rg == 0 && col < n) atomicAdd(db + col, cs);
  }
}

// bf16 [rows, cols] (row stride in_stride) -> bf16 [cols, rows_padded] (zero padded rows beyond `rows`)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                             int64_t in_stride, int64_t rows, int64_t cols,
                                                             int64_t rows_padded,
                                                             __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = int64_t(blockIdx.y) * 32 + ty * 4 + k, c = int64_t(blockIdx.x) * 32 + tx;
    tile[ty * 4 + k][tx] = (r < rows && c < cols) ? in[r * in_stride + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t orow = int64_t(blockIdx.x) * 32 + ty * 4 + k, ocol = int64_t(blockIdx.y) * 32 + tx;
    if (orow < cols && ocol < rows_padded) out[orow * rows_padded + ocol] = tile[tx][ty * 4 + k];
  }
}

// fp32 W [n, k] -> bf16 W^T [k_padded, n_padded] (zero padded): the dgrad operand
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ in, int64_t n, int64_t k,
                                                             int64_t kp, int64_t np,
                                                             __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t r = int64_t(blockIdx.y) * 32 + ty * 4 + j, c = int64_t(blockIdx.x) * 32 + tx;
    tile[ty * 4 + j][tx] = (r < n && c < k) ? __ldg(in + r * k + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t orow = int64_t(blockIdx.x) * 32 + ty * 4 + j, ocol = int64_t(blockIdx.y) * 32 + tx;
    if (orow < kp && ocol < np) out[orow * np + ocol] = __float2bfloat16_rn(tile[tx][ty * 4 + j]);
  }
}

}  // namespace mmbs

extern "C" int mmbs_dropout_cast_bf16(const void* in, int32_t in_is_bf16, int64_t in_stride, void* out, int64_t rows,
                                      int64_t cols, int64_t cols_padded, float p, uint64_t seed, uint32_t tag,
                                      void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && rows > 0 && cols > 0 && cols_padded >= cols && cols_padded % 4 == 0 && p >= 0.f && p < 1.f,
               "mmbs_dropout_cast_bf16: bad argument");
  const bool vec = cols % 8 == 0 && cols_padded % 8 == 0 && in_stride % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  if (vec) {
    dropout_cast_vec8_kernel<<<blocks_for(rows * (cols_padded / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, in_is_bf16, in_stride, static_cast<uint4*>(out), rows, cols, cols_padded / 8, p, seed, tag);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  dropout_cast_kernel<<<blocks_for(rows * (cols_padded / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_is_bf16, in_stride, static_cast<uint2*>(out), rows, cols, cols_padded / 4, p, seed, tag);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_mlp_bwd_elementwise(const void* g, int32_t g_is_bf16, int64_t g_stride, const void* act,
                                        int64_t act_stride, int32_t relu, float p, uint64_t seed, uint32_t tag,
                                        int64_t m, int64_t n, int64_t n_padded, int64_t m_padded, void* dz, void* dzt,
                                        float* db, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && dz && m > 0 && n > 0 && n_padded >= n && m_padded >= m && (!relu || act) && p >= 0.f && p < 1.f,
               "mmbs_mlp_bwd_elementwise: bad argument");   // dzt may be NULL (TN weight-gradient GEMMs read dz itself)
  const bool vec = g_is_bf16 && g_stride % 8 == 0 && (!relu || act_stride % 8 == 0) && n_padded % 8 == 0 &&
                   m_padded % 8 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
                   (!relu || reinterpret_cast<uintptr_t>(act) % 16 == 0) && reinterpret_cast<uintptr_t>(dz) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(dzt) % 16 == 0 && ceil_div(n_padded, 64) <= 65535;
  if (vec) {
    // rows beyond m up to m_padded: the transposed output is written (zeros) for every row group the grid covers
    dim3 vgrid(unsigned(ceil_div(ceil_div(m_padded, 64), MBV_RT)), unsigned(ceil_div(n_padded, 64)));
    mlp_bwd_elementwise_vec_kernel<<<vgrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(g), g_stride, static_cast<const __nv_bfloat16*>(act), act_stride, relu, p, seed,
        tag, m, n, n_padded, m_padded, static_cast<__nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(dzt), db);
    MMBS_LAUNCH_CHECK();
    return MMBS_OK;
  }
  dim3 grid(unsigned(ceil_div(n_padded, 32)), unsigned(ceil_div(m_padded, 32)));
  mlp_bwd_elementwise_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, g_is_bf16, g_stride, static_cast<const __nv_bfloat16*>(act), act_stride, relu, p, seed, tag, m, n, n_padded,
      m_padded, static_cast<__nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(dzt), db);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_transpose_bf16(const void* in, int64_t in_stride, int64_t rows, int64_t cols, int64_t rows_padded,
                                   void* out, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && rows > 0 && cols > 0 && rows_padded >= rows, "mmbs_transpose_bf16: bad argument");
  if (in_stride == cols && cols % 8 == 0 && rows_padded % 8 == 0 && rows < (int64_t(1) << 31) &&
      reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
      ceil_div(cols, 64) <= 65535) {
    // csrc/train.cu: 64x64 shared-memory transpose with 16-byte global accesses (1x1 "im2col" = plain transpose)
    return mmbs_im2col_t(in, out, rows, 1, 1, cols, 1, 1, rows_padded, 0, stream);
  }
  dim3 grid(unsigned(ceil_div(cols, 32)), unsigned(ceil_div(rows_padded, 32)));
  transpose_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), in_stride, rows, cols, rows_padded, static_cast<__nv_bfloat16*>(out));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_cast_transpose_pad_bf16(const float* in, int64_t n, int64_t k, int64_t k_padded, int64_t n_padded,
                                            void* out, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && out && n > 0 && k > 0 && k_padded >= k && n_padded >= n, "mmbs_cast_transpose_pad_bf16: bad argument");
  dim3 grid(unsigned(ceil_div(k_padded, 32)), unsigned(ceil_div(n_padded, 32)));
  cast_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, k, k_padded, n_padded,
                                                                            static_cast<__nv_bfloat16*>(out));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
