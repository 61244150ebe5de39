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
__global__ void __launch_bounds__(256) maxpool_3x3s2_kernel(const uint4* __restrict__ in,
                                                            uint4* __restrict__ out, int64_t batch, int h,
                                                            int w, int c8) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const int64_t total = batch * oh * ow * c8;
  if (i >= total) return;
  const int g = int(i % c8);
  const int x = int((i / c8) % ow), y = int((i / (int64_t(c8) * ow)) % oh);
  const int64_t n = i / (int64_t(c8) * ow * oh);
  const uint32_t NEG = 0xff80ff80u;  // (-inf, -inf) in bf16
  uint4 m = make_uint4(NEG, NEG, NEG, NEG);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = 2 * y - 1 + dy;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = 2 * x - 1 + dx;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = __ldg(in + ((n * h + iy) * w + ix) * c8 + g);
      m.x = max_bf16x2(m.x, v.x); m.y = max_bf16x2(m.y, v.y);
      m.z = max_bf16x2(m.z, v.z); m.w = max_bf16x2(m.w, v.w);
    }
  }
  out[i] = m;
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
  pack_conv_weight_kernel<<<blocks_for(c_out * c_in * k * k, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(out), c_out, c_in, k);
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
  maxpool_3x3s2_kernel<<<blocks_for(batch * oh * ow * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), static_cast<uint4*>(out), batch, int(h), int(w), int(c / 8));
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
