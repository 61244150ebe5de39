// Training-mode ResNet trunk glue (sm_100a, HBM-bound): batch-statistics BatchNorm forward and
// backward, the operand layouts of the layer4 dgrad / wgrad GEMMs, gradient routing.
//
// Reference semantics: Bottleneck.forward in model.train()
//   /root/reference/5_JointFusion/resnet.py:70-90 (nn.BatchNorm2d in training mode: batch mean /
//   biased variance for the normalisation, momentum 0.1 update of running_mean / unbiased
//   running_var), autograd through layer4 + fc as configured by
//   /root/reference/1_HistoPathology/2_HistoPath_train.py:541-551 (n_layers_to_train).
// The per-channel sums come out of the conv kernel's epilogue (gemm_tcgen05.cu, ConvParams::stats).
#include <algorithm>
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"

namespace mmbs {

constexpr int BNA_ITER = 8;   // 8-channel items per thread in the BatchNorm apply kernels

__device__ __forceinline__ uint32_t tr_pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void tr_unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 tr_pack8(const float* f) {
  return make_uint4(tr_pack2(f[0], f[1]), tr_pack2(f[2], f[3]), tr_pack2(f[4], f[5]), tr_pack2(f[6], f[7]));
}
__device__ __forceinline__ void tr_ld8(const float* p, int g, float* v) {
  *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(p) + 2 * g);
  *reinterpret_cast<float4*>(v + 4) = __ldg(reinterpret_cast<const float4*>(p) + 2 * g + 1);
}

// ---- BatchNorm (training) finalize: sums -> scale/shift for the apply pass, saved mean / invstd for the
// backward pass, running-statistics update exactly like nn.BatchNorm2d (momentum, unbiased variance).
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int64_t c, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double mean = double(stats[i]) / count;
  double var = double(stats[c + i]) / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = float(1.0 / sqrt(var + double(eps)));
  const float g = gamma ? gamma[i] : 1.0f, b = beta ? beta[i] : 0.0f;
  const float sc = g * invstd;
  scale[i] = sc;
  shift[i] = b - float(mean) * sc;
  mean_out[i] = float(mean);
  invstd_out[i] = invstd;
  if (running_mean) running_mean[i] = (1.0f - momentum) * running_mean[i] + momentum * float(mean);
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[i] = (1.0f - momentum) * running_var[i] + momentum * float(unbiased);
  }
}

// ---- y = [relu]( x*scale + shift  (+ r)  |  (+ r*rscale + rshift) ), bf16 NHWC, 8 channels per thread
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const uint4* __restrict__ r,
                                                       const float* __restrict__ rscale,
                                                       const float* __restrict__ rshift, int relu,
                                                       uint4* __restrict__ out, int64_t total, int c8) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int g = int(i % c8);
  float v[8], sc[8], sh[8];
  tr_unpack8(x[i], v);
  *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g);
  *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g + 1);
  *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g);
  *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g + 1);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
  if (r != nullptr) {
    float rv[8];
    tr_unpack8(r[i], rv);
    if (rscale != nullptr) {
      *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(rscale) + 2 * g);
      *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(rscale) + 2 * g + 1);
      *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(rshift) + 2 * g);
      *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(rshift) + 2 * g + 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) rv[j] = fmaf(rv[j], sc[j], sh[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += rv[j];
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  out[i] = tr_pack8(v);
}

// ---- stem tail in training mode: MaxPool2d(3,2,1)(relu(x*scale + shift)); grid = (ceil(ow*c8/256), oh, B)
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(const uint4* __restrict__ in,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              uint4* __restrict__ out, int h, int w, int c8,
                                                              int c8_shift) {
  const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const int xc = blockIdx.x * 256 + threadIdx.x;
  if (xc >= ow * c8) return;
  const int g = xc & (c8 - 1), x = xc >> c8_shift;
  const int y = blockIdx.y;
  const int64_t n = blockIdx.z;
  float sc[8], sh[8], m[8];
  *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g);
  *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g + 1);
  *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g);
  *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g + 1);
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = 0.f;   // relu(.) >= 0 and every window holds an in-range pixel
  const uint4* base = in + n * int64_t(h) * w * c8 + g;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = 2 * y - 1 + dy;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = 2 * x - 1 + dx;
      if (ix < 0 || ix >= w) continue;
      float v[8];
      tr_unpack8(__ldg(base + (int64_t(iy) * w + ix) * c8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaf(v[j], sc[j], sh[j]));
    }
  }
  out[((n * oh + y) * int64_t(ow) + x) * c8 + g] = tr_pack8(m);
}

// ---- fused finalize + apply (training forward): the per-channel scale/shift are recomputed by every thread
// from the epilogue sums (a handful of FMAs + one rsqrt per channel, free next to the 32-48 bytes the thread
// moves), so no separate finalize launch sits between the convolution and its normalisation.  The first c8
// threads of the grid also publish mean / invstd / scale / shift for the backward pass and update the running
// statistics.  Launched with programmatic dependent launch: the prologue overlaps the tail of the convolution.
struct BnTrain {
  const float* stats;      // [2][c] sum | sum of squares from the conv epilogue
  const float* gamma;
  const float* beta;
  float* running_mean;     // may be null
  float* running_var;
  float* scale_out;        // [c] each, for the backward pass
  float* shift_out;
  float* mean_out;
  float* invstd_out;
  float inv_count, unbias, eps, momentum;
  int c;
};

__device__ __forceinline__ void bn_train_coeffs(const BnTrain& b, int g, bool publish, float* sc, float* sh) {
  float s1[8], s2[8], ga[8], be[8];
  tr_ld8(b.stats, g, s1);
  tr_ld8(b.stats + b.c, g, s2);
  tr_ld8(b.gamma, g, ga);
  tr_ld8(b.beta, g, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = s1[j] * b.inv_count;
    const float var = fmaxf(fmaf(-mean, mean, s2[j] * b.inv_count), 0.f);
    const float invstd = rsqrtf(var + b.eps);
    sc[j] = ga[j] * invstd;
    sh[j] = fmaf(-mean, sc[j], be[j]);
    if (publish) {
      const int ch = g * 8 + j;
      b.scale_out[ch] = sc[j];
      b.shift_out[ch] = sh[j];
      b.mean_out[ch] = mean;
      b.invstd_out[ch] = invstd;
      if (b.running_mean) b.running_mean[ch] = (1.0f - b.momentum) * b.running_mean[ch] + b.momentum * mean;
      if (b.running_var) b.running_var[ch] = (1.0f - b.momentum) * b.running_var[ch] + b.momentum * var * b.unbias;
    }
  }
}

// res_mode: 0 none, 1 identity residual, 2 residual normalised with its own BatchNorm (downsample branch).
// A block covers BNA_ITER * 256 consecutive 8-channel items; c8 divides 256, so every item of a thread belongs to
// the same channel group and the coefficients are derived once per thread (the per-item parameter loads of a naive
// kernel move 4x more L1 bytes than the payload moves HBM bytes and bound it).
__global__ void __launch_bounds__(256) bn_train_apply_kernel(const __grid_constant__ BnTrain bn,
                                                             const __grid_constant__ BnTrain rbn,
                                                             const uint4* __restrict__ x, const uint4* __restrict__ r,
                                                             int res_mode, int relu, uint4* __restrict__ out,
                                                             int64_t total, int c8) {
  // (no early launch_dependents: a persistent conv CTA that became resident while this HBM-bound kernel is still
  //  running would take 46 K registers per SM away from it)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t base = int64_t(blockIdx.x) * (BNA_ITER * 256) + threadIdx.x;
  if (base >= total) return;
  const int g = int(base & (c8 - 1));
  const bool publish = base < c8;
  float sc[8], sh[8], rsc[8], rsh[8];
  bn_train_coeffs(bn, g, publish, sc, sh);
  if (res_mode == 2) bn_train_coeffs(rbn, g, publish, rsc, rsh);
  uint4 xv[BNA_ITER], rv4[BNA_ITER];
#pragma unroll
  for (int k = 0; k < BNA_ITER; ++k) {
    const int64_t i = base + int64_t(k) * 256;
    if (i < total) {
      xv[k] = x[i];
      if (res_mode != 0) rv4[k] = __ldg(r + i);
    }
  }
#pragma unroll
  for (int k = 0; k < BNA_ITER; ++k) {
    const int64_t i = base + int64_t(k) * 256;
    if (i >= total) break;
    float v[8];
    tr_unpack8(xv[k], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (res_mode != 0) {
      float rv[8];
      tr_unpack8(rv4[k], rv);
      if (res_mode == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rv[j] = fmaf(rv[j], rsc[j], rsh[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rv[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    out[i] = tr_pack8(v);
  }
}

constexpr int TMP_ROWS = 4;   // vertically adjacent outputs per thread (each input row is normalised and reduced once)
__global__ void __launch_bounds__(256) bn_train_relu_maxpool_kernel(const __grid_constant__ BnTrain bn,
                                                                    const uint4* __restrict__ in,
                                                                    uint4* __restrict__ out, int h, int w, int c8,
                                                                    int c8_shift) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const int xc = blockIdx.x * 256 + threadIdx.x;
  if (xc >= ow * c8) return;
  const int g = xc & (c8 - 1), x = xc >> c8_shift;
  const int y0 = blockIdx.y * TMP_ROWS;
  const int64_t n = blockIdx.z;
  float sc[8], sh[8];
  bn_train_coeffs(bn, g, (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && xc < c8, sc, sh);
  float m[TMP_ROWS][8];
#pragma unroll
  for (int k = 0; k < TMP_ROWS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) m[k][j] = 0.f;   // relu(.) >= 0 and every window holds an in-range pixel
  const uint4* base = in + n * int64_t(h) * w * c8 + g;
#pragma unroll
  for (int r = 0; r < 2 * TMP_ROWS + 1; ++r) {
    const int iy = 2 * y0 - 1 + r;
    float rm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rm[j] = 0.f;
    if (iy >= 0 && iy < h) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = 2 * x - 1 + dx;
        if (ix < 0 || ix >= w) continue;
        float v[8];
        tr_unpack8(__ldg(base + (int64_t(iy) * w + ix) * c8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) rm[j] = fmaxf(rm[j], fmaf(v[j], sc[j], sh[j]));
      }
    }
#pragma unroll
    for (int k = 0; k < TMP_ROWS; ++k) {
      if (r >= 2 * k && r <= 2 * k + 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) m[k][j] = fmaxf(m[k][j], rm[j]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < TMP_ROWS; ++k)
    if (y0 + k < oh) out[((n * oh + y0 + k) * int64_t(ow) + x) * c8 + g] = tr_pack8(m[k]);
}

// ---- grad of AvgPool2d(7)+flatten: dfeat fp32 [B, C] -> bf16 [B, hw, C] = dfeat / hw
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dfeat, uint4* __restrict__ out,
                                                          int64_t total, int hw, int c8) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int g = int(i % c8);
  const int64_t n = i / (int64_t(c8) * hw);
  const float inv = 1.0f / float(hw);
  float v[8];
  *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(dfeat + (n * c8 + g) * 8));
  *reinterpret_cast<float4*>(v + 4) = __ldg(reinterpret_cast<const float4*>(dfeat + (n * c8 + g) * 8) + 1);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= inv;
  out[i] = tr_pack8(v);
}

// ---- BatchNorm backward, pass 1: per channel  s1 = sum dz,  s2 = sum dz * xhat   with
// dz = g * (mask > 0 if mask given), xhat = (raw - mean) * invstd.  block (32, 8): x = 8-channel group,
// y = row lane; sums[2][C] pre-zeroed.
constexpr int BNB_ROWS = 64;
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const uint4* __restrict__ g, const uint4* __restrict__ mask,
                                                            const uint4* __restrict__ raw,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            float* __restrict__ sums, int64_t rows, int c8) {
  __shared__ float red[8][32][17];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cg = blockIdx.x * 32 + tx;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (cg < c8) {
    float mu[8], is[8];
    *reinterpret_cast<float4*>(mu) = __ldg(reinterpret_cast<const float4*>(mean) + 2 * cg);
    *reinterpret_cast<float4*>(mu + 4) = __ldg(reinterpret_cast<const float4*>(mean) + 2 * cg + 1);
    *reinterpret_cast<float4*>(is) = __ldg(reinterpret_cast<const float4*>(invstd) + 2 * cg);
    *reinterpret_cast<float4*>(is + 4) = __ldg(reinterpret_cast<const float4*>(invstd) + 2 * cg + 1);
    const int64_t r0 = int64_t(blockIdx.y) * BNB_ROWS;
    const int64_t r1 = min(rows, r0 + BNB_ROWS);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float gv[8], xv[8];
      tr_unpack8(__ldg(g + r * c8 + cg), gv);
      tr_unpack8(__ldg(raw + r * c8 + cg), xv);
      if (mask != nullptr) {
        float mv[8];
        tr_unpack8(__ldg(mask + r * c8 + cg), mv);
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] = mv[j] > 0.f ? gv[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += gv[j];
        s2[j] = fmaf(gv[j], (xv[j] - mu[j]) * is[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[ty][tx][j] = s1[j];
    red[ty][tx][8 + j] = s2[j];
  }
  __syncthreads();
  // 32 channel groups x 16 values = 512 sums, 256 threads take two each
  for (int o = threadIdx.x; o < 512; o += 256) {
    const int gx = o >> 4, j = o & 15;
    float acc = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) acc += red[y][gx][j];
    const int c = (blockIdx.x * 32 + gx) * 8 + (j & 7);
    if (c < c8 * 8) atomicAdd(sums + (j >> 3) * (int64_t(c8) * 8) + c, acc);
  }
}

// ---- BatchNorm backward, pass 2: draw = scale * (dz - s1/n - xhat * s2/n), scale = gamma * invstd, i.e.
// draw = A*dz + Bx*raw + C with per-channel A = scale, Bx = -scale*invstd*s2/n, C = -scale*s1/n - Bx*mean.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ mask,
                                                           const uint4* __restrict__ raw,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ sums, float inv_count,
                                                           uint4* __restrict__ out, int64_t total, int c8) {
  // c8 divides 256: all items of a thread share one channel group, the coefficients are derived once
  const int64_t base = int64_t(blockIdx.x) * (BNA_ITER * 256) + threadIdx.x;
  if (base >= total) return;
  const int cg = int(base & (c8 - 1));
  float a[8], bx[8], c0[8];
  {
    float mu[8], is[8], sc[8], s1[8], s2[8];
    tr_ld8(mean, cg, mu);
    tr_ld8(invstd, cg, is);
    tr_ld8(scale, cg, sc);
    tr_ld8(sums, cg, s1);
    tr_ld8(sums + int64_t(c8) * 8, cg, s2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = sc[j];
      bx[j] = -sc[j] * is[j] * s2[j] * inv_count;
      c0[j] = -sc[j] * s1[j] * inv_count - bx[j] * mu[j];
    }
  }
  uint4 gv4[BNA_ITER], xv4[BNA_ITER], mv4[BNA_ITER];
#pragma unroll
  for (int k = 0; k < BNA_ITER; ++k) {
    const int64_t i = base + int64_t(k) * 256;
    if (i < total) {
      gv4[k] = __ldg(g + i);
      xv4[k] = __ldg(raw + i);
      if (mask != nullptr) mv4[k] = __ldg(mask + i);
    }
  }
#pragma unroll
  for (int k = 0; k < BNA_ITER; ++k) {
    const int64_t i = base + int64_t(k) * 256;
    if (i >= total) break;
    float gv[8], xv[8], o[8];
    tr_unpack8(gv4[k], gv);
    tr_unpack8(xv4[k], xv);
    if (mask != nullptr) {
      float mv[8];
      tr_unpack8(mv4[k], mv);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = mv[j] > 0.f ? gv[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j], gv[j], fmaf(bx[j], xv[j], c0[j]));
    out[i] = tr_pack8(o);
  }
}

// ---- wgrad B operand: x NHWC bf16 [B,H,W,C] -> colT [k*k*C, Pp] bf16,
// colT[(kh*k+kw)*C + c][p] = x[n, ho*s + kh - pad, wo*s + kw - pad, c] (0 outside / for p >= P), p = (n, ho, wo).
// One block = 64 channels x 64 pixels of one tap, transposed through shared memory with 16-byte global
// accesses on both sides (8 channels of a pixel in, 8 pixels of a channel out).  C % 8 == 0, Pp % 8 == 0.
// row order: (tap, channel) [oihw_rows = 0: matches the packed weight layout] or (channel, tap) [oihw_rows = 1: the
// weight-gradient GEMM then produces nn.Conv2d.weight.grad's OIHW layout directly].
__global__ void __launch_bounds__(256) im2col_t_kernel(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ out, int batch, int h, int w,
                                                       int c, int k, int stride, int pad, int oh, int ow,
                                                       int64_t p_total, int64_t p_padded, int oihw_rows) {
  __shared__ __align__(16) uint16_t tile[64][66];   // [pixel][channel], 33-word rows: conflict-free writes, 2-way reads
  const int tap = blockIdx.z, kh = tap / k, kw = tap % k;
  const int c0 = blockIdx.y * 64;
  const int64_t p0 = int64_t(blockIdx.x) * 64;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * 256;   // 64 pixels x 8 channel groups
    const int pl = idx >> 3, cgp = idx & 7;
    const int64_t p = p0 + pl;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p < p_total && c0 + cgp * 8 < c) {
      const int wo = int(p % ow), ho = int((p / ow) % oh);
      const int64_t n = p / (int64_t(ow) * oh);
      const int iy = ho * stride + kh - pad, ix = wo * stride + kw - pad;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = __ldg(reinterpret_cast<const uint4*>(x + ((n * h + iy) * int64_t(w) + ix) * c + c0 + cgp * 8));
    }
    uint32_t* trow = reinterpret_cast<uint32_t*>(&tile[pl][cgp * 8]);
    trow[0] = v.x; trow[1] = v.y; trow[2] = v.z; trow[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * 256;   // 64 channels x 8 pixel groups
    const int cl = idx >> 3, pg = idx & 7;
    const int ch = c0 + cl;
    const int64_t p = p0 + pg * 8;
    if (ch < c && p < p_padded) {
      uint16_t e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] = tile[pg * 8 + j][cl];
      const uint4 v = make_uint4(uint32_t(e[0]) | (uint32_t(e[1]) << 16), uint32_t(e[2]) | (uint32_t(e[3]) << 16),
                                 uint32_t(e[4]) | (uint32_t(e[5]) << 16), uint32_t(e[6]) | (uint32_t(e[7]) << 16));
      const int64_t orow = oihw_rows ? (int64_t(ch) * (k * k) + tap) : (int64_t(tap) * c + ch);
      *reinterpret_cast<uint4*>(out + orow * p_padded + p) = v;
    }
  }
}

// ---- dgrad weights: OIHW fp32 [O,I,k,k] -> bf16 [I][k][k][O] with the taps flipped
// (out[i][kh][kw][o] = w[o][i][k-1-kh][k-1-kw]); k = 1: the transposed matrix.
__global__ void pack_conv_weight_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                              int64_t c_out, int64_t c_in, int64_t k) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = c_out * k * k * c_in;
  if (i >= total) return;
  const int64_t o = i % c_out, kw = (i / c_out) % k, kh = (i / (c_out * k)) % k, ci = i / (c_out * k * k);
  out[i] = __float2bfloat16_rn(w[((o * c_in + ci) * k + (k - 1 - kh)) * k + (k - 1 - kw)]);
}

// k = 3 variant: one thread per (ci, o), o fastest: 9 writes coalesced over o, one 36-byte read
__global__ void __launch_bounds__(256) pack_conv_weight_dgrad3_kernel(const float* __restrict__ w,
                                                                      __nv_bfloat16* __restrict__ out, int64_t c_out,
                                                                      int64_t c_in) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= c_out * c_in) return;
  const int64_t o = i % c_out, ci = i / c_out;
  const float* src = w + (o * c_in + ci) * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) out[(ci * 9 + t) * c_out + o] = __float2bfloat16_rn(__ldg(src + 8 - t));
}

// ---- wgrad result [O][kh][kw][I] fp32 -> OIHW fp32 (the layout of nn.Conv2d.weight.grad)
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t c_out,
                                         int64_t c_in, int64_t k) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = c_out * k * k * c_in;
  if (i >= total) return;
  const int64_t kw = i % k, kh = (i / k) % k, ci = (i / (k * k)) % c_in, o = i / (k * k * c_in);
  out[i] = g[((o * k + kh) * k + kw) * c_in + ci];
}

// ---- stride-2 dgrad input: u[n, 2y, 2x, :] = g[n, y, x, :]  (u zero elsewhere: zeroed once by the caller)
__global__ void __launch_bounds__(256) scatter_stride2_kernel(const uint4* __restrict__ g, uint4* __restrict__ u,
                                                              int64_t total, int h, int w, int c8) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int cg = int(i % c8);
  const int64_t pix = i / c8;
  const int x = int(pix % w), y = int((pix / w) % h);
  const int64_t n = pix / (int64_t(w) * h);
  u[((n * (2 * h) + 2 * y) * int64_t(2 * w) + 2 * x) * c8 + cg] = g[i];
}

// ---- out = a + g * (mask > 0)   (gradient into a block input: conv path + identity shortcut)
__global__ void __launch_bounds__(256) add_relu_mask_kernel(const uint4* __restrict__ a, const uint4* __restrict__ g,
                                                            const uint4* __restrict__ mask, uint4* __restrict__ out,
                                                            int64_t total) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  float av[8], gv[8], mv[8];
  tr_unpack8(__ldg(g + i), gv);
  tr_unpack8(__ldg(mask + i), mv);
  if (a != nullptr) tr_unpack8(__ldg(a + i), av);
#pragma unroll
  for (int j = 0; j < 8; ++j) gv[j] = (mv[j] > 0.f ? gv[j] : 0.f) + (a != nullptr ? av[j] : 0.f);
  out[i] = tr_pack8(gv);
}

}  // namespace mmbs

using namespace mmbs;

static inline unsigned tr_blocks(int64_t total, int threads) {
  return unsigned(std::max<int64_t>(1, ceil_div(total, threads)));
}
static inline bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

extern "C" int mmbs_bn_finalize(const float* stats, int64_t c, int64_t count, const float* gamma, const float* beta,
                                float eps, float momentum, float* running_mean, float* running_var, float* scale,
                                float* shift, float* mean_out, float* invstd_out, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(stats && scale && shift && mean_out && invstd_out && c > 0 && count > 0, "mmbs_bn_finalize: bad argument");
  bn_finalize_kernel<<<tr_blocks(c, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, c, double(count), gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean_out,
      invstd_out);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_bn_apply(const void* x, const float* scale, const float* shift, const void* residual,
                             const float* res_scale, const float* res_shift, int32_t relu, void* out, int64_t rows,
                             int64_t c, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x && scale && shift && out && rows > 0 && c > 0 && c % 8 == 0 && al16(x) && al16(out) &&
                   al16(residual) && al16(scale) && al16(shift) && (!res_scale == !res_shift) &&
                   (!res_scale || residual),
               "mmbs_bn_apply: bad argument");
  const int64_t total = rows * (c / 8);
  bn_apply_kernel<<<tr_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), scale, shift, static_cast<const uint4*>(residual), res_scale, res_shift, relu,
      static_cast<uint4*>(out), total, int(c / 8));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_bn_relu_maxpool_3x3s2(const void* in, const float* scale, const float* shift, void* out,
                                          int64_t batch, int64_t h, int64_t w, int64_t c, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in && scale && shift && out && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0,
               "mmbs_bn_relu_maxpool_3x3s2: bad argument");
  const int64_t oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const int c8 = int(c / 8);
  int shift_bits = 0;
  while ((1 << shift_bits) < c8) ++shift_bits;
  MMBS_REQUIRE((1 << shift_bits) == c8 && batch <= 65535 && oh <= 65535,
               "mmbs_bn_relu_maxpool_3x3s2: c/8 must be a power of two");
  dim3 grid(unsigned(ceil_div(ow * c8, 256)), unsigned(oh), unsigned(batch));
  bn_relu_maxpool_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), scale, shift, static_cast<uint4*>(out), int(h), int(w), c8, shift_bits);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_avgpool_global_bwd(const float* dfeat, void* out, int64_t batch, int64_t hw, int64_t c,
                                       void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(dfeat && out && batch > 0 && hw > 0 && c > 0 && c % 8 == 0 && al16(dfeat) && al16(out),
               "mmbs_avgpool_global_bwd: bad argument");
  const int64_t total = batch * hw * (c / 8);
  avgpool_bwd_kernel<<<tr_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dfeat, static_cast<uint4*>(out), total, int(hw), int(c / 8));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_bn_bwd_reduce(const void* g, const void* relu_mask, const void* raw, const float* mean,
                                  const float* invstd, float* sums, int64_t rows, int64_t c, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && raw && mean && invstd && sums && rows > 0 && c > 0 && c % 8 == 0 && al16(g) && al16(raw) &&
                   al16(relu_mask) && al16(mean) && al16(invstd),
               "mmbs_bn_bwd_reduce: bad argument");
  const int c8 = int(c / 8);
  dim3 grid(unsigned(ceil_div(c8, 32)), unsigned(ceil_div(rows, BNB_ROWS)));
  bn_bwd_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g), static_cast<const uint4*>(relu_mask), static_cast<const uint4*>(raw), mean, invstd,
      sums, rows, c8);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_bn_bwd_apply(const void* g, const void* relu_mask, const void* raw, const float* mean,
                                 const float* invstd, const float* scale, const float* sums, void* out, int64_t rows,
                                 int64_t c, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && raw && mean && invstd && scale && sums && out && rows > 0 && c > 0 && c % 8 == 0 && al16(g) &&
                   al16(raw) && al16(relu_mask) && al16(out),
               "mmbs_bn_bwd_apply: bad argument");
  const int64_t total = rows * (c / 8);
  MMBS_REQUIRE(c / 8 <= 256 && ((c / 8) & (c / 8 - 1)) == 0, "mmbs_bn_bwd_apply: c/8 must be a power of two <= 256");
  bn_bwd_apply_kernel<<<tr_blocks(total, 256 * BNA_ITER), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g), static_cast<const uint4*>(relu_mask), static_cast<const uint4*>(raw), mean, invstd,
      scale, sums, 1.0f / float(rows), static_cast<uint4*>(out), total, int(c / 8));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_im2col_t(const void* x, void* out, int64_t batch, int64_t h, int64_t w, int64_t c, int64_t ksize,
                             int64_t stride, int64_t p_padded, int32_t oihw_rows, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x && out && batch > 0 && h > 0 && w > 0 && c > 0 && (ksize == 1 || ksize == 3) &&
                   (stride == 1 || stride == 2),
               "mmbs_im2col_t: bad argument");
  const int pad = int(ksize / 2);
  const int64_t oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
  const int64_t p_total = batch * oh * ow;
  MMBS_REQUIRE(p_padded >= p_total && p_padded % 8 == 0 && c % 8 == 0 && al16(x) && al16(out),
               "mmbs_im2col_t: need p_padded >= B*oh*ow, p_padded %% 8 == 0, c %% 8 == 0, 16-byte aligned pointers");
  dim3 grid(unsigned(ceil_div(p_padded, 64)), unsigned(ceil_div(c, 64)), unsigned(ksize * ksize));
  im2col_t_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), int(batch), int(h), int(w), int(c),
      int(ksize), int(stride), pad, int(oh), int(ow), p_total, p_padded, int(oihw_rows));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_pack_conv_weight_dgrad(const float* w, void* out, int64_t c_out, int64_t c_in, int64_t k,
                                           void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(w && out && c_out > 0 && c_in > 0 && k > 0, "mmbs_pack_conv_weight_dgrad: bad argument");
  if (k == 1)   // the transposed matrix: 32x32 tiled cast-transpose (csrc/elementwise.cu)
    return mmbs_cast_transpose_pad_bf16(w, c_out, c_in, c_in, c_out, out, stream);
  if (k == 3)
    pack_conv_weight_dgrad3_kernel<<<tr_blocks(c_out * c_in, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(out), c_out, c_in);
  else
    pack_conv_weight_dgrad_kernel<<<tr_blocks(c_out * c_in * k * k, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(out), c_out, c_in, k);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_unpack_conv_wgrad(const float* g, float* out, int64_t c_out, int64_t c_in, int64_t k,
                                      void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && out && c_out > 0 && c_in > 0 && k > 0, "mmbs_unpack_conv_wgrad: bad argument");
  unpack_conv_wgrad_kernel<<<tr_blocks(c_out * c_in * k * k, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, out, c_out, c_in, k);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_scatter_stride2(const void* g, void* u, int64_t batch, int64_t h, int64_t w, int64_t c,
                                    void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && u && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && al16(g) && al16(u),
               "mmbs_scatter_stride2: bad argument");
  const int64_t total = batch * h * w * (c / 8);
  scatter_stride2_kernel<<<tr_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g), static_cast<uint4*>(u), total, int(h), int(w), int(c / 8));
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_add_relu_mask(const void* a, const void* g, const void* mask, void* out, int64_t elems,
                                  void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(g && mask && out && elems > 0 && elems % 8 == 0 && al16(a) && al16(g) && al16(mask) && al16(out),
               "mmbs_add_relu_mask: bad argument");
  add_relu_mask_kernel<<<tr_blocks(elems / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(g), static_cast<const uint4*>(mask),
      static_cast<uint4*>(out), elems / 8);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

// ------------------------------------------------------------------ fused finalize + apply (training forward)
static int make_bn_train(const mmbs_bn_train_desc* d, BnTrain* b, const char* what) {
  MMBS_REQUIRE(d && d->stats && d->gamma && d->beta && d->scale_out && d->shift_out && d->mean_out && d->invstd_out &&
                   d->c > 0 && d->c % 8 == 0 && d->count > 0 && al16(d->stats) && al16(d->gamma) && al16(d->beta),
               "%s: bad BatchNorm descriptor", what);
  b->stats = d->stats; b->gamma = d->gamma; b->beta = d->beta;
  b->running_mean = d->running_mean; b->running_var = d->running_var;
  b->scale_out = d->scale_out; b->shift_out = d->shift_out; b->mean_out = d->mean_out; b->invstd_out = d->invstd_out;
  b->inv_count = float(1.0 / double(d->count));
  b->unbias = d->count > 1 ? float(double(d->count) / double(d->count - 1)) : 1.0f;
  b->eps = d->eps; b->momentum = d->momentum; b->c = int(d->c);
  return MMBS_OK;
}

template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // Programmatic dependent launch is opt-in here (MMBS_TRAIN_PDL=1): measured on B200 it makes the step SLOWER
  // (forward 3.79 -> 4.08 ms at B = 128) - blocks of these wide grids become resident behind the running
  // persistent convolution and take issue slots / registers from it.
  static const bool use_pdl = []() {
    const char* e = getenv("MMBS_TRAIN_PDL");
    return e && e[0] == '1';
  }();
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  MMBS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, args...));
  count_launch();
  return MMBS_OK;
}

extern "C" int mmbs_bn_train_apply(const mmbs_bn_train_desc* bn, const void* x, const void* residual,
                                   const mmbs_bn_train_desc* res_bn, int32_t relu, void* out, int64_t rows,
                                   void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  BnTrain b, rb;
  if (int rc = make_bn_train(bn, &b, "mmbs_bn_train_apply")) return rc;
  rb = b;
  if (res_bn) {
    if (int rc = make_bn_train(res_bn, &rb, "mmbs_bn_train_apply (residual)")) return rc;
    MMBS_REQUIRE(res_bn->c == bn->c && residual, "mmbs_bn_train_apply: residual BatchNorm width mismatch");
  }
  MMBS_REQUIRE(x && out && rows > 0 && al16(x) && al16(out) && al16(residual), "mmbs_bn_train_apply: bad argument");
  const int c8 = int(bn->c / 8);
  MMBS_REQUIRE(c8 <= 256 && (c8 & (c8 - 1)) == 0, "mmbs_bn_train_apply: c/8 must be a power of two <= 256 (c=%lld)",
               (long long)bn->c);
  const int64_t total = rows * c8;
  const int res_mode = residual ? (res_bn ? 2 : 1) : 0;
  return launch_pdl(bn_train_apply_kernel, dim3(tr_blocks(total, 256 * BNA_ITER)), dim3(256), static_cast<cudaStream_t>(stream), b,
                    rb, static_cast<const uint4*>(x), static_cast<const uint4*>(residual), res_mode, int(relu),
                    static_cast<uint4*>(out), total, c8);
}

extern "C" int mmbs_bn_train_relu_maxpool_3x3s2(const mmbs_bn_train_desc* bn, const void* in, void* out, int64_t batch,
                                                int64_t h, int64_t w, void* stream) {
  if (int rc = mmbs_device_check()) return rc;
  BnTrain b;
  if (int rc = make_bn_train(bn, &b, "mmbs_bn_train_relu_maxpool_3x3s2")) return rc;
  MMBS_REQUIRE(in && out && batch > 0 && h > 0 && w > 0, "mmbs_bn_train_relu_maxpool_3x3s2: bad argument");
  const int64_t oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
  const int c8 = int(bn->c / 8);
  int shift_bits = 0;
  while ((1 << shift_bits) < c8) ++shift_bits;
  MMBS_REQUIRE((1 << shift_bits) == c8 && batch <= 65535 && oh <= 65535,
               "mmbs_bn_train_relu_maxpool_3x3s2: c/8 must be a power of two");
  dim3 grid(unsigned(ceil_div(ow * c8, 256)), unsigned(ceil_div(oh, TMP_ROWS)), unsigned(batch));
  return launch_pdl(bn_train_relu_maxpool_kernel, grid, dim3(256), static_cast<cudaStream_t>(stream), b,
                    static_cast<const uint4*>(in), static_cast<uint4*>(out), int(h), int(w), c8, shift_bits);
}
