// Training-time patch augmentation on the device (sm_100a): the uint8 arithmetic of
//   transforms.RandomHorizontalFlip -> RandomVerticalFlip -> ColorJitter(64/255, 0.75, 0.25, 0.04)
// of /root/reference/1_HistoPathology/2_HistoPath_train.py:474-488, applied to decoded patches (uint8 HWC, what
// PatchBagDataset.__getitem__ holds after Image.open(...).convert('RGB'), 1_HistoPathology/models.py:280-286) and written
// as uint8 NCHW - the layout ResNet.forward_extract takes for raw pixels (ToTensor + Normalize are fused into its stem
// pack kernel).  Bit-exact with torchvision 0.26 / Pillow 12.2 on PIL images (oracle/augment_oracle.py, pinned
// exhaustively): Image.blend in single precision without contraction, rgb2l in 16-bit fixed point, the HSV round trip of
// libImaging/Convert.c with its float / double mix, adjust_hue's uint8 wrap-around.
//
// ColorJitter's contrast blends with the mean luma of the image AS IT IS at that point of the (random) operation order:
// pass 1 runs the operations in front of the contrast step and sums L per image (integer atomics: exact, order-free),
// pass 2 runs the whole chain.  One thread per pixel.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace mmbs {

struct AugParams {          // one per image (mmbs.h: mmbs_aug_params)
  int32_t hflip, vflip;
  int32_t order[4];         // operations in application order: 0 brightness, 1 contrast, 2 saturation, 3 hue, -1 none
  float factor[3];          // brightness, contrast, saturation factors (Image.blend takes a C float)
  int32_t hue_shift;        // torchvision adjust_hue: np.int32(hue_factor * 255).astype(np.uint8), computed by the host
};

__device__ __forceinline__ uint32_t luma(uint32_t r, uint32_t g, uint32_t b) {   // Convert.c rgb2l
  return (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16;
}
// Blend.c: in1 + alpha * (in2 - in1) in float, multiply and add rounded separately, truncated (clipped outside [0, 1])
__device__ __forceinline__ uint32_t blend1(int in1, int in2, float alpha, bool inside) {
  const float t = __fadd_rn(float(in1), __fmul_rn(alpha, float(in2 - in1)));
  if (inside) return uint32_t(int(t));
  return t <= 0.0f ? 0u : (t >= 255.0f ? 255u : uint32_t(int(t)));
}
__device__ __forceinline__ void blend3(int d0, int d1, int d2, float alpha, uint32_t& r, uint32_t& g, uint32_t& b) {
  const bool inside = alpha >= 0.0f && alpha <= 1.0f;
  r = blend1(d0, int(r), alpha, inside);
  g = blend1(d1, int(g), alpha, inside);
  b = blend1(d2, int(b), alpha, inside);
}
__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Convert.c rgb2hsv_row -> hue + shift (uint8 wrap) -> hsv2rgb
__device__ __forceinline__ void hue_shift(uint32_t shift, uint32_t& r, uint32_t& g, uint32_t& b) {
  const int ri = int(r), gi = int(g), bi = int(b);
  const int maxc = max(ri, max(gi, bi)), minc = min(ri, min(gi, bi));
  uint32_t uh = 0, us = 0;
  const uint32_t uv = uint32_t(maxc);
  if (minc != maxc) {
    const float cr = float(maxc - minc);
    const float s = __fdiv_rn(cr, float(maxc));
    const float rc = __fdiv_rn(float(maxc - ri), cr), gc = __fdiv_rn(float(maxc - gi), cr), bc = __fdiv_rn(float(maxc - bi), cr);
    float h;   // every assignment rounds a double expression to float, like the C source
    if (ri == maxc) h = __fsub_rn(bc, gc);
    else if (gi == maxc) h = float(__dsub_rn(__dadd_rn(2.0, double(rc)), double(bc)));
    else h = float(__dsub_rn(__dadd_rn(4.0, double(gc)), double(rc)));
    const double t = __dadd_rn(__ddiv_rn(double(h), 6.0), 1.0);   // in [5/6, 11/6): fmod(t, 1) = t - floor(t), exact
    h = float(__dsub_rn(t, floor(t)));
    uh = uint32_t(clip8(int(__dmul_rn(double(h), 255.0))));
    us = uint32_t(clip8(int(__dmul_rn(double(s), 255.0))));
  }
  uh = (uh + shift) & 255u;
  if (us == 0u) {
    r = g = b = uv;
    return;
  }
  const double hf = __ddiv_rn(__dmul_rn(double(float(uh)), 6.0), 255.0);
  const int i = int(floor(hf));
  const float f = float(__dsub_rn(hf, double(float(i))));
  const float fs = float(__ddiv_rn(double(float(us)), 255.0));
  const double vf = double(float(uv)), fsd = double(fs), fd = double(f);
  const uint32_t p = uint32_t(clip8(int(round(__dmul_rn(vf, __dsub_rn(1.0, fsd))))));
  const uint32_t q = uint32_t(clip8(int(round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fsd, fd)))))));
  const uint32_t t2 = uint32_t(clip8(int(round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fsd, __dsub_rn(1.0, fd))))))));
  switch (i % 6) {
    case 0: r = uv; g = t2; b = p; break;
    case 1: r = q; g = uv; b = p; break;
    case 2: r = p; g = uv; b = t2; break;
    case 3: r = p; g = q; b = uv; break;
    case 4: r = t2; g = p; b = uv; break;
    default: r = uv; g = p; b = q; break;
  }
}

// operations order[first .. last) on one pixel; `mean` = the grey level of the contrast step
__device__ __forceinline__ void run_ops(const AugParams& P, int first, int last, int mean, uint32_t& r, uint32_t& g,
                                        uint32_t& b) {
  for (int k = first; k < last; ++k) {
    const int op = P.order[k];
    if (op == 0) {
      blend3(0, 0, 0, P.factor[0], r, g, b);                           // ImageEnhance.Brightness: blend(black, image)
    } else if (op == 1) {
      blend3(mean, mean, mean, P.factor[1], r, g, b);                  // ImageEnhance.Contrast: blend(mean grey, image)
    } else if (op == 2) {
      const int l = int(luma(r, g, b));
      blend3(l, l, l, P.factor[2], r, g, b);                           // ImageEnhance.Color: blend(L image, image)
    } else if (op == 3) {
      hue_shift(uint32_t(P.hue_shift) & 255u, r, g, b);
    }
  }
}
__device__ __forceinline__ int contrast_position(const AugParams& P) {
  for (int k = 0; k < 4; ++k)
    if (P.order[k] == 1) return k;
  return -1;
}

constexpr int AUG_THREADS = 256;

template <bool SUM_ONLY>
__global__ void __launch_bounds__(AUG_THREADS) augment_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h,
                                                             int w, const AugParams* __restrict__ params,
                                                             uint32_t* __restrict__ lsum) {
  __shared__ uint32_t s_red[AUG_THREADS / 32];
  const int64_t n = blockIdx.y;
  const AugParams P = params[n];
  const int cpos = contrast_position(P);
  if (SUM_ONLY && cpos < 0) return;
  const int pix = blockIdx.x * AUG_THREADS + threadIdx.x;
  const bool ok = pix < h * w;
  uint32_t r = 0, g = 0, b = 0;
  const int y = ok ? pix / w : 0, x = ok ? pix - y * w : 0;
  if (ok) {
    const int sy = P.vflip ? h - 1 - y : y, sx = P.hflip ? w - 1 - x : x;   // flips first (RandomHorizontal/VerticalFlip)
    const uint8_t* src = in + ((n * h + sy) * int64_t(w) + sx) * 3;
    r = src[0]; g = src[1]; b = src[2];
  }
  if (SUM_ONLY) {
    uint32_t l = 0;
    if (ok) {
      run_ops(P, 0, cpos, 0, r, g, b);
      l = luma(r, g, b);
    }
    l = __reduce_add_sync(0xffffffffu, l);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int i = 0; i < AUG_THREADS / 32; ++i) t += s_red[i];
      atomicAdd(lsum + n, t);   // <= 255 * h * w per image: fits (the host checks h * w <= 2^24)
    }
    return;
  }
  if (!ok) return;
  // ImageStat mean of the L image, int(mean + 0.5) (ImageEnhance.Contrast), in double like the Python expression
  const int mean = cpos >= 0 ? int(__dadd_rn(__ddiv_rn(double(lsum[n]), double(h * w)), 0.5)) : 0;
  run_ops(P, 0, 4, mean, r, g, b);
  uint8_t* dst = out + (n * 3 * h + y) * int64_t(w) + x;   // NCHW
  dst[0] = uint8_t(r);
  dst[int64_t(h) * w] = uint8_t(g);
  dst[2 * int64_t(h) * w] = uint8_t(b);
}

// ---- transforms.Resize on PIL images = Image.resize(BILINEAR): Pillow's antialiased triangle filter (Resample.c), a
// horizontal pass then a vertical pass over 22-bit fixed-point coefficient rows, each pass rounded to uint8.
constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// one thread per output pixel of the pass; AXIS 0: along rows (vertical pass), 1: along columns (horizontal pass)
template <int AXIS>
__global__ void __launch_bounds__(256) resample_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t batch,
                                                       int in_h, int in_w, int out_h, int out_w,
                                                       const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                       int ksize) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= batch * out_h * out_w) return;
  const int ox = int(i % out_w), oy = int((i / out_w) % out_h);
  const int64_t n = i / (int64_t(out_w) * out_h);
  const int o = AXIS ? ox : oy;
  const int first = __ldg(bounds + 2 * o), cnt = __ldg(bounds + 2 * o + 1);
  const int32_t* k = kk + int64_t(o) * ksize;
  int32_t a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
  const int64_t step = AXIS ? 3 : int64_t(in_w) * 3;
  const uint8_t* src = in + ((n * in_h + (AXIS ? oy : first)) * int64_t(in_w) + (AXIS ? first : ox)) * 3;
  for (int x = 0; x < cnt; ++x, src += step) {
    const int32_t c = __ldg(k + x);
    a0 += int32_t(src[0]) * c;
    a1 += int32_t(src[1]) * c;
    a2 += int32_t(src[2]) * c;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = uint8_t(clip8(a0 >> RS_PRECISION_BITS));
  dst[1] = uint8_t(clip8(a1 >> RS_PRECISION_BITS));
  dst[2] = uint8_t(clip8(a2 >> RS_PRECISION_BITS));
}

}  // namespace mmbs

using namespace mmbs;

// Pillow precompute_coeffs + normalize_coeffs_8bpc (bilinear filter, the whole axis): host code, double arithmetic in the
// order of the C source.  bounds[2 * out_size] = (first input index, count); kk[out_size * ksize]; returns ksize.
extern "C" int mmbs_resample_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, int kk_capacity) {
  if (in_size < 1 || out_size < 1) return -1;
  const double scale = double(in_size) / double(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = int(ceil(support)) * 2 + 1;
  if (bounds == nullptr || kk == nullptr) return ksize;   // size query
  if (int64_t(out_size) * ksize > int64_t(kk_capacity)) return -1;
  const double ss = 1.0 / filterscale;
  std::vector<double> w(size_t(ksize), 0.0);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      w[size_t(x)] = t < 1.0 ? 1.0 - t : 0.0;
      ww += w[size_t(x)];
    }
    for (int x = 0; x < ksize; ++x) {
      double v = 0.0;
      if (x < xmax) v = (ww != 0.0) ? w[size_t(x)] / ww : w[size_t(x)];
      kk[int64_t(xx) * ksize + x] = v < 0 ? int32_t(-0.5 + v * (1 << RS_PRECISION_BITS)) : int32_t(0.5 + v * (1 << RS_PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

extern "C" int mmbs_resize_bilinear_u8(const uint8_t* in_hwc, uint8_t* out_hwc, uint8_t* tmp, int64_t batch, int in_h, int in_w,
                                       int out_h, int out_w, const int32_t* bounds_w, const int32_t* kk_w, int ksize_w,
                                       const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in_hwc && out_hwc && tmp && bounds_w && kk_w && bounds_h && kk_h, "mmbs_resize_bilinear_u8: null pointer");
  MMBS_REQUIRE(batch >= 1 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1 && ksize_w >= 1 && ksize_h >= 1,
               "mmbs_resize_bilinear_u8: bad shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // horizontal pass: [B, in_h, in_w] -> tmp [B, in_h, out_w]; vertical pass: tmp -> [B, out_h, out_w]
  resample_kernel<1><<<unsigned(ceil_div(batch * in_h * out_w, 256)), 256, 0, stream>>>(in_hwc, tmp, batch, in_h, in_w, in_h,
                                                                                        out_w, bounds_w, kk_w, ksize_w);
  MMBS_LAUNCH_CHECK();
  resample_kernel<0><<<unsigned(ceil_div(batch * out_h * out_w, 256)), 256, 0, stream>>>(tmp, out_hwc, batch, in_h, out_w, out_h,
                                                                                         out_w, bounds_h, kk_h, ksize_h);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_augment_u8(const uint8_t* in_hwc, uint8_t* out_chw, int64_t batch, int h, int w, const void* params_dev,
                               uint32_t* lsum_ws, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(in_hwc && out_chw && params_dev && lsum_ws, "mmbs_augment_u8: null pointer");
  MMBS_REQUIRE(batch >= 1 && batch <= 65535 && h >= 1 && w >= 1 && int64_t(h) * w <= (int64_t(1) << 24),
               "mmbs_augment_u8: batch=%lld h=%d w=%d out of range", (long long)batch, h, w);
  static_assert(sizeof(AugParams) == 40, "mmbs_aug_params layout");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const dim3 grid(unsigned(ceil_div(int64_t(h) * w, AUG_THREADS)), unsigned(batch));
  MMBS_CUDA_TRY(cudaMemsetAsync(lsum_ws, 0, sizeof(uint32_t) * size_t(batch), stream));
  augment_kernel<true><<<grid, AUG_THREADS, 0, stream>>>(in_hwc, out_chw, h, w, static_cast<const AugParams*>(params_dev),
                                                         lsum_ws);
  MMBS_LAUNCH_CHECK();
  augment_kernel<false><<<grid, AUG_THREADS, 0, stream>>>(in_hwc, out_chw, h, w, static_cast<const AugParams*>(params_dev),
                                                          lsum_ws);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
