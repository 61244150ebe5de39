// Fused multi-tensor Adam step (L2 weight-decay form) for sm_100a.
//
// Replaces the arithmetic of torch.optim.Adam.step() as the reference's training scripts use it
// (/root/reference/2_GeneExpression/1_GeneExpress_train.py:303-305, two parameter groups lr_rna / lr_mlp;
//  /root/reference/1_HistoPathology/2_HistoPath_train.py:558; /root/reference/5_JointFusion/1_JointFusion_train.py:413-416,
//  three groups): weight_decay is ADDED TO THE GRADIENT (Adam, not AdamW), no amsgrad.
//
// torch's foreach implementation makes ~9 elementwise passes over (p, g, m, v); this is ONE pass: 16 B read +
// 12 B written per parameter (28 B/param: 1.70 GB for the 60.7 M parameters of the RNA model = 0.26 ms at the
// measured HBM peak) - the step-time floor of the RNA / joint models.
// All tensors of up to MMBS_ADAM_MAX_TENSORS parameters go out in one launch: the descriptor table travels
// as a kernel parameter (no device-side table, no host->device copy, graph-capturable).
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace mmbs {

constexpr int AD_THREADS = 256;
constexpr int AD_ITEMS = 16;                       // 4 x float4 per thread
constexpr int AD_CHUNK = AD_THREADS * AD_ITEMS;    // elements per block

struct AdamLaunch {
  mmbs_adam_tensor t[MMBS_ADAM_MAX_TENSORS];
  mmbs_adam_group g[MMBS_ADAM_MAX_GROUPS];
  int32_t block_end[MMBS_ADAM_MAX_TENSORS];       // exclusive prefix end of every tensor's block range
  int32_t n_tensors;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const mmbs_adam_group& h) {
  // same operation order as torch._multi_tensor_adam (add, lerp, mul + addcmul, sqrt / bc2_sqrt + eps, addcdiv)
  g = fmaf(h.weight_decay, p, g);
  m = fmaf(h.one_minus_beta1, g - m, m);
  v = fmaf(h.one_minus_beta2 * g, g, v * h.beta2);
  const float denom = sqrtf(v) / h.bias_correction2_sqrt + h.eps;
  p = p - h.step_size * (m / denom);
}

__global__ void __launch_bounds__(AD_THREADS) adam_step_kernel(const __grid_constant__ AdamLaunch L) {
  // which tensor does this block belong to? (<= 64 entries: binary search on the prefix table in param space)
  int lo = 0, hi = L.n_tensors - 1;
  const int b = int(blockIdx.x);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (b < L.block_end[mid]) hi = mid; else lo = mid + 1;
  }
  const mmbs_adam_tensor& T = L.t[lo];
  const mmbs_adam_group h = L.g[T.group];
  const int64_t first = int64_t(b - (lo ? L.block_end[lo - 1] : 0)) * AD_CHUNK;
  float* __restrict__ p = static_cast<float*>(T.p);
  const float* __restrict__ g = static_cast<const float*>(T.g);
  float* __restrict__ m = static_cast<float*>(T.m);
  float* __restrict__ v = static_cast<float*>(T.v);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec && first + AD_CHUNK <= T.n) {
    float4 pv[4], gv[4], mv[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // all loads first: 16 x 16 B in flight per thread
      const int64_t i = first + (int64_t(j) * AD_THREADS + threadIdx.x) * 4;
      pv[j] = *reinterpret_cast<const float4*>(p + i);
      gv[j] = __ldcs(reinterpret_cast<const float4*>(g + i));   // the gradient is dead after this step: streaming
      mv[j] = *reinterpret_cast<const float4*>(m + i);
      vv[j] = *reinterpret_cast<const float4*>(v + i);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = first + (int64_t(j) * AD_THREADS + threadIdx.x) * 4;
      adam_update(pv[j].x, gv[j].x, mv[j].x, vv[j].x, h);
      adam_update(pv[j].y, gv[j].y, mv[j].y, vv[j].y, h);
      adam_update(pv[j].z, gv[j].z, mv[j].z, vv[j].z, h);
      adam_update(pv[j].w, gv[j].w, mv[j].w, vv[j].w, h);
      *reinterpret_cast<float4*>(p + i) = pv[j];
      *reinterpret_cast<float4*>(m + i) = mv[j];
      *reinterpret_cast<float4*>(v + i) = vv[j];
    }
  } else {
    const int64_t end = min(T.n, first + AD_CHUNK);
    for (int64_t i = first + threadIdx.x; i < end; i += AD_THREADS) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(pp, g[i], mm, vv, h);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_adam_step(const mmbs_adam_tensor* tensors_host, int32_t n_tensors, const mmbs_adam_group* groups_host,
                              int32_t n_groups, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(tensors_host && groups_host && n_tensors >= 0 && n_groups >= 1 && n_groups <= MMBS_ADAM_MAX_GROUPS,
               "mmbs_adam_step: bad argument (n_tensors=%d n_groups=%d, at most %d groups)", n_tensors, n_groups,
               MMBS_ADAM_MAX_GROUPS);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int32_t i = 0; i < n_tensors; ++i) {
    const mmbs_adam_tensor& t = tensors_host[i];
    MMBS_REQUIRE(t.p && t.g && t.m && t.v && t.n >= 0 && t.group >= 0 && t.group < n_groups,
                 "mmbs_adam_step: tensor %d: null pointer, negative size or group %d out of range", i, t.group);
    MMBS_REQUIRE((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                  reinterpret_cast<uintptr_t>(t.v)) % 4 == 0, "mmbs_adam_step: tensor %d: pointers must be 4-byte aligned", i);
  }
  int32_t done = 0;
  while (done < n_tensors) {
    AdamLaunch L;
    std::memset(&L, 0, sizeof(L));
    std::memcpy(L.g, groups_host, sizeof(mmbs_adam_group) * size_t(n_groups));
    int64_t blocks = 0;
    int32_t k = 0;
    while (done < n_tensors && k < MMBS_ADAM_MAX_TENSORS) {
      const mmbs_adam_tensor& t = tensors_host[done];
      const int64_t nb = ceil_div(t.n, AD_CHUNK);
      if (blocks + nb > 0x7fffffff) break;
      ++done;
      if (nb == 0) continue;
      L.t[k] = t;
      blocks += nb;
      L.block_end[k] = int32_t(blocks);
      ++k;
    }
    L.n_tensors = k;
    if (k == 0) {
      MMBS_REQUIRE(done >= n_tensors || tensors_host[done].n < (int64_t(0x7fffffff) * AD_CHUNK), "mmbs_adam_step: tensor too large");
      continue;
    }
    adam_step_kernel<<<unsigned(blocks), AD_THREADS, 0, stream>>>(L);
    MMBS_LAUNCH_CHECK();
  }
  return MMBS_OK;
}
