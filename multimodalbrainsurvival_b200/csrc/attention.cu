// Bag aggregation of TanhAttention (sm_100a), the tail of
//   /root/reference/1_HistoPathology/models.py:22-33 (TanhAttention.forward) and the projection head of
//   AggregationProjectModel.extract (:59-88, F.tanh after nn.Linear):
//     logits  = tanh(x W^T) . vector          (x W^T comes from the tcgen05 linear plan, fp32 [rows, dim])
//     weights = softmax(logits, dim = bag)
//     out     = x * weights * bag             (per patch, what the aggregator returns)
//     pooled  = sum_p x[b, p, :] weights[p]   (= out.mean(dim = 1), what AggregationModel.extract keeps)
// Two kernels: one warp per (case, patch) row for the logits, one block per case for the softmax and the weighted rows.
#include "common.cuh"

namespace mmbs {

constexpr int AT_THREADS = 256;
constexpr int AT_MAX_BAG = 8192;   // softmax weights of one bag in shared memory

__global__ void __launch_bounds__(AT_THREADS) attn_logits_kernel(const float* __restrict__ h, const float* __restrict__ v,
                                                                int64_t rows, int dim, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (AT_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* hr = h + row * dim;
  float acc = 0.f;
  if ((dim & 3) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0) {
    for (int d = lane * 4; d < dim; d += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(hr + d));
      const float4 w = __ldg(reinterpret_cast<const float4*>(v + d));
      acc += tanhf(a.x) * w.x + tanhf(a.y) * w.y + tanhf(a.z) * w.z + tanhf(a.w) * w.w;
    }
  } else {
    for (int d = lane; d < dim; d += 32) acc += tanhf(__ldg(hr + d)) * __ldg(v + d);
  }
  acc = warp_sum(acc);
  if (lane == 0) logits[row] = acc;
}

__global__ void __launch_bounds__(AT_THREADS) attn_apply_kernel(const float* __restrict__ x, float* attn, int bag, int dim,
                                                               float* __restrict__ out, float* __restrict__ pooled) {
  __shared__ float s_w[AT_MAX_BAG];
  __shared__ float s_red[AT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x;
  float* a = attn + b * bag;
  float m = -INFINITY;
  for (int p = tid; p < bag; p += AT_THREADS) {
    s_w[p] = a[p];
    m = fmaxf(m, s_w[p]);
  }
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
  for (int w = 1; w < AT_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float s = 0.f;
  for (int p = tid; p < bag; p += AT_THREADS) {
    const float e = expf(s_w[p] - m);
    s_w[p] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < AT_THREADS / 32; ++w) s += s_red[w];
  const float inv = 1.0f / s;
  for (int p = tid; p < bag; p += AT_THREADS) {
    const float w = s_w[p] * inv;
    s_w[p] = w;
    a[p] = w;
  }
  __syncthreads();
  const float* xb = x + b * int64_t(bag) * dim;
  const float scale = float(bag);
  for (int d = tid; d < dim; d += AT_THREADS) {
    float acc = 0.f;
    for (int p = 0; p < bag; ++p) {
      const float xv = __ldg(xb + int64_t(p) * dim + d);
      const float w = s_w[p];
      acc = fmaf(xv, w, acc);
      if (out != nullptr) out[(b * bag + p) * dim + d] = xv * w * scale;   // x * attention_weights * x.shape[1] (models.py:32)
    }
    if (pooled != nullptr) pooled[b * dim + d] = acc;
  }
}

__global__ void __launch_bounds__(256) tanh_inplace_kernel(float* __restrict__ x, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) x[i] = tanhf(x[i]);
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_attention_pool(const float* x, const float* h, const float* vector, int64_t batch, int bag, int dim,
                                   float* attn, float* out, float* pooled, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x && h && vector && attn && (out || pooled), "mmbs_attention_pool: null pointer");
  MMBS_REQUIRE(batch >= 1 && bag >= 1 && bag <= AT_MAX_BAG && dim >= 1, "mmbs_attention_pool: batch=%lld bag=%d dim=%d",
               (long long)batch, bag, dim);
  MMBS_REQUIRE(batch * bag < (int64_t(1) << 31), "mmbs_attention_pool: too many rows");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t rows = batch * bag;
  attn_logits_kernel<<<unsigned(ceil_div(rows, AT_THREADS / 32)), AT_THREADS, 0, stream>>>(h, vector, rows, dim, attn);
  MMBS_LAUNCH_CHECK();
  attn_apply_kernel<<<unsigned(batch), AT_THREADS, 0, stream>>>(x, attn, bag, dim, out, pooled);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_tanh_inplace_f32(float* x, int64_t n, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(x && n >= 0, "mmbs_tanh_inplace_f32: bad argument");
  if (n == 0) return MMBS_OK;
  tanh_inplace_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, n);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
