// Discrete-time survival negative log-likelihood (the `survival_bin` task), forward + gradient in one pass (sm_100a).
//
// Replaces nll_loss() / NLLSurvLoss of the reference
//   /root/reference/1_HistoPathology/models.py:120-153 (NLLSurvLoss), :155-232 (nll_loss);  SURVEY.md 8(f) row 4:
//   hazards = sigmoid(h); S = cumprod(1 - hazards); S_padded = [1, S];
//   uncensored = -(1 - c) (log clamp(S_padded[y], eps) + log clamp(hazards[y], eps));
//   censored   = -c log clamp(S_padded[y + 1], eps);   loss = (1 - alpha) censored + uncensored;  mean | sum over n.
// One thread per sample walks its n_bins logits once (the reference launches ~15 elementwise / gather kernels and
// autograd replays as many): it writes the sample's loss and d loss_i / d h_i[:], the backward pass only scales.
// clamp(min = eps) passes the gradient where the clamped value is >= eps (torch's clamp backward).
#include <algorithm>

#include "common.cuh"

namespace mmbs {

constexpr int NLL_THREADS = 256;

__global__ void __launch_bounds__(NLL_THREADS) nll_surv_kernel(const float* __restrict__ h, const int64_t* __restrict__ y,
                                                               const float* __restrict__ c, int64_t n, int k, float alpha,
                                                               float eps, float* __restrict__ loss_i,
                                                               float* __restrict__ grad_unit, int32_t* __restrict__ bad) {
  const int64_t i = int64_t(blockIdx.x) * NLL_THREADS + threadIdx.x;
  if (i >= n) return;
  const int64_t yi = y[i];
  const float ci = c[i];
  if (yi < 0 || yi >= k) {   // torch.gather would raise: poison the sample and flag it
    loss_i[i] = __int_as_float(0x7fc00000);
    for (int j = 0; j < k; ++j) grad_unit[i * k + j] = __int_as_float(0x7fc00000);
    atomicOr(bad, 1);
    return;
  }
  const float* hi = h + i * k;
  float* gi = grad_unit + i * k;
  // S_padded[y] = prod_{j < y} (1 - hz_j),  hazards[y],  S_padded[y + 1] = S_padded[y] (1 - hz_y)
  float s_prev = 1.f, hz_y = 0.f;
  for (int j = 0; j <= int(yi); ++j) {
    const float hz = 1.f / (1.f + expf(-hi[j]));
    if (j < int(yi)) s_prev *= 1.f - hz;
    else hz_y = hz;
  }
  const float s_this = s_prev * (1.f - hz_y);
  const float unc = -(1.f - ci) * (logf(fmaxf(s_prev, eps)) + logf(fmaxf(hz_y, eps)));
  const float cen = -ci * logf(fmaxf(s_this, eps));
  loss_i[i] = (1.f - alpha) * cen + unc;
  // d log(1 - sigmoid(h)) / dh = -sigmoid(h);  d log sigmoid(h) / dh = 1 - sigmoid(h)
  const float w_prev = (s_prev >= eps) ? (1.f - ci) : 0.f;                 // weight of -log S_padded[y]
  const float w_hz = (hz_y >= eps) ? (1.f - ci) : 0.f;                     // weight of -log hazards[y]
  const float w_this = (s_this >= eps) ? (1.f - alpha) * ci : 0.f;         // weight of -log S_padded[y + 1]
  for (int j = 0; j < k; ++j) {
    float g = 0.f;
    if (j <= int(yi)) {
      const float hz = 1.f / (1.f + expf(-hi[j]));
      if (j < int(yi)) g = (w_prev + w_this) * hz;
      else g = w_this * hz - w_hz * (1.f - hz);
    }
    gi[j] = g;
  }
}

// loss = sum_i loss_i (/ n for the mean), fp64 accumulation in a fixed order
__global__ void __launch_bounds__(1024) nll_reduce_kernel(const float* __restrict__ loss_i, int64_t n, int mean,
                                                          float* __restrict__ loss_out) {
  __shared__ double s_red[32];
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) t += double(loss_i[i]);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += s_red[w];
    loss_out[0] = float(mean ? s / double(n) : s);
  }
}

__global__ void __launch_bounds__(NLL_THREADS) nll_scale_kernel(const float* __restrict__ grad_unit,
                                                                const float* __restrict__ grad_loss, float scale,
                                                                int64_t total, float* __restrict__ grad_h) {
  const float g = grad_loss[0] * scale;
  for (int64_t i = int64_t(blockIdx.x) * NLL_THREADS + threadIdx.x; i < total; i += int64_t(gridDim.x) * NLL_THREADS)
    grad_h[i] = grad_unit[i] * g;
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_nll_surv_forward(const float* h, const int64_t* y, const float* c, int64_t n, int32_t n_bins,
                                     float alpha, float eps, int32_t reduction_mean, float* loss_i, float* grad_unit,
                                     float* loss_out, int32_t* flags_out, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(h && y && c && loss_i && grad_unit && loss_out && flags_out, "mmbs_nll_surv_forward: null pointer");
  MMBS_REQUIRE(n >= 1 && n_bins >= 1 && n_bins <= 4096, "mmbs_nll_surv_forward: n=%lld n_bins=%d", (long long)n, n_bins);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMBS_CUDA_TRY(cudaMemsetAsync(flags_out, 0, sizeof(int32_t), stream));
  nll_surv_kernel<<<unsigned(ceil_div(n, NLL_THREADS)), NLL_THREADS, 0, stream>>>(h, y, c, n, n_bins, alpha, eps, loss_i,
                                                                                  grad_unit, flags_out);
  MMBS_LAUNCH_CHECK();
  nll_reduce_kernel<<<1, 1024, 0, stream>>>(loss_i, n, reduction_mean, loss_out);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

extern "C" int mmbs_nll_surv_backward(const float* grad_unit, const float* grad_loss, int64_t n, int32_t n_bins,
                                      int32_t reduction_mean, float* grad_h, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(grad_unit && grad_loss && grad_h, "mmbs_nll_surv_backward: null pointer");
  MMBS_REQUIRE(n >= 1 && n_bins >= 1, "mmbs_nll_surv_backward: n=%lld n_bins=%d", (long long)n, n_bins);
  const int64_t total = n * n_bins;
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, NLL_THREADS), int64_t(sm_count()) * 8)));
  nll_scale_kernel<<<grid, NLL_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
      grad_unit, grad_loss, reduction_mean ? 1.f / float(n) : 1.f, total, grad_h);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
