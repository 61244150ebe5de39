// Concordance index (Harrell's C) pair counting on the GPU (sm_100a).
//
// Replaces the third-party call at the tail of get_survival_CI:
//   CI = concordance_index(survival_months, -score, vital_status)
//   /root/reference/1_HistoPathology/3_HistoPath_savescore.py:147 (+ 7 copies).
// lifelines is neither vendored nor pinned by the reference and is absent from this image, so the pair rule
// is a restatement of lifelines.utils.concordance._concordance_summary_statistics (PARITY UNPINNED):
//   every subject i is compared with every OBSERVED death j that exited strictly earlier (t_j < t_i), and a
//   censored subject also with the deaths at its own exit time (t_j == t_i); two deaths at the same time are
//   not comparable.  correct += pred_j < pred_i,  tied += pred_j == pred_i,  pairs += 1.
//   C = (correct + tied / 2) / pairs.
// Exact integer counts (u64), O(n^2) tiled through shared memory: n is the number of patients (hundreds to a
// few 10^5), where this is microseconds to milliseconds; all inputs are float64 like the pandas columns.
#include <algorithm>

#include "common.cuh"

namespace mmbs {

constexpr int CI_THREADS = 256;
constexpr int CI_JTILE = 1024;

__global__ void __launch_bounds__(CI_THREADS) cindex_count_kernel(const double* __restrict__ t,
                                                                  const double* __restrict__ p,
                                                                  const uint8_t* __restrict__ e, int64_t n,
                                                                  unsigned long long* __restrict__ out) {
  __shared__ double s_t[CI_JTILE];
  __shared__ double s_p[CI_JTILE];
  __shared__ unsigned long long s_red[3][CI_THREADS / 32];
  const int64_t i = int64_t(blockIdx.x) * CI_THREADS + threadIdx.x;
  const bool valid = i < n;
  const double ti = valid ? t[i] : 0.0, pi = valid ? p[i] : 0.0;
  const bool censored = valid ? (e[i] == 0) : false;
  unsigned long long pairs = 0, correct = 0, tied = 0;
  const int64_t j0 = int64_t(blockIdx.y) * CI_JTILE;
  const int cnt = int(min((long long)CI_JTILE, (long long)(n - j0)));
  // stage the deaths of this j tile (censored subjects never sit on the "earlier" side of a pair)
  for (int k = threadIdx.x; k < CI_JTILE; k += CI_THREADS) {
    const bool death = (k < cnt) && (e[j0 + k] != 0);
    s_t[k] = death ? t[j0 + k] : INFINITY;   // +inf is never earlier than anything
    s_p[k] = death ? p[j0 + k] : 0.0;
  }
  __syncthreads();
  if (valid) {
    uint32_t c_pairs = 0, c_correct = 0, c_tied = 0;
#pragma unroll 4
    for (int k = 0; k < CI_JTILE; ++k) {
      const double tj = s_t[k], pj = s_p[k];
      const bool adm = (tj < ti) || (censored && tj == ti);
      c_pairs += adm ? 1u : 0u;
      c_correct += (adm && pj < pi) ? 1u : 0u;
      c_tied += (adm && pj == pi) ? 1u : 0u;
    }
    pairs = c_pairs; correct = c_correct; tied = c_tied;
  }
  // block reduction -> one atomic triple per block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    correct += __shfl_xor_sync(0xffffffffu, correct, o);
    tied += __shfl_xor_sync(0xffffffffu, tied, o);
  }
  if (lane == 0) { s_red[0][warp] = pairs; s_red[1][warp] = correct; s_red[2][warp] = tied; }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned long long acc = 0;
    for (int w = 0; w < CI_THREADS / 32; ++w) acc += s_red[threadIdx.x][w];
    if (acc) atomicAdd(out + threadIdx.x, acc);
  }
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_concordance_counts(const double* event_times, const double* predicted, const uint8_t* event_observed,
                                       int64_t n, unsigned long long* counts_out, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(event_times && predicted && event_observed && counts_out && n >= 1 && n <= (int64_t(1) << 22),
               "mmbs_concordance_counts: bad argument (n=%lld, supported 1 .. 2^22)", (long long)n);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMBS_CUDA_TRY(cudaMemsetAsync(counts_out, 0, 3 * sizeof(unsigned long long), stream));
  dim3 grid(unsigned(ceil_div(n, CI_THREADS)), unsigned(ceil_div(n, CI_JTILE)));
  MMBS_REQUIRE(grid.y <= 65535, "mmbs_concordance_counts: n too large");
  cindex_count_kernel<<<grid, CI_THREADS, 0, stream>>>(event_times, predicted, event_observed, n, counts_out);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}
