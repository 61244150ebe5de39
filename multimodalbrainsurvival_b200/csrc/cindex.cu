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

// ---- large cohorts: the same counts as a 2-D dominance problem, O(n * S) instead of O(n^2).
// With the deaths sorted by exit time (position k) and, separately, by prediction (perm[s] = time position of the s-th
// smallest prediction, inv = its inverse), a subject with L admissible deaths (a prefix of the time order) and lo / hi =
// number of deaths with a smaller / smaller-or-equal prediction has
//     correct = F(lo, L),   tied = F(hi, L) - F(lo, L),   F(m, L) = #{ s < m : perm[s] < L }.
// F is evaluated from a table of block corners T[bs][bv] = #{ s < bs S : perm[s] < bv S } (built by the caller: a 2-D
// histogram and two prefix sums) plus two partial scans of at most S entries: the tail of the row block in perm and the
// tail of the column block in inv.  One warp per subject, coalesced scans.
constexpr int CI_DOM_THREADS = 256;

__device__ __forceinline__ unsigned long long dominance_count(const int32_t* __restrict__ perm, const int32_t* __restrict__ inv,
                                                              const int64_t* __restrict__ table, int row_len, int shift,
                                                              int64_t m, int64_t l, int lane) {
  const int64_t bs = m >> shift, bv = l >> shift;
  const int64_t s0 = bs << shift, v0 = bv << shift;
  uint32_t cnt = 0;
  for (int64_t s = s0 + lane; s < m; s += 32) cnt += (int64_t(__ldg(perm + s)) < l) ? 1u : 0u;
  for (int64_t v = v0 + lane; v < l; v += 32) cnt += (int64_t(__ldg(inv + v)) < s0) ? 1u : 0u;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  return static_cast<unsigned long long>(__ldg(table + bs * row_len + bv)) + cnt;
}

__global__ void __launch_bounds__(CI_DOM_THREADS) cindex_dominance_kernel(
    const int32_t* __restrict__ perm, const int32_t* __restrict__ inv, const int64_t* __restrict__ table, int row_len, int shift,
    const int64_t* __restrict__ lo, const int64_t* __restrict__ hi, const int64_t* __restrict__ adm, int64_t n,
    unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (CI_DOM_THREADS / 32);
  unsigned long long correct = 0, tied = 0;
  for (int64_t q = int64_t(blockIdx.x) * (CI_DOM_THREADS / 32) + (threadIdx.x >> 5); q < n; q += warps) {
    const int64_t l = __ldg(adm + q), a = __ldg(lo + q), b = __ldg(hi + q);
    if (l == 0) continue;   // (warp-uniform)
    const unsigned long long fa = dominance_count(perm, inv, table, row_len, shift, a, l, lane);
    const unsigned long long fb = (b == a) ? fa : dominance_count(perm, inv, table, row_len, shift, b, l, lane);
    correct += fa;
    tied += fb - fa;
  }
  if (lane == 0) {
    if (correct) atomicAdd(out + 1, correct);
    if (tied) atomicAdd(out + 2, tied);
  }
}

}  // namespace mmbs

using namespace mmbs;

extern "C" int mmbs_concordance_dominance(const int32_t* perm, const int32_t* inv, const int64_t* table, int64_t n_deaths,
                                          int block_shift, const int64_t* lo, const int64_t* hi, const int64_t* admissible,
                                          int64_t n, unsigned long long* counts_out, void* stream_) {
  if (int rc = mmbs_device_check()) return rc;
  MMBS_REQUIRE(perm && inv && table && lo && hi && admissible && counts_out, "mmbs_concordance_dominance: null pointer");
  MMBS_REQUIRE(n >= 1 && n_deaths >= 1 && n_deaths < (int64_t(1) << 31) && block_shift >= 4 && block_shift <= 20,
               "mmbs_concordance_dominance: bad argument");
  const int row_len = int(((n_deaths + (int64_t(1) << block_shift) - 1) >> block_shift) + 1);
  const int64_t warps = ceil_div(n, 1);
  const unsigned grid = unsigned(std::min<int64_t>(ceil_div(warps, CI_DOM_THREADS / 32), int64_t(sm_count()) * 16));
  cindex_dominance_kernel<<<grid, CI_DOM_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
      perm, inv, table, row_len, block_shift, lo, hi, admissible, n, counts_out);
  MMBS_LAUNCH_CHECK();
  return MMBS_OK;
}

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
