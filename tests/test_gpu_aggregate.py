"""GPU parity: csrc/segmean.cu vs the aggregation oracle and the reference golden vectors.
Grouping (ids, counts, order) bit-exact; means fp32-accumulated: 2e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import aggregate_oracle

pytestmark = pytest.mark.gpu


def test_case_features_match_reference(golden):
    from multimodalbrainsurvival_b200 import aggregate
    g = golden("aggregate_reference.npz")
    cases = [str(c) for c in g["cases"]]
    order = [str(c) for c in g["case_uniques"]]
    uniq, feats = aggregate.aggregate_case_features(g["features"], cases, case_order=order)
    assert uniq == order and feats.dtype == np.float64
    np.testing.assert_allclose(feats, g["features_final"], rtol=2e-6, atol=2e-6)
    uniq2, feats2 = aggregate.aggregate_case_features(g["features"], cases)
    assert sorted(uniq2) == sorted(order)
    np.testing.assert_allclose(feats2[[uniq2.index(c) for c in order]], g["features_final"], rtol=2e-6, atol=2e-6)


def test_survival_grouping_matches_reference(golden):
    from multimodalbrainsurvival_b200 import aggregate
    g = golden("aggregate_reference.npz")
    cases = [str(c) for c in g["cases"]]
    ids, score, sm, vs = aggregate.survival_grouping(g["outputs"], cases, g["survival"], g["vital"])
    assert ids == [str(c) for c in g["ci_ids"]]
    np.testing.assert_allclose(score, g["ci_score"], rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(sm, g["ci_survival"])
    np.testing.assert_array_equal(vs, g["ci_vital"])


@pytest.mark.parametrize("n,d,g", [(1, 1, 1), (1000, 1, 37), (5000, 2048, 50), (4097, 12, 300), (20000, 64, 70000),
                                    (3000, 7, 5)])
def test_segmented_mean_random(n, d, g):
    from multimodalbrainsurvival_b200 import aggregate
    rng = np.random.default_rng(n + d + g)
    v = rng.standard_normal((n, d)).astype(np.float32)
    seg = rng.integers(0, g, n).astype(np.int32)
    mean, cnt, last = aggregate.segmented_mean(torch.tensor(v, device="cuda:0"), torch.tensor(seg, device="cuda:0"), g)
    o_mean, o_cnt = aggregate_oracle.segment_mean(v, seg, g)
    assert np.array_equal(cnt.cpu().numpy(), o_cnt)                       # grouping: bit-exact
    o_last = np.full(g, -1, np.int64)
    o_last[seg] = np.arange(n)
    assert np.array_equal(last.cpu().numpy(), o_last)
    m = mean.cpu().numpy()
    nz = o_cnt > 0
    np.testing.assert_allclose(m[nz], o_mean[nz], rtol=2e-5, atol=2e-6)
    assert np.isnan(m[~nz]).all()


def test_out_of_range_segment_id_poisons_the_result():
    """A segment id outside [0, n_seg) must not shift the other segments silently: NaN means, counts -1."""
    from multimodalbrainsurvival_b200 import aggregate
    v = torch.randn(300, 8, device="cuda")
    seg = (torch.arange(300, device="cuda") % 7).to(torch.int32)
    seg[123] = 7
    mean, counts, last = aggregate.segmented_mean(v, seg, 7)
    assert torch.isnan(mean).all() and (counts == -1).all() and (last == -1).all()
    seg[123] = -2
    mean, counts, last = aggregate.segmented_mean(v, seg, 7)
    assert torch.isnan(mean).all() and (counts == -1).all()
    seg[123] = 3
    mean, counts, _ = aggregate.segmented_mean(v, seg, 7)
    assert torch.isfinite(mean).all() and int(counts.sum()) == 300


def test_hundred_patches_per_case_2048():
    """BASELINE config 2 tail: 100 patches per case, 2048-d features."""
    from multimodalbrainsurvival_b200 import aggregate
    n_case, per = 64, 100
    torch.manual_seed(1)
    feats = torch.randn(n_case * per, 2048, device="cuda:0")
    cases = [f"case{(i * 7919) % n_case:03d}" for i in range(n_case * per)]
    uniq, out = aggregate.aggregate_case_features(feats, cases)
    ref = aggregate_oracle.case_mean_features(feats.cpu().numpy(), cases, uniq)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)


def _ci_case(seed, n, tie_times=False, tie_pred=False, p_event=0.6):
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 200, n)
    if tie_times:
        t = np.round(t / 8) * 8          # many equal exit times (deaths with deaths, deaths with censorings)
    p = rng.normal(size=n)
    if tie_pred:
        p = np.round(p, 1)
    e = rng.random(n) < p_event
    return t, p, e


@pytest.mark.parametrize("n,tt,tp", [(1, False, False), (2, False, False), (37, True, True), (1000, False, False),
                                     (1500, True, False), (1025, False, True), (3000, True, True)])
def test_concordance_counts_bit_exact_vs_oracle(n, tt, tp):
    from multimodalbrainsurvival_b200 import aggregate
    from oracle import cindex_oracle
    t, p, e = _ci_case(100 + n, n, tt, tp)
    got = aggregate.concordance_counts(t, p, e)
    assert got == cindex_oracle.concordance_counts(t, p, e)


def test_concordance_index_edge_cases_and_get_survival_ci():
    from multimodalbrainsurvival_b200 import aggregate
    from oracle import cindex_oracle
    t = np.array([1.0, 2.0, 3.0, 4.0])
    assert aggregate.concordance_index(t, t, np.ones(4)) == 1.0
    assert aggregate.concordance_index(t, -t, np.ones(4)) == 0.0
    assert aggregate.concordance_index(t, np.zeros(4), np.ones(4)) == 0.5
    with pytest.raises(ZeroDivisionError):
        aggregate.concordance_index(t, t, np.zeros(4))           # all censored: no admissible pair
    # a death and a censoring at the same time are comparable (the death is earlier), two deaths are not
    assert aggregate.concordance_counts([5.0, 5.0], [0.0, 1.0], [1, 0]) == (1, 1, 0)
    assert aggregate.concordance_counts([5.0, 5.0], [0.0, 1.0], [1, 1]) == (0, 0, 0)
    # through get_survival_CI (reference call: concordance_index(survival_months, -score, vital_status))
    rng = np.random.default_rng(5)
    ids = [f"case{i % 40:03d}" for i in range(400)]
    out = torch.tensor(rng.normal(size=(400, 1)).astype(np.float32)).cuda()
    sm = np.repeat(rng.uniform(1, 100, 40), 10).reshape(10, 40).T.reshape(-1)[:400]
    sm = np.array([sm[i % 40] for i in range(400)])
    vs = np.array([(i % 40) % 3 != 0 for i in range(400)]).astype(np.int64)
    ci, df = aggregate.get_survival_CI(out, ids, sm, vs)
    ref = cindex_oracle.concordance_index(df["survival_months"], -df["score"], df["vital_status"])
    assert ci == ref and 0.0 <= ci <= 1.0


@pytest.mark.parametrize("ties", [False, True])
def test_concordance_dominance_path_matches_pair_kernel(ties, monkeypatch):
    """Large cohorts take the O(n S) dominance count (cindex_dominance_kernel): the three integer counts must equal the
    O(n^2) pair kernel's, which is bit-exact with the oracle (ties in time and in prediction included)."""
    from multimodalbrainsurvival_b200 import aggregate
    from oracle import cindex_oracle
    rng = np.random.default_rng(5 + int(ties))
    n = 30_000
    t = np.floor(rng.uniform(0, 200, n)) if ties else rng.uniform(0, 200, n)
    p = np.round(rng.standard_normal(n), 1) if ties else rng.standard_normal(n)
    e = (rng.uniform(size=n) < 0.6).astype(np.int64)
    want = aggregate.concordance_counts(t, p, e)                      # n <= PAIRWISE_MAX_N: the pair kernel
    monkeypatch.setattr(aggregate, "PAIRWISE_MAX_N", 1000)
    got = aggregate.concordance_counts(t, p, e)
    assert got == want and want[0] > 0
    small = slice(0, 700)                                              # and the oracle itself on a small slice
    monkeypatch.setattr(aggregate, "PAIRWISE_MAX_N", 100)
    assert aggregate.concordance_counts(t[small], p[small], e[small]) == \
        cindex_oracle.concordance_counts(t[small], p[small], e[small])
    # degenerate inputs: nobody died / everybody died at the same time
    assert aggregate.concordance_counts(t[small], p[small], np.zeros(700)) == (0, 0, 0)
    assert aggregate.concordance_counts(np.ones(700), p[small], np.ones(700)) == (0, 0, 0)
