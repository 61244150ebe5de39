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
