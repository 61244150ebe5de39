"""Pin the Cox oracle to the reference: SURVEY §4 KAT + golden vectors produced by
the reference's cox_loss (tools/make_golden.py)."""
import numpy as np
import pytest

from oracle import cox_oracle

KAT_T = np.array([5, 3, 5, 1, 3, 5, 0, 0], np.float32)
KAT_S = np.array([0.5, -1, 2, 0, 1.5, -0.5, 0.25, -2], np.float32)
KAT_E = np.array([1, 0, 1, 1, 0, 1, 1, 0], np.float32)
KAT_LOSS = 1.0347948074
KAT_GRAD = np.array([0.069751039, 0.005708719, 0.187636316, -0.109482095,
                     0.069546439, -0.107726723, -0.115433700, 0.0])


def test_kat_from_survey():
    loss, grad, perm = cox_oracle.cox_loss_and_grad(KAT_S, KAT_T, KAT_E)
    assert abs(loss - KAT_LOSS) < 2e-7
    np.testing.assert_allclose(grad, KAT_GRAD, atol=2e-8)
    assert perm.tolist() == [0, 2, 5, 1, 4, 3, 6, 7]


def _cases(g):
    return sorted({k.split("/")[0] for k in g.files})


def test_against_reference_golden(golden):
    g = golden("cox_reference.npz")
    for name in _cases(g):
        s, t, e = g[name + "/scores"], g[name + "/times"], g[name + "/status"]
        loss, grad, perm = cox_oracle.cox_loss_and_grad(s, t, e)
        assert np.array_equal(perm, g[name + "/perm"]), name          # bit-exact order
        ref_loss = float(g[name + "/loss"])
        # 1e-5 relative (north_star) + an absolute floor of 2 fp32 ulps of the O(1)
        # intermediates (s~, log(C+eps)) whose difference forms each term.
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss) + 2.4e-7, name
        loss32, _ = cox_oracle.cox_forward(s, t, e, np.float32)     # fp32-faithful mode
        assert abs(float(loss32) - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-9, name
        ref_grad = g[name + "/grad"].astype(np.float64)
        scale = max(np.abs(ref_grad).max(), 1e-12)
        assert np.abs(grad - ref_grad).max() <= 1e-5 * scale + 1e-10, name


def test_key_transform_orders_like_sort(golden):
    g = golden("cox_reference.npz")
    for name in _cases(g):
        t = g[name + "/times"]
        key = cox_oracle.order_key_u32(t)
        perm = np.argsort(key, kind="stable")
        assert np.array_equal(perm, g[name + "/perm"]), name


def test_key_transform_special_values():
    t = np.array([np.inf, -np.inf, 0.0, -0.0, 1e-45, -1e-45, 3.4e38, -3.4e38, 1.0, 1.0], np.float32)
    key = cox_oracle.order_key_u32(t)
    assert np.array_equal(np.argsort(key, kind="stable"), np.argsort(-t, kind="stable"))
    assert key[2] == key[3]


def test_edge_cases():
    loss, _ = cox_oracle.cox_forward(np.array([0.7], np.float32), np.array([3.0], np.float32),
                                     np.array([1.0], np.float32))
    assert abs(loss - np.log(1 + 1e-5)) < 1e-12
    loss, _ = cox_oracle.cox_forward(np.random.randn(10), np.arange(10), np.zeros(10))
    assert loss == 0.0


@pytest.mark.parametrize("n", [3, 50, 2000])
def test_grad_matches_finite_difference(n):
    rng = np.random.default_rng(n)
    s = rng.standard_normal(n).astype(np.float32)
    t = rng.integers(0, 30, n).astype(np.float32)
    e = (rng.uniform(size=n) < 0.6).astype(np.float32)
    g = cox_oracle.cox_backward(s, t, e)
    assert abs(g.sum()) < 1e-12  # shift invariance
    for i in rng.integers(0, n, 3):
        h = 1e-3
        sp, sm = s.copy(), s.copy()
        sp[i] += h
        sm[i] -= h
        fd = (cox_oracle.cox_forward(sp, t, e)[0] - cox_oracle.cox_forward(sm, t, e)[0]) / (float(sp[i]) - float(sm[i]))
        assert abs(fd - g[i]) < 1e-6 + 1e-4 * abs(g[i])
