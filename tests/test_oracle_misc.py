"""Pin the aggregation / ResNet / MLP oracles to vectors produced by the reference."""
import numpy as np
import pytest
import torch

from conftest import det_input
from oracle import aggregate_oracle, cindex_oracle, mlp_oracle, resnet_oracle


def test_case_mean_features_matches_reference(golden):
    g = golden("aggregate_reference.npz")
    cases = [str(c) for c in g["cases"]]
    order = [str(c) for c in g["case_uniques"]]
    got = aggregate_oracle.case_mean_features(g["features"], cases, order)
    assert got.dtype == g["features_final"].dtype
    np.testing.assert_array_equal(got, g["features_final"])
    uniq, seg = aggregate_oracle.group_ids(cases, order=order)
    mean, cnt = aggregate_oracle.segment_mean(g["features"], seg, len(uniq))
    np.testing.assert_allclose(mean, g["features_final"], rtol=2e-6, atol=2e-6)
    assert cnt.sum() == len(cases)


def test_survival_grouping_matches_reference(golden):
    g = golden("aggregate_reference.npz")
    cases = [str(c) for c in g["cases"]]
    ids, score, sm, vs = aggregate_oracle.survival_grouping(g["outputs"], cases, g["survival"], g["vital"])
    assert ids == [str(c) for c in g["ci_ids"]]
    np.testing.assert_array_equal(score, g["ci_score"])
    np.testing.assert_array_equal(sm, g["ci_survival"])
    np.testing.assert_array_equal(vs, g["ci_vital"])


def test_cindex_basic():
    t = np.array([1, 2, 3, 4.0])
    assert cindex_oracle.concordance_index(t, t, np.ones(4)) == 1.0
    assert cindex_oracle.concordance_index(t, -t, np.ones(4)) == 0.0
    assert cindex_oracle.concordance_index(t, np.zeros(4), np.ones(4)) == 0.5


def test_cindex_tie_rules_and_bruteforce():
    """The restated lifelines rule: the earlier exit must be a death; equal-time deaths are not comparable,
    a death and a censoring at the same time are.  Checked against a literal double loop."""
    assert cindex_oracle.concordance_counts([5.0, 5.0], [0.0, 1.0], [1, 0]) == (1, 1, 0)
    assert cindex_oracle.concordance_counts([5.0, 5.0], [0.0, 1.0], [1, 1]) == (0, 0, 0)
    assert cindex_oracle.concordance_counts([5.0, 5.0], [0.0, 1.0], [0, 0]) == (0, 0, 0)
    assert cindex_oracle.concordance_counts([1.0, 2.0], [3.0, 3.0], [1, 0]) == (1, 0, 1)
    assert cindex_oracle.concordance_counts([1.0, 2.0], [3.0, 3.0], [0, 1]) == (0, 0, 0)   # censored first: unusable
    rng = np.random.default_rng(3)
    for n in (1, 7, 60):
        t = np.round(rng.uniform(0, 10, n))
        p = np.round(rng.normal(size=n), 1)
        e = rng.random(n) < 0.6
        pairs = correct = tied = 0
        for i in range(n):
            for j in range(n):
                if e[j] and (t[j] < t[i] or (t[j] == t[i] and not e[i])):
                    pairs += 1
                    correct += p[j] < p[i]
                    tied += p[j] == p[i]
        assert cindex_oracle.concordance_counts(t, p, e) == (pairs, int(correct), int(tied))


def _check_rng_fingerprint(sd, g):
    for k, v in zip(g["fp_keys"], g["fp_vals"]):
        got = float(sd[str(k)].double().abs().sum())
        if abs(got - float(v)) > 1e-6 * abs(float(v)):
            pytest.skip("torch CPU RNG stream differs from the one that produced the golden weights")


def test_resnet_oracle_matches_reference(golden):
    g = golden("resnet_reference.npz")
    sd = resnet_oracle.init_state_dict(seed=1111)
    _check_rng_fingerprint(sd, g)
    x = torch.tensor(det_input((2, 3, 224, 224)))
    torch.set_num_threads(max(1, torch.get_num_threads()))
    f = resnet_oracle.forward_extract(sd, x).numpy()
    ref = g["features"]
    assert f.shape == (2, 2048)
    assert np.abs(f - ref).max() <= 2e-4 * np.abs(ref).max()


def test_resnet_oracle_bf16_emulation_within_tolerance(golden):
    g = golden("resnet_reference.npz")
    sd = resnet_oracle.init_state_dict(seed=1111)
    _check_rng_fingerprint(sd, g)
    x = torch.tensor(det_input((1, 3, 224, 224)))
    f = resnet_oracle.forward_extract(sd, x, emulate_bf16=True).numpy()
    ref = g["features"][:1]
    rel = np.linalg.norm(f - ref) / np.linalg.norm(ref)
    assert rel < 1e-2, rel     # north_star: bf16 features within 1e-2 relative


def _rna_model_weights(seed):
    import torch.nn as nn
    torch.manual_seed(seed)
    rna = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048))
    head = nn.Sequential(nn.Linear(2048, 1))
    return rna, head


def test_mlp_oracle_matches_reference(golden):
    g = golden("mlp_reference.npz")
    rna, head = _rna_model_weights(1111)
    sd = {"rna_mlp." + k: v for k, v in rna.state_dict().items()}
    x = torch.tensor(det_input((6, 12778), a=0.11))
    outs = mlp_oracle.mlp_forward(x, mlp_oracle.rna_layers(sd))
    feat = outs[-1]
    y = feat @ head[0].weight.t() + head[0].bias
    np.testing.assert_allclose(feat[:, :64].detach().numpy(), g["rna_feat_head"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y.detach().numpy(), g["rna_out"], rtol=1e-4, atol=1e-5)


def test_resnet_train_oracle_matches_reference_golden(golden):
    """oracle.resnet_oracle.train_step (model.train(): batch-statistics BatchNorm, autograd through layer4) against
    the reference module's own step (tools/make_golden.py gen_resnet_train): features, layer4 gradients, running
    statistics, all to fp32 round-off."""
    from conftest import det_input
    import torch
    g = golden("resnet_train_reference.npz")
    sd = resnet_oracle.init_state_dict(seed=2222, bn3_gamma_scale=0.1)
    x = torch.tensor(det_input((4, 3, 224, 224), a=0.7))
    gw = torch.tensor(det_input((4, 2048), a=1.3))
    f, grads, stats = resnet_oracle.train_step(sd, x, gw)
    ref = g["features"]
    if np.linalg.norm(f.numpy() - ref) > 1e-3 * np.linalg.norm(ref):
        fp = float(sd["conv1.weight"].double().abs().sum())
        pytest.skip(f"torch CPU RNG stream differs from the one that produced the golden weights (fingerprint {fp})")
    assert np.linalg.norm(f.numpy() - ref) <= 1e-5 * np.linalg.norm(ref)
    for key in g.files:
        if key.startswith("grad/"):
            gr = grads[key[5:]].flatten()
            sample = (gr if gr.numel() <= 4096 else gr[::997]).numpy()
            assert np.linalg.norm(sample - g[key]) <= 1e-3 * np.linalg.norm(g[key]) + 1e-12, key
        elif key.startswith("stat/") and not key.endswith("num_batches_tracked"):
            assert np.allclose(stats[key[5:]].numpy(), g[key], rtol=1e-5, atol=1e-6), key
