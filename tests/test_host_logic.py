"""CPU tests: the C-ABI library loads and exports every declared symbol, fails loudly without
a GPU, and the drop-in modules keep the reference's import surface / state_dict layout."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mmbs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmbs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from multimodalbrainsurvival_b200 import _lib
    l = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(l, s), f"libmmbs.so does not export {s}"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert l.mmbs_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    from multimodalbrainsurvival_b200 import _lib, cox
    l = _lib.lib()
    assert l.mmbs_device_check() == -4                      # MMBS_ERR_DEVICE
    assert b"no CPU fallback" in l.mmbs_last_error()
    buf = (ctypes.c_float * 8)()
    rc = l.mmbs_cox_forward(buf, buf, buf, 8, buf, buf, buf, buf, buf, buf, 1 << 20, None)
    assert rc == -4
    with pytest.raises(_lib.MMBSError):
        _lib.check(rc, "mmbs_cox_forward")
    with pytest.raises(RuntimeError, match="CUDA"):
        cox.cox_loss(torch.zeros(4, requires_grad=True), torch.ones(4), torch.ones(4))
    from multimodalbrainsurvival_b200 import aggregate
    with pytest.raises(RuntimeError, match="CUDA"):
        aggregate.segmented_mean(torch.zeros(4, 2), torch.zeros(4, dtype=torch.int32), 2)


def test_workspace_size_functions_are_pure_host():
    from multimodalbrainsurvival_b200 import _lib
    l = _lib.lib()
    a, b = l.mmbs_cox_workspace_bytes(1000), l.mmbs_cox_workspace_bytes(10_000_000)
    assert 0 < a < b and b < 400_000_000          # ~16 B/sample sort buffers + look-back words
    assert l.mmbs_segmented_mean_workspace_bytes(1000, 10) > 0


def test_dropin_import_surface():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for k in ("models", "resnet"):
            sys.modules.pop(k, None)
        import models
        import resnet
        for name in ("AggregationModel", "Identity", "TanhAttention", "CoxLoss", "PatchBagDataset", "NLLSurvLoss",
                     "cox_loss", "RNAOnlyModel", "HistopathologyRNAModel", "AggregationProjectModel",
                     "BagHistopathologyRNAModel"):
            assert hasattr(models, name), name
        for name in ("resnet50", "ResNet", "Bottleneck", "BasicBlock", "resnet18", "resnet34", "resnet101",
                     "resnet152", "RNfour", "RNone", "ResNetProject"):
            assert hasattr(resnet, name), name
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for k in ("models", "resnet"):
            sys.modules.pop(k, None)


def test_state_dict_contract_and_torch_path_matches_oracle():
    from multimodalbrainsurvival_b200 import models, resnet
    from oracle import resnet_oracle
    net = resnet.resnet50()
    assert sum(p.numel() for p in net.parameters()) == 25_557_032
    model = models.AggregationModel(net, models.Identity(), 2048, 2048, 1)
    sd = model.state_dict()
    assert len(sd) == 322 and sd["fc.weight"].shape == (1, 2048)
    assert sd["resnet.layer2.0.downsample.0.weight"].shape == (512, 256, 1, 1)
    osd = resnet_oracle.init_state_dict(seed=3)
    net.load_state_dict(osd, strict=True)                    # reference key names load both ways
    net.eval()
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with pytest.raises(RuntimeError, match="CUDA tensor"):   # no CPU path behind the public entry point
        net.forward_extract(x)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        model.eval()(x.view(1, 1, 3, 224, 224))
    with torch.no_grad():
        f = net._features_torch(x)     # the stock module graph itself: the tree is wired like the reference's
    ref = resnet_oracle.forward_extract(osd, x)
    assert float((f - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_state_dict_keys_equal_reference():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import make_golden
    finally:
        sys.path.pop(0)
    from multimodalbrainsurvival_b200 import models, resnet
    ref_resnet = make_golden.load_ref("5_JointFusion/resnet.py", "ref_resnet_cmp")
    ref_models = make_golden.load_ref("5_JointFusion/models.py", "ref_models_cmp")
    rna = lambda: nn.Sequential(nn.Dropout(), nn.Linear(32, 16), nn.ReLU(), nn.Dropout(), nn.Linear(16, 8))  # noqa: E731
    head = lambda: nn.Sequential(nn.Dropout(0.8), nn.Linear(2056, 1))  # noqa: E731
    ours = models.BagHistopathologyRNAModel(resnet.resnet50(), rna(), head())
    theirs = ref_models.BagHistopathologyRNAModel(ref_resnet.resnet50(), rna(), head())
    a, b = ours.state_dict(), theirs.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    theirs.load_state_dict(a)
    ours.load_state_dict(b)


def test_mlp_pattern_matching_and_cpu_path():
    from multimodalbrainsurvival_b200 import mlp, models
    seq = nn.Sequential(nn.Dropout(), nn.Linear(12, 8), nn.ReLU(), nn.Dropout(), nn.Linear(8, 4))
    layers = mlp._parse(seq)
    assert [(l.in_features, l.out_features, r) for l, r in layers] == [(12, 8, True), (8, 4, False)]
    assert mlp._parse(nn.Sequential(nn.Linear(4, 4), nn.Tanh())) is None
    m = models.RNAOnlyModel(seq, nn.Sequential(nn.Linear(4, 1))).eval()
    x = torch.randn(5, 12)
    torch.testing.assert_close(m(x), m.final_mlp(seq(x)))     # CPU tensors: module graph
    acc = models.accelerate(seq)
    assert list(acc.state_dict().keys()) == list(seq.state_dict().keys())
    torch.testing.assert_close(acc.eval()(x), seq.eval()(x))


def test_nll_loss_matches_manual():
    from multimodalbrainsurvival_b200 import models
    h = torch.tensor([[0.2, -0.3, 0.5], [1.0, 0.1, -0.7]])
    y = torch.tensor([1, 2])
    c = torch.tensor([0.0, 1.0])
    loss = models.NLLSurvLoss()(h, y, c)
    hz = torch.sigmoid(h)
    s = torch.cumprod(1 - hz, 1)
    manual = (-(torch.log(s[0, 0]) + torch.log(hz[0, 1])) - torch.log(s[1, 2])) / 2
    torch.testing.assert_close(loss, manual)


def test_accelerate_optimizer_keeps_the_torch_object_and_state_layout():
    """Host logic of optim.accelerate_optimizer: same object, same param groups, CPU parameters step through torch's
    own Adam (the fused kernel needs CUDA tensors), state_dict layout identical to a stock optimizer's."""
    from multimodalbrainsurvival_b200 import optim
    torch.manual_seed(0)
    m1, m2 = nn.Linear(5, 3), nn.Linear(5, 3)
    m2.load_state_dict(m1.state_dict())
    groups = lambda m: [{"params": [m.weight], "lr": 1e-3}, {"params": [m.bias], "lr": 1e-2}]  # noqa: E731
    o1 = torch.optim.Adam(groups(m1), weight_decay=1e-5)
    o2 = torch.optim.Adam(groups(m2), weight_decay=1e-5)
    assert optim.accelerate_optimizer(o1) is o1 and optim.accelerate_optimizer(o1) is o1
    x = torch.randn(4, 5)
    for _ in range(3):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad()
            m(x).pow(2).sum().backward()
            o.step()
    assert torch.equal(m1.weight, m2.weight) and torch.equal(m1.bias, m2.bias)
    s1, s2 = o1.state_dict(), o2.state_dict()
    assert s1["param_groups"] == s2["param_groups"]
    assert {k: sorted(v) for k, v in s1["state"].items()} == {k: sorted(v) for k, v in s2["state"].items()}


def test_feature_csv_writer_is_byte_identical_to_savetxt(tmp_path):
    """np.savetxt(path, features, delimiter=",") of 4_HistoPath_extractfeatures.py:184-192 vs the multi-threaded
    writer (csrc/csv.cu, host code): identical bytes, including non-finite values, 1-D input and float32 input."""
    from multimodalbrainsurvival_b200 import aggregate
    rng = np.random.default_rng(3)
    cases = [rng.standard_normal((37, 2048)) * 10.0 ** rng.integers(-30, 30, (37, 2048)),
             np.array([[0.0, -0.0, np.inf, -np.inf, np.nan, 1e-320, 1.7976931348623157e308, 5e-324]]),
             rng.standard_normal(11), rng.standard_normal((700, 3)).astype(np.float32)]
    for i, a in enumerate(cases):
        ref, out = tmp_path / f"ref{i}.csv", tmp_path / f"out{i}.csv"
        np.savetxt(ref, a, delimiter=",")
        aggregate.save_features_csv(out, a, threads=1 + i)
        assert ref.read_bytes() == out.read_bytes(), f"case {i}"
