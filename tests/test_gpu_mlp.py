"""GPU parity of the fused MLP paths (multimodalbrainsurvival_b200/mlp.py over the tcgen05 GEMM):
inference vs the reference golden outputs / fp32 oracle, training (forward, dgrad, wgrad, Philox
dropout) vs torch autograd on an emulation that reuses the kernel's own dropout masks.
Tolerance: bf16 operands, fp32 accumulation -> 2e-2 relative to the tensor's max."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import det_input
from oracle import mlp_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rna(seed):
    torch.manual_seed(seed)
    rna = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048))
    head = nn.Sequential(nn.Linear(2048, 1))
    return rna, head


def _close(a, b, tol, name):
    """Relative Frobenius error <= tol and max error <= 5 tol of the largest entry.  (A hidden unit whose
    pre-activation is within rounding of 0 may take the other ReLU branch than the reference; that moves
    single entries by a whole term, which a pure max-norm test would flag although the tensors agree.)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    assert rel <= tol, f"{name}: relative error {rel:.4g}"
    assert err <= 5 * tol * scale + 1e-6, f"{name}: max err {err:.4g} vs scale {scale:.4g}"


def test_rna_inference_matches_reference_golden(golden):
    from multimodalbrainsurvival_b200 import mlp, models
    g = golden("mlp_reference.npz")
    rna, head = _rna(1111)
    model = models.RNAOnlyModel(rna, head).to(DEV).eval()
    x = torch.tensor(det_input((6, 12778), a=0.11), device=DEV)
    with torch.no_grad():
        y = model(x)
        feat = model.extract(x)
    assert mlp._ENGINES, "the fused inference engine did not run"
    _close(feat[:, :64], torch.tensor(g["rna_feat_head"]), 2e-2, "rna features")
    _close(y, torch.tensor(g["rna_out"]), 2e-2, "rna output")


def test_early_fusion_inference_matches_reference_golden(golden):
    from multimodalbrainsurvival_b200 import models
    g = golden("mlp_reference.npz")
    torch.manual_seed(2222)
    early = nn.Sequential(nn.Dropout(), nn.Linear(4096, 2048), nn.ReLU(), nn.Dropout(), nn.Linear(2048, 200),
                          nn.ReLU(), nn.Dropout(), nn.Linear(200, 1))
    acc = models.accelerate(early).to(DEV).eval()
    xe = torch.tensor(det_input((5, 4096), a=0.23), device=DEV)
    with torch.no_grad():
        ye = acc(xe)
    _close(ye, torch.tensor(g["early_out"]), 2e-2, "early fusion output")


@pytest.mark.parametrize("dims,M", [((300, 256, 128, 1), 37), ((12778, 4096, 2048), 128)])
def test_training_without_dropout_matches_torch_autograd(dims, M):
    from multimodalbrainsurvival_b200 import mlp
    torch.manual_seed(5)
    mods = []
    for i in range(len(dims) - 1):
        mods += [nn.Dropout(0.0), nn.Linear(dims[i], dims[i + 1])]
        if i < len(dims) - 2:
            mods.append(nn.ReLU())
    seq = nn.Sequential(*mods).to(DEV).train()
    x = torch.randn(M, dims[0], device=DEV)
    R = torch.randn(M, dims[-1], device=DEV)
    out = mlp.run_mlp(seq, x)
    assert mlp._TRAIN_ENGINES, "the fused training engine did not run"
    (out * R).sum().backward()
    got = {n: p.grad.clone() for n, p in seq.named_parameters()}
    seq.zero_grad()
    # reference: torch autograd on the same graph with bf16-rounded GEMM operands (the ReLU masks of a
    # mixed-precision forward legitimately differ from an fp32 forward for pre-activations near zero)
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    h = x
    lins = [m for m in seq if isinstance(m, nn.Linear)]
    for i, lin in enumerate(lins):
        h = bf(h) @ bf(lin.weight).t() + lin.bias
        if i < len(lins) - 1:
            h = torch.relu(h)
    ref_out = h
    (ref_out * R).sum().backward()
    _close(out, ref_out, 2e-2, "forward")
    for n, p in seq.named_parameters():
        _close(got[n], p.grad, 3e-2, f"grad {n}")
    fp32_out = nn.Sequential.forward(seq, x)
    _close(out, fp32_out, 3e-2, "forward vs fp32 module graph")


def test_training_with_dropout_is_consistent_with_its_own_masks():
    """p = 0.5 / 0.8: recover the Philox masks from the engine's dropped activations and check
    forward, dx, dW, db against torch autograd on the same masked graph; mask rates ~ 1-p."""
    from multimodalbrainsurvival_b200 import mlp
    torch.manual_seed(9)
    seq = nn.Sequential(nn.Dropout(0.8), nn.Linear(512, 256), nn.ReLU(), nn.Dropout(0.5), nn.Linear(256, 64)).to(DEV).train()
    M = 200
    x = (torch.rand(M, 512, device=DEV) + 0.5).requires_grad_(True)   # strictly positive: mask = (dropped != 0)
    R = torch.randn(M, 64, device=DEV)
    out = mlp.run_mlp(seq, x)
    (out * R).sum().backward()
    eng = next(iter(mlp._TRAIN_ENGINES.values())) if len(mlp._TRAIN_ENGINES) == 1 else \
        [e for e in mlp._TRAIN_ENGINES.values() if e.m == M and e.need_dx][-1]
    mask0 = (eng.hin[0][:, :512].float() != 0).float()
    assert abs(mask0.mean().item() - 0.2) < 0.02
    got = {n: p.grad.clone() for n, p in seq.named_parameters()}
    got_dx = x.grad.clone()
    # emulate with the recovered masks (bf16 rounding of operands like the kernels)
    lin1, lin2 = seq[1], seq[4]
    xr = x.detach().clone().requires_grad_(True)
    h0 = (xr * mask0 * 5.0)
    a1 = torch.relu(h0.to(torch.bfloat16).float() @ lin1.weight.to(torch.bfloat16).float().t() + lin1.bias)
    a1b = eng.act[0][:, :256].float()
    mask1 = ((eng.hin[1][:, :256].float() != 0) | (a1b == 0)).float()      # where a1 == 0 the mask is unobservable
    assert abs(((eng.hin[1][:, :256].float() != 0).float().sum() / (a1b != 0).float().sum()).item() - 0.5) < 0.03
    h1 = a1 * mask1 * 2.0
    ref_out = h1.to(torch.bfloat16).float() @ lin2.weight.to(torch.bfloat16).float().t() + lin2.bias
    _close(out, ref_out, 2e-2, "forward with dropout")
    seq.zero_grad()
    (ref_out * R).sum().backward()
    _close(got["4.weight"], lin2.weight.grad, 3e-2, "dW2")
    _close(got["4.bias"], lin2.bias.grad, 3e-2, "db2")
    _close(got["1.weight"], lin1.weight.grad, 3e-2, "dW1")
    _close(got["1.bias"], lin1.bias.grad, 3e-2, "db1")
    _close(got_dx, xr.grad, 3e-2, "dx")
    assert float((got_dx[mask0 == 0]).abs().max()) == 0.0


def test_rna_model_train_step_reduces_cox_loss():
    """End to end (config 1 shape): RNAOnlyModel + cox_loss + Adam, all through the kernels."""
    from multimodalbrainsurvival_b200 import models
    rna, head = _rna(3)
    model = models.RNAOnlyModel(rna, head).to(DEV).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator(device=DEV).manual_seed(3333)
    x = torch.randn(128, 12778, device=DEV, generator=g)
    t = torch.rand(128, device=DEV, generator=g) * 200
    e = (torch.rand(128, device=DEV, generator=g) < 0.6).float()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = models.cox_loss(model(x).view(-1), t, e)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and min(losses[3:]) < losses[0], losses
