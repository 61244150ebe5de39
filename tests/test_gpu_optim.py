"""GPU parity of the fused multi-tensor Adam (csrc/adam.cu behind torch.optim.Adam's interface)
vs torch.optim.Adam itself: same parameters, same gradients, 10 steps.  Tolerance 1e-6 relative to the largest entry
(fp32, same operation order as torch's foreach implementation; FMA contraction differs)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(4096, 1277), (4096,), (2048, 4096), (2048,), (1, 2048), (1,), (7, 3, 3, 5), (333,), (1000003,)]


def _make(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in SHAPES]


def _groups(ps):
    # the reference's layout: several groups with their own learning rate, one shared weight decay
    # (2_GeneExpression/1_GeneExpress_train.py:303-305)
    return [{"params": ps[:4], "lr": 1e-3}, {"params": ps[4:6], "lr": 5e-2, "betas": (0.8, 0.95)},
            {"params": ps[6:], "lr": 3e-4, "weight_decay": 0.0, "eps": 1e-6}]


def _close(a, b, tol, what):
    err = float((a.double() - b.double()).abs().max())
    scale = float(b.double().abs().max()) + 1e-30
    assert err <= tol * scale, f"{what}: max err {err:.3g} vs scale {scale:.3g}"


def test_fused_adam_matches_torch_adam_over_10_steps():
    from multimodalbrainsurvival_b200 import _lib, optim
    pa, pb = _make(1), _make(1)
    oa = optim.accelerate_optimizer(torch.optim.Adam(_groups(pa), weight_decay=1e-5))
    ob = torch.optim.Adam(_groups(pb), weight_decay=1e-5)
    g = torch.Generator(device=DEV).manual_seed(2)
    for step in range(10):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g) * (0.1 + step)
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 4:            # a parameter without a gradient is skipped, like torch does
            pa[3].grad = pb[3].grad = None
        l0 = _lib.launch_count()
        oa.step()
        assert _lib.launch_count() == l0 + 1, "one fused launch per step"
        ob.step()            # (after step 4 group 0 holds parameters at two step counts: two hyper-parameter rows)
    for i, (a, b) in enumerate(zip(pa, pb)):
        _close(a, b, 1e-6, f"param {i} {tuple(a.shape)}")
        sa, sb = oa.state[a], ob.state[b]
        assert float(sa["step"]) == float(sb["step"])
        _close(sa["exp_avg"], sb["exp_avg"], 1e-6, f"exp_avg {i}")
        _close(sa["exp_avg_sq"], sb["exp_avg_sq"], 1e-6, f"exp_avg_sq {i}")


def test_fused_step_bumps_version_counters():
    """The packed bf16 weight caches of the MLP / ResNet engines are keyed on `_version`: a fused step that left it
    alone would let the next forward run on stale weights (found by tests/test_gpu_script_rna.py)."""
    from multimodalbrainsurvival_b200 import optim
    p = torch.randn(1000, device=DEV, requires_grad=True)
    o = optim.accelerate_optimizer(torch.optim.Adam([p], lr=1e-2))
    p.grad = torch.ones_like(p)
    v0 = p._version
    o.step()
    assert p._version > v0 and o.state[p]["exp_avg"]._version > 0


def test_state_dict_round_trips_between_fused_and_stock():
    from multimodalbrainsurvival_b200 import optim
    pa = _make(3)
    oa = optim.accelerate_optimizer(torch.optim.Adam(_groups(pa), weight_decay=1e-5))
    g = torch.Generator(device=DEV).manual_seed(4)
    for _ in range(3):
        for a in pa:
            a.grad = torch.randn(a.shape, device=DEV, generator=g)
        oa.step()
    sd = copy.deepcopy(oa.state_dict())
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    pb = [a.detach().clone().requires_grad_(True) for a in pa]
    ob = torch.optim.Adam(_groups(pb), weight_decay=1e-5)
    # (load_state_dict does not clone: every optimizer gets its own copy of the checkpoint's tensors)
    ob.load_state_dict(copy.deepcopy(sd))          # a checkpoint written by the fused path resumes on stock torch
    pc = [a.detach().clone().requires_grad_(True) for a in pa]
    oc = optim.accelerate_optimizer(torch.optim.Adam(_groups(pc), weight_decay=1e-5))
    oc.load_state_dict(copy.deepcopy(sd))          # ... and the other way round
    for _ in range(2):
        for a, b, c in zip(pa, pb, pc):
            gr = torch.randn(a.shape, device=DEV, generator=g)
            a.grad, b.grad, c.grad = gr.clone(), gr.clone(), gr.clone()
        oa.step(); ob.step(); oc.step()
    for i, (a, b, c) in enumerate(zip(pa, pb, pc)):
        _close(a, b, 1e-6, f"param {i} fused vs resumed stock")
        _close(c, b, 1e-6, f"param {i} resumed fused vs resumed stock")


def test_uncovered_configurations_run_torchs_own_step():
    from multimodalbrainsurvival_b200 import _lib, optim
    p = torch.randn(100, device=DEV, requires_grad=True)
    q = p.detach().clone().requires_grad_(True)
    oa = optim.accelerate_optimizer(torch.optim.Adam([p], lr=1e-2, amsgrad=True))
    ob = torch.optim.Adam([q], lr=1e-2, amsgrad=True)
    p.grad, q.grad = torch.ones_like(p), torch.ones_like(q)
    l0 = _lib.launch_count()
    oa.step(); ob.step()
    assert _lib.launch_count() == l0 and torch.equal(p, q)
    sgd = torch.optim.SGD([p], lr=0.1)
    assert optim.accelerate_optimizer(sgd) is sgd and not getattr(sgd, "_mmbs_accelerated", False)


def test_cached_tables_follow_changing_gradient_sets_and_reloaded_state():
    """The fused step caches its tensor table per optimizer: a parameter that skips steps (grad None), a changed set of
    parameters with gradients, and a load_state_dict in the middle must all keep matching torch's own Adam."""
    from multimodalbrainsurvival_b200 import optim
    torch.manual_seed(3)
    ref_p = [torch.randn(257, 33, device=DEV, requires_grad=True), torch.randn(1000, device=DEV, requires_grad=True),
             torch.randn(64, 64, device=DEV, requires_grad=True)]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=1e-2, weight_decay=1e-3)
    ours = optim.accelerate_optimizer(torch.optim.Adam(our_p, lr=1e-2, weight_decay=1e-3))
    g = torch.Generator(device=DEV).manual_seed(1)
    for step in range(9):
        active = [0, 1, 2] if step % 3 == 0 else ([0, 2] if step % 3 == 1 else [1])   # parameters that get a gradient
        for i in range(3):
            grad = torch.randn(ref_p[i].shape, device=DEV, generator=g) if i in active else None
            ref_p[i].grad = grad
            our_p[i].grad = None if grad is None else grad.clone()
        if step == 5:   # reload the state (new state objects): the cache must notice
            ours.load_state_dict(ours.state_dict())
        ref.step()
        ours.step()
    for a, b in zip(ref_p, our_p):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), float((a - b).abs().max())
    for a, b in zip(ref_p, our_p):
        assert float(ref.state[a]["step"]) == float(ours.state[b]["step"])
