"""Fused multi-tensor Adam behind ``torch.optim.Adam``'s interface (csrc/adam.cu).

The reference's scripts build their optimizers themselves::

    optimizer = torch.optim.Adam([{'params': rna_mlp.parameters(), 'lr': lr_rna}, {'params': mlp.parameters(), ...}],
                                 weight_decay=config['weight_decay'])
        /root/reference/2_GeneExpression/1_GeneExpress_train.py:303-305
        /root/reference/1_HistoPathology/2_HistoPath_train.py:558
        /root/reference/5_JointFusion/1_JointFusion_train.py:413-416

so the replacement keeps that object: ``accelerate_optimizer(opt)`` swaps only what ``opt.step()`` executes.  Parameter
groups, hyper-parameters, ``opt.state`` (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter, exactly torch's layout)
and therefore ``state_dict()`` / ``load_state_dict()`` stay torch's own; the arithmetic is one pass of
``mmbs_adam_step`` over (p, g, m, v) instead of torch's ~9 foreach passes.  ``install()`` (or ``MMBS_FUSED_ADAM=1``
with the drop-in ``models.py``) does the same for every ``torch.optim.Adam`` the unmodified scripts create.

Covered: fp32 CUDA parameters, L2 weight decay, any number of groups, ``amsgrad=False``, ``maximize=False``.
Anything else (CPU parameters, amsgrad, sparse gradients, a closure) runs torch's stock ``step`` - that code is the
script's own optimizer, not a fallback of the hot path.
"""
from __future__ import annotations

import math
import os

import torch

from . import _lib

_MAX_GROUPS = 8


def _bump_versions(tensors) -> None:
    try:
        torch._C._increment_version(tensors)          # torch >= 2.3: an iterable of tensors
    except TypeError:
        for t in tensors:
            torch._C._increment_version(t)


def _fusable_group(group) -> bool:
    return not (group.get("amsgrad") or group.get("maximize") or group.get("differentiable")
                or group.get("capturable"))


def _fused_step(opt) -> bool:
    """One Adam step of every parameter that has a gradient.  Returns False when this optimizer (in its current
    state) is not covered and nothing was touched."""
    groups = opt.param_groups
    if len(groups) > _MAX_GROUPS or not all(_fusable_group(g) for g in groups):
        return False
    work = []
    for gi, group in enumerate(groups):
        for p in group["params"]:
            if p.grad is None:
                continue
            g = p.grad
            if not (p.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32 and not g.is_sparse
                    and p.is_contiguous() and g.is_contiguous() and g.device == p.device):
                return False
            work.append((gi, p, g))
    if not work:
        return True
    devices = {p.device for _, p, _ in work}
    if len(devices) != 1:
        return False
    # hyper-parameter rows: one per (group, step count) - the bias corrections depend on the step, and a parameter
    # whose gradient was None for a while lags behind the rest of its group
    rows, row_of = {}, []
    for gi, p, g in work:
        st = opt.state[p]
        if len(st) == 0:   # torch.optim.Adam._init_group: lazily created, same keys / dtypes / layouts
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        m, v = st["exp_avg"], st["exp_avg_sq"]
        if not (m.is_contiguous() and v.is_contiguous() and m.dtype == torch.float32 and v.dtype == torch.float32
                and m.device == p.device and v.device == p.device):
            return False
        row_of.append(rows.setdefault((gi, float(st["step"]) + 1.0), len(rows)))
    if len(rows) > _MAX_GROUPS:
        return False
    hyper = (_lib.AdamGroup * len(rows))()
    for (gi, t), r in rows.items():
        group = groups[gi]
        beta1, beta2 = group["betas"]
        hyper[r].step_size = float(group["lr"]) / (1.0 - beta1 ** t)
        hyper[r].beta1, hyper[r].beta2 = beta1, beta2
        hyper[r].one_minus_beta1, hyper[r].one_minus_beta2 = 1.0 - beta1, 1.0 - beta2   # in double, then rounded
        hyper[r].eps, hyper[r].weight_decay = group["eps"], group["weight_decay"]
        hyper[r].bias_correction2_sqrt = math.sqrt(1.0 - beta2 ** t)
    tensors = (_lib.AdamTensor * len(work))()
    for i, (gi, p, g) in enumerate(work):
        st = opt.state[p]
        st["step"] += 1
        tensors[i].p, tensors[i].g = p.data_ptr(), g.data_ptr()
        tensors[i].m, tensors[i].v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
        tensors[i].n, tensors[i].group = p.numel(), row_of[i]
    n_rows = len(rows)
    dev = next(iter(devices))
    with _lib.on_device(dev):
        _lib.check(_lib.lib().mmbs_adam_step(tensors, len(work), hyper, n_rows, _lib.stream_ptr()),
                   "mmbs_adam_step")
    # the kernel wrote through raw pointers: bump the version counters like an in-place torch op would, so that
    # autograd's saved-tensor checks and the packed-weight caches keyed on `_version` (mlp.py, engine.py,
    # train_engine.py) see the update
    touched = []
    for _, p, _ in work:
        st = opt.state[p]
        touched += [p, st["exp_avg"], st["exp_avg_sq"]]
    _bump_versions(touched)
    return True


def accelerate_optimizer(opt: torch.optim.Optimizer) -> torch.optim.Optimizer:
    """Route ``opt.step()`` of a ``torch.optim.Adam`` through the fused kernel.  Returns the SAME object (param
    groups, state, ``state_dict`` untouched); a no-op for other optimizer classes."""
    if type(opt) is not torch.optim.Adam or getattr(opt, "_mmbs_accelerated", False):
        return opt
    stock = opt.step          # torch's bound (hook-wrapped) step: what runs when the fused path does not apply

    def step(closure=None):
        if closure is None:
            with torch.no_grad():
                if _fused_step(opt):
                    return None
            return stock()
        return stock(closure)

    opt.step = step
    opt._mmbs_accelerated = True
    return opt


_installed = False


def install() -> None:
    """Patch ``torch.optim.Adam.step`` process-wide (what ``MMBS_FUSED_ADAM=1`` does when the drop-in ``models.py``
    is imported): every Adam the unmodified reference scripts construct steps through the fused kernel."""
    global _installed
    if _installed:
        return
    stock = torch.optim.Adam.step

    def step(self, closure=None):
        if closure is None and type(self) is torch.optim.Adam and not getattr(self, "_mmbs_accelerated", False):
            with torch.no_grad():
                if _fused_step(self):
                    return None
        return stock(self, closure)

    step.hooked = True        # Optimizer._patch_step_function must not wrap it a second time
    torch.optim.Adam.step = step
    _installed = True


if os.environ.get("MMBS_FUSED_ADAM", "0") == "1":
    install()
