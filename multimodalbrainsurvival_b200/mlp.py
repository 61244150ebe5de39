"""Fused execution of the reference's inline MLPs on the tcgen05 GEMM kernel.

The reference builds its MLPs from stock ``nn.Linear`` inside the scripts
(2_GeneExpression/1_GeneExpress_train.py:247-257, 3_EarlyFusion/2_EarlyFusion_train.py:242-253,
5_JointFusion/1_JointFusion_train.py:314-325), so this module only ever *receives* an
``nn.Sequential``; it pattern-matches ``Dropout / Linear / ReLU`` stacks and runs them as
GEMMs with the bias and ReLU fused into the epilogue, reading the ``nn.Linear`` parameters in
place (bf16 padded copies are caches keyed on the parameters' version counters).

Served by the kernels: inference (eval mode or autograd disabled) on CUDA tensors.
Training mode (dropout masks + autograd) runs the module graph on the tensor's device.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import engine


def _pad(n, m):
    return (n + m - 1) // m * m


def _parse(seq):
    """-> list of (linear, relu_after) or None when the stack is not a plain MLP."""
    if not isinstance(seq, nn.Sequential):
        return None
    layers = []
    for m in seq:
        if isinstance(m, nn.Dropout):
            continue
        if isinstance(m, nn.Linear):
            layers.append([m, False])
        elif isinstance(m, nn.ReLU) and layers and not layers[-1][1]:
            layers[-1][1] = True
        else:
            return None
    return layers or None


class _MLPEngine:
    def __init__(self, layers, m_rows, device):
        self.layers = layers
        self.m = m_rows
        self.device = device
        self.k0 = layers[0][0].in_features
        with torch.cuda.device(device):
            self.x_bf = torch.empty((m_rows, _pad(self.k0, 64)), dtype=torch.bfloat16, device=device)
            self.w_bf, self.b_pad, self.outs, self.plans = [], [], [], []
            cur = self.x_bf
            for i, (lin, relu) in enumerate(layers):
                last = i == len(layers) - 1
                kp = cur.shape[1]
                npad = _pad(lin.out_features, 32 if last else 64)
                w = torch.zeros((npad, kp), dtype=torch.bfloat16, device=device)
                b = torch.zeros(npad, dtype=torch.float32, device=device)
                out = torch.empty((m_rows, npad), dtype=torch.float32 if last else torch.bfloat16, device=device)
                self.w_bf.append(w)
                self.b_pad.append(b)
                self.outs.append(out)
                self.plans.append(engine.linear_plan(cur, w, b, out, relu=relu))
                cur = out
        self.version = None

    def _weights_version(self):
        v = []
        for lin, _ in self.layers:
            v.append((lin.weight.data_ptr(), lin.weight._version))
            if lin.bias is not None:
                v.append((lin.bias.data_ptr(), lin.bias._version))
        return tuple(v)

    def refresh(self):
        ver = self._weights_version()
        if ver == self.version:
            return
        for (lin, _), w, b in zip(self.layers, self.w_bf, self.b_pad):
            n = lin.out_features
            engine.cast_pad_bf16(lin.weight, w.shape[1], out=w[:n])
            if lin.bias is not None:
                b[:n].copy_(lin.bias.detach())
        self.version = ver

    def run(self, x):
        self.refresh()
        engine.cast_pad_bf16(x, self.x_bf.shape[1], out=self.x_bf)
        for p in self.plans:
            p.run()
        return self.outs[-1][:, :self.layers[-1][0].out_features].clone()


_ENGINES = {}


def _fusable(seq, x):
    if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1":
        return None
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2):
        return None
    if torch.is_grad_enabled() and (seq.training or x.requires_grad or any(p.requires_grad for p in seq.parameters())):
        return None
    layers = _parse(seq)
    if layers is None or layers[0][0].in_features != x.shape[1]:
        return None
    return layers


def run_mlp(seq, x):
    """Evaluate ``seq(x)``: fused kernels when the stack is a plain MLP in inference mode."""
    layers = _fusable(seq, x)
    if layers is None:
        return seq(x)
    key = (id(seq), x.shape[0], x.device.index)
    eng = _ENGINES.get(key)
    if eng is None or any(a is not b[0] for a, b in zip([l[0] for l in eng.layers], layers)):
        eng = _MLPEngine(layers, x.shape[0], x.device)
        if len(_ENGINES) > 64:
            _ENGINES.clear()
        _ENGINES[key] = eng
    return eng.run(x)


class AcceleratedSequential(nn.Sequential):
    """``nn.Sequential`` whose forward goes through :func:`run_mlp` (same children/keys)."""

    @classmethod
    def wrap(cls, seq):
        if isinstance(seq, cls):
            return seq
        if not isinstance(seq, nn.Sequential):
            raise TypeError("accelerate() expects an nn.Sequential MLP")
        new = cls()
        for name, child in seq.named_children():
            new.add_module(name, child)
        new.train(seq.training)
        return new

    def forward(self, x):
        layers = _fusable(self, x)
        if layers is None:
            return nn.Sequential.forward(self, x)
        return run_mlp(self, x)
