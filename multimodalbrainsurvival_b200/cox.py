"""Drop-in ``cox_loss`` / ``CoxLoss`` backed by the sm_100a kernels in csrc/cox.cu.

Mirrors the reference's interface (same name, argument meaning, 0-d tensor result,
differentiable w.r.t. ``cox_scores`` only):

    cox_loss(cox_scores, times, status)
        /root/reference/1_HistoPathology/models.py:90-111
        /root/reference/5_JointFusion/models.py:119-140
        /root/reference/2_GeneExpression/models.py:24-45
        /root/reference/3_EarlyFusion/models.py:24-45
    CoxLoss().forward(cox_scores, times, status)
        /root/reference/1_HistoPathology/models.py:113-118, 5_JointFusion/models.py:142-147

Differences, all deliberate:
  * tie order is the stable one (``torch.sort(-times, stable=True)``);
  * a NaN term does not drop into ``pdb`` (models.py:107-109); the device-side flag
    is checked only when ``MMBS_COX_CHECK_NAN=1`` (it costs a host sync);
  * tensors must live on a CUDA device: there is no CPU fallback.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib


SMALL_N = 2048  # csrc/cox.cu SM_MAX: one fused block, no workspace


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _prep(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"cox_loss: `{name}` must be a CUDA tensor (no CPU fallback in this build)")
    t = t.detach().reshape(-1)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def risk_order(times: torch.Tensor) -> torch.Tensor:
    """Stable argsort(-times) as int32 (the risk-set order), via the radix sort."""
    t = _prep(times, "times")
    n = t.numel()
    perm = torch.empty(n, dtype=torch.int32, device=t.device)
    if n == 0:
        return perm
    L = _lib.lib()
    with torch.cuda.device(t.device):
        nbytes = L.mmbs_cox_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
        _lib.check(L.mmbs_risk_order(_lib.ptr(t), n, _lib.ptr(perm), _lib.ptr(ws), nbytes, _lib.stream_ptr()),
                   "mmbs_risk_order")
    return perm


def pipeline_state(times: torch.Tensor) -> dict:
    """Diagnostics: sort `times` like cox_loss would and report whether the bucketed pipeline kept the input
    (``state`` 0) or handed it to the LSD-sort pipeline (bit 0: bucket overflow, bit 1: oversized sub-bucket;
    -1: n outside the bucketed range).  Synchronises the device; for tests and tools/cox_fallback_scan.py."""
    import ctypes
    t = _prep(times, "times")
    n = t.numel()
    perm = torch.empty(n, dtype=torch.int32, device=t.device)
    L = _lib.lib()
    out = (ctypes.c_int32 * 4)()
    with torch.cuda.device(t.device):
        nbytes = L.mmbs_cox_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
        _lib.check(L.mmbs_risk_order(_lib.ptr(t), n, _lib.ptr(perm), _lib.ptr(ws), nbytes, _lib.stream_ptr()),
                   "mmbs_risk_order")
        _lib.check(L.mmbs_cox_debug_state(_lib.ptr(ws), nbytes, n, ctypes.cast(out, ctypes.c_void_p)),
                   "mmbs_cox_debug_state")
    return {"state": int(out[0]), "largest_bucket": int(out[1]), "buckets": int(out[2]), "capacity": int(out[3])}


_WS_BYTES = {}   # n -> mmbs_cox_workspace_bytes(n): a pure function of n, one ctypes round trip saved per call


def _ws_bytes(L, n: int) -> int:
    v = _WS_BYTES.get(n)
    if v is None:
        if len(_WS_BYTES) > 256:
            _WS_BYTES.clear()
        v = _WS_BYTES[n] = int(L.mmbs_cox_workspace_bytes(n))
    return v


def _flat_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """_prep without tensor operations for the common case (a contiguous 1-D fp32 CUDA tensor)."""
    if t.is_cuda and t.dtype == torch.float32 and t.dim() == 1 and t.is_contiguous():
        return t.detach() if t.requires_grad else t
    return _prep(t, name)


class _CoxLossFn(torch.autograd.Function):
    # The host side of a call is on the critical path when the caller synchronises between steps (the GPU idles until
    # the first launch): no tensor views, the output pointers are computed from ONE allocation's base address.
    @staticmethod
    def forward(ctx, cox_scores, times, status):
        s = _flat_f32(cox_scores, "cox_scores")
        t = _flat_f32(times, "times")
        d = _flat_f32(status, "status")
        n = s.numel()
        if t.numel() != n or d.numel() != n:
            raise ValueError(f"cox_loss: size mismatch scores={tuple(cox_scores.shape)} "
                             f"times={tuple(times.shape)} status={tuple(status.shape)}")
        dev = s.device
        if n == 0:  # mean over an empty batch
            ctx.n = 0
            ctx.in_shape = cox_scores.shape
            ctx.in_dtype = cox_scores.dtype
            return torch.full((), float("nan"), device=dev)
        L = _lib.lib()
        npad = _pad64(n)
        switch = torch.cuda.current_device() != dev.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(dev)
        try:
            # one allocation: [perm | saved s~ | saved w | loss, flags]  (all 4-byte words)
            buf = torch.empty(3 * npad + 64, dtype=torch.float32, device=dev)
            base = buf.data_ptr()
            if n <= SMALL_N:
                ws, nbytes, ws_ptr = None, 0, 0
            else:
                nbytes = _ws_bytes(L, n)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                ws_ptr = ws.data_ptr()
            out_ptr = base + 12 * npad
            _lib.check(L.mmbs_cox_forward(s.data_ptr(), t.data_ptr(), d.data_ptr(), n, base, base + 4 * npad,
                                          base + 8 * npad, out_ptr, out_ptr + 4, ws_ptr, nbytes,
                                          _lib.stream_ptr()), "mmbs_cox_forward")
        finally:
            if switch:
                torch.cuda.set_device(prev)
        if os.environ.get("MMBS_COX_CHECK_NAN", "0") == "1" and int(buf[3 * npad + 1:3 * npad + 2].view(torch.int32).item()) != 0:
            raise FloatingPointError(f"cox_loss: NaN in the loss terms (n={n})")
        ctx.n = n
        ctx.in_shape = cox_scores.shape
        ctx.in_dtype = cox_scores.dtype
        ctx.ws_bytes = nbytes
        ctx.save_for_backward(s, d, buf, *([ws] if ws is not None else []))
        ctx.npad = npad
        return buf[3 * npad].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        if ctx.n == 0:
            return torch.zeros(ctx.in_shape, dtype=ctx.in_dtype, device=grad_loss.device), None, None
        s, d, buf = ctx.saved_tensors[:3]
        ws = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
        n, npad = ctx.n, ctx.npad
        g = grad_loss
        if not (g.dtype == torch.float32 and g.is_contiguous() and g.device == s.device):
            g = g.detach().reshape(1).to(s.device).float().contiguous()
        grad = torch.empty(n, dtype=torch.float32, device=s.device)
        L = _lib.lib()
        base = buf.data_ptr()
        switch = torch.cuda.current_device() != s.device.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(s.device)
        try:
            _lib.check(L.mmbs_cox_backward(s.data_ptr(), d.data_ptr(), base, base + 4 * npad, base + 8 * npad,
                                           g.data_ptr(), n, grad.data_ptr(), ws.data_ptr() if ws is not None else 0,
                                           ctx.ws_bytes, _lib.stream_ptr()), "mmbs_cox_backward")
        finally:
            if switch:
                torch.cuda.set_device(prev)
        if grad.shape != ctx.in_shape:
            grad = grad.reshape(ctx.in_shape)
        return (grad if ctx.in_dtype == torch.float32 else grad.to(ctx.in_dtype)), None, None


def cox_loss(cox_scores, times, status):
    """
    :param cox_scores: cox scores, size (batch_size)
    :param times: event times (either death or censor), size batch_size
    :param status: event status (1 for death, 0 for censor), size batch_size
    :return: 0-d loss tensor: mean over the batch of the negative log partial likelihood terms
    """
    return _CoxLossFn.apply(cox_scores, times, status)


class CoxLoss(nn.Module):
    def __init__(self):
        super(CoxLoss, self).__init__()

    def forward(self, cox_scores, times, status):
        return cox_loss(cox_scores, times, status)


def cox_loss_with_order(cox_scores, times, status):
    """(loss, perm int32) - the loss plus the risk-set order the kernel used (parity tests)."""
    s = _prep(cox_scores, "cox_scores")
    t = _prep(times, "times")
    d = _prep(status, "status")
    n = s.numel()
    L = _lib.lib()
    with torch.cuda.device(s.device):
        nbytes = L.mmbs_cox_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=s.device)
        perm = torch.empty(n, dtype=torch.int32, device=s.device)
        saved = torch.empty(2, _pad64(n), dtype=torch.float32, device=s.device)
        out = torch.empty(2, dtype=torch.float32, device=s.device)
        _lib.check(L.mmbs_cox_forward(_lib.ptr(s), _lib.ptr(t), _lib.ptr(d), n, _lib.ptr(perm),
                                      _lib.ptr(saved[0]), _lib.ptr(saved[1]), _lib.ptr(out),
                                      _lib.ptr(out[1:]), _lib.ptr(ws), nbytes, _lib.stream_ptr()),
                   "mmbs_cox_forward")
    return out[0].clone(), perm & 0x7FFFFFFF  # bit 31 carries the event indicator
