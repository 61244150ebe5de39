"""Oracle for the Cox negative log partial likelihood.  TEST INFRASTRUCTURE ONLY.

Restates ``cox_loss`` of the reference
(/root/reference/1_HistoPathology/models.py:90-111, textually identical copies at
5_JointFusion/models.py:119-140, 2_GeneExpression/models.py:24-45,
3_EarlyFusion/models.py:24-45) in numpy:

    pi      = stable argsort(-times)                       models.py:99
    s~      = scores[pi] - max(scores)                     models.py:100,102
    C_i     = sum_{j<=i} exp(s~_j)                         models.py:103-104
    L       = -(1/N) sum_i status[pi]_i (s~_i - log(C_i + 1e-5))   :104-111

Tie rule: the reference calls ``torch.sort(-times)`` (unstable by default); the
only well-defined order is the stable one (ties keep ascending original index),
which is what ``torch.sort(..., stable=True)`` and this oracle produce
(SURVEY.md §7 hard part 4).

Pinned by tests/test_oracle_cox.py against the SURVEY §4 known-answer vector and
tests/golden/cox_*.npz (generated from the reference by tools/make_golden.py).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-5


def order_key_u32(times: np.ndarray) -> np.ndarray:
    """Order-preserving u32 image of -times (the radix-sort key).

    ascending key order == ascending (-t) == descending t.  -0.0 and +0.0 are
    canonicalised to the same key because torch.sort compares them equal.
    NaN times are not supported by the reference either (sort order undefined).
    """
    neg = (-np.asarray(times, dtype=np.float32)).astype(np.float32)
    neg = neg + np.float32(0.0)  # -0.0 + 0.0 -> +0.0
    bits = neg.view(np.uint32)
    sign = (bits >> np.uint32(31)).astype(bool)
    return np.where(sign, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)


def risk_order(times: np.ndarray) -> np.ndarray:
    """pi = stable argsort(-times); int64 (models.py:99)."""
    neg = -np.asarray(times, dtype=np.float32)
    return np.argsort(neg, kind="stable").astype(np.int64)


def cox_forward(scores, times, status, dtype=np.float64):
    """Returns (loss, perm).  dtype=float64 is the exact-arithmetic oracle,
    dtype=float32 mimics the reference's working precision (fp64 cumsum carry,
    as torch's CPU cumsum does)."""
    s = np.asarray(scores, dtype=np.float32).astype(dtype)
    d = np.asarray(status, dtype=np.float32).astype(dtype)
    n = s.shape[0]
    perm = risk_order(times)
    if n == 0:
        return dtype(np.nan), perm
    st = s[perm] - s.max()
    e = np.exp(st)
    c = np.cumsum(e.astype(np.float64)).astype(dtype)
    terms = -(st - np.log(c + dtype(EPS))) * d[perm]
    return dtype(terms.astype(np.float64).sum() / n), perm


def cox_backward(scores, times, status, grad_loss=1.0):
    """d loss / d scores evaluated in fp64, returned as fp64.

    w_i = delta_i / (C_i + eps); W_k = sum_{i>=k} w_i;
    g~_k = -(delta_k - exp(s~_k) W_k) / N; the gradient through ``- max(s)``
    subtracts sum_k g~_k, split evenly over all positions attaining the max
    (torch's full-reduction max backward); un-permute.
    """
    s = np.asarray(scores, dtype=np.float32).astype(np.float64)
    d = np.asarray(status, dtype=np.float32).astype(np.float64)
    n = s.shape[0]
    perm = risk_order(times)
    smax = s.max()
    st = s[perm] - smax
    e = np.exp(st)
    c = np.cumsum(e)
    w = d[perm] / (c + EPS)
    W = np.cumsum(w[::-1])[::-1]
    gt = -(d[perm] - e * W) / n
    g = np.zeros(n, dtype=np.float64)
    g[perm] = gt
    is_max = s == smax
    g[is_max] -= gt.sum() / is_max.sum()
    return g * float(grad_loss)


def cox_loss_and_grad(scores, times, status):
    loss, perm = cox_forward(scores, times, status, np.float64)
    return float(loss), cox_backward(scores, times, status), perm
