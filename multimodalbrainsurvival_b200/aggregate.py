"""Per-patient aggregation on the GPU (csrc/segmean.cu), mirroring the reference's tails.

* ``aggregate_case_features``  <->  the per-case mean at the end of ``extract_features``
  (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:75-89,
   /root/reference/2_GeneExpression/3_GeneExpress_extractfeatures.py:70-82)
* ``survival_grouping`` / ``get_survival_CI``  <->  ``get_survival_CI``
  (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:126-152 and its 7 copies)

The id -> segment mapping (string handling) is host logic and bit-exact with the
reference (``set`` / ``sorted(set)``); the means are computed by the segmented-mean
kernel.  No CPU fallback: the value tensors must be CUDA tensors (or are moved to
the current CUDA device when given as numpy arrays).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def segmented_mean(values: torch.Tensor, seg_ids: torch.Tensor, n_seg: int):
    """values [n, ...] fp32 CUDA, seg_ids [n] int32 CUDA in [0, n_seg) ->
    (mean [n_seg, ...] fp32, counts [n_seg] int32, last_row [n_seg] int32).
    An id outside [0, n_seg) does not raise (that would cost a host synchronisation per call): the kernel poisons
    the whole result instead - NaN means, counts == -1, last_row == -1."""
    if not values.is_cuda or not seg_ids.is_cuda:
        raise RuntimeError("segmented_mean: CUDA tensors required (no CPU fallback in this build)")
    n = values.shape[0]
    if seg_ids.numel() != n:
        raise ValueError(f"segmented_mean: {n} rows but {seg_ids.numel()} segment ids")
    v = values.detach().reshape(n, -1).float().contiguous()
    s = seg_ids.detach().reshape(-1).to(torch.int32).contiguous()
    d = v.shape[1]
    dev = v.device
    out = torch.empty((n_seg, d), dtype=torch.float32, device=dev)
    counts = torch.empty(n_seg, dtype=torch.int32, device=dev)
    last = torch.empty(n_seg, dtype=torch.int32, device=dev)
    if n == 0 or n_seg == 0:
        return out.fill_(float("nan")).reshape((n_seg,) + tuple(values.shape[1:])), counts.zero_(), last.fill_(-1)
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L.mmbs_segmented_mean_workspace_bytes(n, n_seg)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(L.mmbs_segmented_mean(_lib.ptr(v), _lib.ptr(s), n, d, n_seg, _lib.ptr(out), _lib.ptr(counts),
                                         _lib.ptr(last), _lib.ptr(ws), nbytes, _lib.stream_ptr()),
                   "mmbs_segmented_mean")
    return out.reshape((n_seg,) + tuple(values.shape[1:])), counts, last


def _segments(ids, uniq):
    lut = {k: i for i, k in enumerate(uniq)}
    return np.fromiter((lut[i] for i in ids), dtype=np.int32, count=len(ids))


def _to_cuda(x, device=None):
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.cuda(device) if device is not None else x.cuda()
        return x
    return torch.as_tensor(np.asarray(x)).cuda(device) if device is not None else torch.as_tensor(np.asarray(x)).cuda()


def aggregate_case_features(features, case_list, case_order=None):
    """Per-case mean of feature rows.

    features: (N, D) CUDA tensor or numpy array; case_list: N hashable ids.
    Returns (case_uniques: list, features_final: np.ndarray float64 (n_cases, D)) exactly
    like the tail of the reference's ``extract_features``; the case order is the
    reference's ``set(case_list)`` iteration order unless ``case_order`` is given.
    """
    case_list = list(case_list)
    case_uniques = list(set(case_list)) if case_order is None else list(case_order)
    feats = _to_cuda(features)
    seg = torch.from_numpy(_segments(case_list, case_uniques)).to(feats.device)
    mean, _, _ = segmented_mean(feats, seg, len(case_uniques))
    return case_uniques, mean.double().cpu().numpy()


def survival_grouping(output_list, ids_list, survival_months, vital_status):
    """The grouping part of ``get_survival_CI``: (ids_unique, score float32[n_ids],
    survival_months[n_ids], vital_status[n_ids]) with the "last row seen" rule."""
    ids_list = list(ids_list)
    ids_unique = sorted(list(set(ids_list)))
    out = _to_cuda(output_list)
    seg = torch.from_numpy(_segments(ids_list, ids_unique)).to(out.device)
    mean, _, last = segmented_mean(out.reshape(len(ids_list), -1)[:, :1], seg, len(ids_unique))
    last = last.cpu().numpy().astype(np.int64)
    sm = np.asarray(survival_months.cpu() if isinstance(survival_months, torch.Tensor) else survival_months)[last]
    vs = np.asarray(vital_status.cpu() if isinstance(vital_status, torch.Tensor) else vital_status)[last]
    return ids_unique, mean[:, 0].cpu().numpy(), sm, vs


PAIRWISE_MAX_N = 50_000     # above: the O(n S) dominance count instead of the O(n^2) pair kernel
_DOMINANCE_SHIFT = 11       # S = 2048 deaths per table block


def _concordance_counts_dominance(t, p, e):
    """The same exact counts for large cohorts (csrc/cindex.cu, cindex_dominance_kernel).  torch does the preparatory
    sorts / binary searches / prefix sums of the index arrays; the pair counting itself is the kernel's."""
    n = t.numel()
    dev = t.device
    d = e != 0
    nd = int(d.sum())
    if nd == 0:
        return 0, 0, 0
    td, order = torch.sort(t[d], stable=True)                     # deaths by exit time: time position k
    pd_ = p[d][order]
    # admissible deaths of subject i = a prefix of that order: strictly earlier deaths, plus - for a censored subject -
    # the deaths at its own exit time
    adm = torch.where(d, torch.searchsorted(td, t, right=False), torch.searchsorted(td, t, right=True))
    pr, perm = torch.sort(pd_, stable=True)                       # deaths by prediction: perm[s] = time position
    lo = torch.searchsorted(pr, p, right=False)
    hi = torch.searchsorted(pr, p, right=True)
    perm32 = perm.to(torch.int32).contiguous()
    inv32 = torch.empty_like(perm32)
    inv32[perm] = torch.arange(nd, device=dev, dtype=torch.int32)
    S = 1 << _DOMINANCE_SHIFT
    nb = (nd + S - 1) // S
    cell = (torch.arange(nd, device=dev) >> _DOMINANCE_SHIFT) * nb + (perm >> _DOMINANCE_SHIFT)
    hist = torch.bincount(cell, minlength=nb * nb).view(nb, nb)
    table = torch.zeros((nb + 1, nb + 1), dtype=torch.int64, device=dev)
    table[1:, 1:] = hist.cumsum(0).cumsum(1)                      # exclusive 2-D prefix: #{s < bs S, perm[s] < bv S}
    out = torch.zeros(3, dtype=torch.int64, device=dev)
    out[0] = adm.sum()
    with _lib.on_device(dev):
        _lib.check(_lib.lib().mmbs_concordance_dominance(_lib.ptr(perm32), _lib.ptr(inv32), _lib.ptr(table), nd,
                                                         _DOMINANCE_SHIFT, _lib.ptr(lo.contiguous()), _lib.ptr(hi.contiguous()),
                                                         _lib.ptr(adm.contiguous()), n, _lib.ptr(out), _lib.stream_ptr()),
                   "mmbs_concordance_dominance")
    pairs, correct, tied = (int(v) for v in out.cpu().tolist())
    return pairs, correct, tied


def concordance_counts(event_times, predicted_scores, event_observed, device=None):
    """(admissible pairs, correct, tied) of Harrell's C, counted exactly on the GPU (csrc/cindex.cu).
    Inputs: array-likes / tensors of length n; compared in float64 like the reference's pandas columns."""
    def f64(x):
        if isinstance(x, torch.Tensor):
            return x.detach().reshape(-1).to(torch.float64)
        return torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(-1))
    t, p = f64(event_times), f64(predicted_scores)
    e = (event_observed.detach().reshape(-1) != 0) if isinstance(event_observed, torch.Tensor) \
        else torch.as_tensor(np.asarray(event_observed).reshape(-1) != 0)
    n = t.numel()
    if p.numel() != n or e.numel() != n:
        raise ValueError(f"concordance_index: lengths differ ({n}, {p.numel()}, {e.numel()})")
    if n == 0:
        return 0, 0, 0
    dev = next((x.device for x in (t, p) if x.is_cuda), None) or torch.device("cuda", torch.cuda.current_device()
                                                                             if device is None else device)
    t, p, e = t.to(dev).contiguous(), p.to(dev).contiguous(), e.to(dev).to(torch.uint8).contiguous()
    if n > PAIRWISE_MAX_N:
        return _concordance_counts_dominance(t, p, e)
    out = torch.empty(3, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mmbs_concordance_counts(_lib.ptr(t), _lib.ptr(p), _lib.ptr(e), n, _lib.ptr(out),
                                                      _lib.stream_ptr()), "mmbs_concordance_counts")
    pairs, correct, tied = (int(v) for v in out.cpu().tolist())
    return pairs, correct, tied


def concordance_index(event_times, predicted_scores, event_observed=None):
    """Drop-in for ``lifelines.utils.concordance_index(event_times, predicted_scores, event_observed)`` as the
    reference calls it (3_HistoPath_savescore.py:147).  Raises ZeroDivisionError without admissible pairs, like
    lifelines."""
    if event_observed is None:
        event_observed = np.ones(len(event_times))
    pairs, correct, tied = concordance_counts(event_times, predicted_scores, event_observed)
    if pairs == 0:
        raise ZeroDivisionError("No admissable pairs in the dataset.")
    return (correct + tied / 2) / pairs


def get_survival_CI(output_list, ids_list, survival_months, vital_status, concordance_index=None):
    """Same contract as the reference's ``get_survival_CI``: returns (CI, DataFrame with
    columns id, score, survival_months, vital_status).  ``concordance_index`` defaults to the GPU pair count
    above (the reference's third-party ``lifelines.utils.concordance_index`` can be passed in instead)."""
    import pandas as pd
    if concordance_index is None:
        concordance_index = globals()["concordance_index"]
    ids_unique, score_list, sm, vs = survival_grouping(output_list, ids_list, survival_months, vital_status)
    CI = concordance_index(sm, -score_list, vs)
    pandas_output = pd.DataFrame({'id': ids_unique, 'score': score_list, 'survival_months': sm,
                                  'vital_status': vs})
    return CI, pandas_output


def save_features_csv(path, features, threads: int = 0) -> None:
    """Drop-in for ``np.savetxt(path, features, delimiter=",")`` as the extraction scripts call it
    (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:184-192,
    /root/reference/2_GeneExpression/3_GeneExpress_extractfeatures.py:143-149): the same bytes ('%.18e' per value),
    written by the multi-threaded formatter of csrc/csv.cu (host code)."""
    import ctypes
    a = np.ascontiguousarray(np.asarray(features, dtype=np.float64))
    if a.ndim == 1:               # np.savetxt writes a 1-D array as one value per row
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError(f"save_features_csv: expected a 1-D or 2-D array, got shape {a.shape}")
    if a.shape[1] == 0:
        raise ValueError("save_features_csv: no columns")
    _lib.check(_lib.lib().mmbs_write_matrix_csv(a.ctypes.data_as(ctypes.c_void_p), a.shape[0], a.shape[1],
                                                str(path).encode(), int(threads)), "mmbs_write_matrix_csv")
