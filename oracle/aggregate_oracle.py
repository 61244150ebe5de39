"""Oracle for per-patient aggregation.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy, the two aggregation tails of the reference:

* ``extract_features`` tail - per-case mean of patch feature rows
  (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:75-89 and
  2_GeneExpression/3_GeneExpress_extractfeatures.py:70-82): for every case in
  ``set(case_list)``: boolean mask . features / mask.sum().
  The reference iterates ``set(case_list)`` (hash order); the *grouping* is what
  must be bit-exact, so this oracle takes the case order as an argument.

* ``get_survival_CI`` grouping
  (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:126-152 and its 7
  copies): ``ids_unique = sorted(set(ids))``; score[id] =
  np.mean of that id's float32 outputs; survival/vital = last row seen.

Pinned by tests/golden/aggregate_*.npz (tools/make_golden.py runs the
reference's own function bodies).
"""
from __future__ import annotations

import numpy as np


def group_ids(ids, order="sorted"):
    """Map a list of hashable ids to (unique_ids, int32 segment index per row).

    order="sorted"      -> sorted(set(ids))              (savescore.py:128)
    order="first_seen"  -> order of first appearance     (deterministic stand-in
                           for the reference's hash-ordered ``set`` iteration)
    """
    if order == "sorted":
        uniq = sorted(set(ids))
    elif order == "first_seen":
        uniq = list(dict.fromkeys(ids))
    else:
        uniq = list(order)
    lut = {k: i for i, k in enumerate(uniq)}
    seg = np.fromiter((lut[i] for i in ids), dtype=np.int32, count=len(ids))
    return uniq, seg


def case_mean_features(features: np.ndarray, case_list, case_order):
    """extractfeatures.py:80-88 with an explicit case order.  float64 result
    (numpy>=2 promotes ``float32 / int64``), accumulated like the reference."""
    out = []
    for case in case_order:
        l = np.array([(x == case) for x in case_list])
        out.append(l.T.dot(features) / (l.sum()))
    return np.asarray(out)


def segment_mean(values: np.ndarray, seg: np.ndarray, n_seg: int):
    """fp64-accumulated segment mean + counts (the exact-arithmetic checker)."""
    values = np.asarray(values)
    v2 = values.reshape(values.shape[0], -1).astype(np.float64)
    acc = np.zeros((n_seg, v2.shape[1]), dtype=np.float64)
    np.add.at(acc, seg, v2)
    cnt = np.bincount(seg, minlength=n_seg).astype(np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = acc / cnt[:, None]
    return mean.reshape((n_seg,) + values.shape[1:]), cnt


def survival_grouping(output_list, ids_list, survival_months, vital_status):
    """savescore.py:128-145 minus the third-party c-index call.  Returns
    (ids_unique, score float32[], survival[], vital[])."""
    ids_unique = sorted(list(set(ids_list)))
    id_to_scores, id_to_sm, id_to_vs = {}, {}, {}
    for i in range(len(output_list)):
        k = ids_list[i]
        id_to_scores.setdefault(k, []).append(output_list[i, 0])
        id_to_sm[k] = survival_months[i]
        id_to_vs[k] = vital_status[i]
    score = np.array([np.mean(id_to_scores[k]) for k in ids_unique])
    sm = np.array([id_to_sm[k] for k in ids_unique])
    vs = np.array([id_to_vs[k] for k in ids_unique])
    return ids_unique, score, sm, vs
