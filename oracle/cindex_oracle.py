"""Harrell's concordance index.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference calls ``lifelines.utils.concordance_index(survival_months,
-score, vital_status)`` (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:147
and 7 copies).  lifelines is a third-party dependency that the
reference neither vendors nor pins (no requirements/lock file) and it is absent
from this image, so this is a restatement of its published algorithm
(``lifelines.utils.concordance._concordance_summary_statistics``), not a checked
port, and no value produced by lifelines itself pins it:

  subjects are visited in order of exit time; at each time the deaths are handled
  (compared with the deaths that exited STRICTLY earlier, then added to the pool),
  then the censored subjects (compared with every death in the pool, i.e. including
  the deaths at their own exit time).  For a visited subject i and a pooled death j:
      pairs += 1;  correct += pred_j < pred_i;  tied += pred_j == pred_i
  C = (correct + tied / 2) / pairs.

So a pair is admissible iff the earlier exit is an observed death; two deaths at the
same time are not comparable; a death and a censoring at the same time are (the death
counts as earlier).  ``concordance_counts`` returns the three integers, which the CUDA
kernel (csrc/cindex.cu) must reproduce exactly.

Used to compare reference-path scores with new-path scores under the SAME
implementation (|dC| <= 0.005, BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np


def concordance_counts(event_times, predicted, event_observed):
    """(pairs, correct, tied) as Python ints.  O(n * deaths), vectorised over the deaths."""
    t = np.asarray(event_times, dtype=np.float64)
    p = np.asarray(predicted, dtype=np.float64)
    e = np.asarray(event_observed).astype(bool)
    td, pd_ = t[e], p[e]
    pairs = correct = tied = 0
    for i in range(t.shape[0]):
        adm = (td < t[i]) | ((td == t[i]) if not e[i] else False)
        pairs += int(adm.sum())
        correct += int((pd_[adm] < p[i]).sum())
        tied += int((pd_[adm] == p[i]).sum())
    return pairs, correct, tied


def concordance_index(event_times, predicted, event_observed) -> float:
    pairs, correct, tied = concordance_counts(event_times, predicted, event_observed)
    return (correct + 0.5 * tied) / pairs if pairs > 0 else float("nan")
