"""Harrell's concordance index.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference calls ``lifelines.utils.concordance_index(survival_months,
-score, vital_status)`` (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:147
and 7 copies).  lifelines is a third-party dependency that the
reference neither vendors nor pins (no requirements/lock file) and it is absent
from this image, so this is a restatement of the published definition, not a
checked port: a pair (i, j) is admissible iff the smaller time is an observed
event (equal times: admissible only when exactly one of the two is an event...
lifelines drops equal-time pairs unless one is censored; we follow the plain
Harrell rule and drop equal-time pairs); concordant when the larger predicted
value belongs to the longer survivor; prediction ties count 1/2.

Used only to compare reference-path scores with new-path scores under the SAME
implementation (|dC| <= 0.005, BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np


def concordance_index(event_times, predicted, event_observed) -> float:
    t = np.asarray(event_times, dtype=np.float64)
    p = np.asarray(predicted, dtype=np.float64)
    e = np.asarray(event_observed).astype(bool)
    n = t.shape[0]
    num = 0.0
    den = 0.0
    # O(n^2) blocked; n is a few thousand cases at most in tests.
    for i in range(n):
        if not e[i]:
            continue
        later = t > t[i]
        den += later.sum()
        num += (p[later] > p[i]).sum() + 0.5 * (p[later] == p[i]).sum()
    return float(num / den) if den > 0 else float("nan")
