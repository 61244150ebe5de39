#!/usr/bin/env python
"""Concordance index on large cohorts: O(n^2) pair kernel vs the O(n S) dominance count.  python tools/cindex_timing.py"""
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import aggregate

g = torch.Generator(device="cuda").manual_seed(0)
for n in (50_000, 200_000, 1_000_000, 10_000_000):
    t = torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 200
    p = torch.randn(n, device="cuda", generator=g, dtype=torch.float64)
    e = (torch.rand(n, device="cuda", generator=g) < 0.6)
    for name, limit in (("dominance", 1000), ("pairs", 1 << 40)):
        if name == "pairs" and n > 200_000:
            continue
        aggregate.PAIRWISE_MAX_N = limit
        aggregate.concordance_counts(t, p, e)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = aggregate.concordance_counts(t, p, e)
        torch.cuda.synchronize()
        print(f"n={n} {name}: {1e3 * (time.perf_counter() - t0):.2f} ms  C = {(c[1] + c[2] / 2) / c[0]:.6f}")
