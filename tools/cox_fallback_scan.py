#!/usr/bin/env python
"""How often does the bucketed Cox pipeline give an input up (device-side fallback to the LSD sort)?  A fallback shows as a
call that takes > 2x the median.   python tools/cox_fallback_scan.py [n] [seeds]"""
import os
import statistics
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import cox

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = "cuda"


def timed(s, t, e):
    out = []
    for _ in range(3):
        s.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cox.cox_loss(s, t, e).backward()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


for dist in ("uniform200", "exponential", "default_rng_unseeded"):
    ts = []
    for seed in range(seeds):
        g = torch.Generator(device=dev).manual_seed(seed)
        if dist == "default_rng_unseeded":
            s = torch.randn(n, device=dev).requires_grad_(True)
            t = torch.rand(n, device=dev) * 200
            e = (torch.rand(n, device=dev) < 0.6).float()
        else:
            s = torch.randn(n, device=dev, generator=g).requires_grad_(True)
            u = torch.rand(n, device=dev, generator=g)
            t = u * 200 if dist == "uniform200" else -torch.log1p(-u) * 30
            e = (torch.rand(n, device=dev, generator=g) < 0.6).float()
        ts.append(timed(s, t, e))
        st = cox.pipeline_state(t)
        if st["state"] != 0:
            print(f"  seed {seed} {dist}: {st}")
    med = statistics.median(ts)
    slow = [(i, round(x, 3)) for i, x in enumerate(ts) if x > 2 * med]
    print(f"n={n} {dist}: median {med:.3f} ms, min {min(ts):.3f}, max {max(ts):.3f}, fallbacks (seed, ms): {slow}")
