#!/usr/bin/env python
"""Cox fwd+bwd on a 10 M-sample risk set (BASELINE config 4): CUDA-event time per call, and - when run under
   ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none  - the per-kernel launch list.
   python tools/cox_profile.py [n] [reps]"""
import os
import statistics
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import cox

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1111)
s = torch.randn(n, device=dev, generator=g).requires_grad_(True)
t = torch.rand(n, device=dev, generator=g) * 200
e = (torch.rand(n, device=dev, generator=g) < 0.6).float()
fw, tot = [], []
for i in range(reps):
    s.grad = None
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    loss = cox.cox_loss(s, t, e)
    b.record()
    loss.backward()
    c.record()
    torch.cuda.synchronize()
    if i >= 2:
        fw.append(a.elapsed_time(b))
        tot.append(a.elapsed_time(c))
# back to back (no host synchronisation between calls): what a training loop sees
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for i in range(reps):
    s.grad = None
    cox.cox_loss(s, t, e).backward()
c.record()
torch.cuda.synchronize()
print(f"n={n} back-to-back fwd+bwd {a.elapsed_time(c) / reps:.4f} ms per call")
print(f"n={n} fwd {statistics.median(fw):.4f} ms  fwd+bwd {statistics.median(tot):.4f} ms  "
      f"-> {n * 112 / statistics.median(tot) / 1e6:.1f} GB/s  loss {float(loss.detach()):.6f}")
