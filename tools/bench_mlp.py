#!/usr/bin/env python
"""Early-fusion MLP (4096 -> 2048 -> 200 -> 1) inference over a large cohort shard (BASELINE config 4):
TFLOP/s of the fused tcgen05 path.   python tools/bench_mlp.py [rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from multimodalbrainsurvival_b200 import models  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
early = nn.Sequential(nn.Dropout(), nn.Linear(4096, 2048), nn.ReLU(), nn.Dropout(), nn.Linear(2048, 200), nn.ReLU(),
                      nn.Dropout(), nn.Linear(200, 1))
acc = models.accelerate(early).cuda().eval()
x = torch.randn(M, 4096, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = acc(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        y = acc(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    ref = early.cuda().eval()(x[:4096])
flops = 2.0 * M * (4096 * 2048 + 2048 * 200 + 200)
print(f"rows {M}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (incl. fp32->bf16 cast of the input)  "
      f"max err vs fp32 module {float((y[:4096] - ref).abs().max()):.3e} (scale {float(ref.abs().max()):.3e})")
