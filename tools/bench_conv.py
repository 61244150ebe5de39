#!/usr/bin/env python
"""Micro-benchmark of single conv plans: python tools/bench_conv.py  (prints us and TFLOP/s per case)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import engine

def bench(B, H, W, Cin, Cout, k, s, res, relu=True, f32=False, stats=False, reps=20, warm=3):
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, k * k * Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
    Ho, Wo = H // s, W // s
    out = torch.empty(B, Ho, Wo, Cout, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    r = torch.randn(B, Ho, Wo, Cout, device=dev).to(torch.bfloat16) if res else None
    sc = None if stats else torch.rand(Cout, device=dev) + 0.5
    sh = None if stats else torch.randn(Cout, device=dev)
    st = torch.zeros(2, Cout, device=dev) if stats else None
    plan = engine.conv_plan(x, w, out, ksize=k, stride=s, c_in=Cin, scale=sc, shift=sh, residual=r,
                            relu=relu and not stats, stats=st)
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        plan.run()
    ts = []
    for _ in range(reps):
        big.zero_()   # flush L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    fl = 2.0 * B * Ho * Wo * Cout * k * k * Cin
    byt = (x.numel() + out.numel() * (2 if res else 1)) * 2
    print(f"B{B} {H}x{W} {Cin}->{Cout} k{k}s{s} res={int(res)} stats={int(stats)}: {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s  "
          f"{byt / us / 1e3:6.0f} GB/s", flush=True)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    bench(B, 7, 7, 512, 2048, 1, 1, True)      # layer4 conv3
    bench(B, 7, 7, 512, 2048, 1, 1, False)
    bench(B, 7, 7, 512, 2048, 1, 1, False, relu=False)
    bench(B, 14, 14, 256, 1024, 1, 1, True)    # layer3 conv3
    bench(B, 14, 14, 256, 1024, 1, 1, False)
    bench(B, 14, 14, 1024, 256, 1, 1, False)   # layer3 conv1
    bench(B, 7, 7, 2048, 512, 1, 1, False)     # layer4 conv1
    bench(B, 14, 14, 256, 256, 3, 1, False)    # layer3 conv2
    bench(B, 7, 7, 512, 512, 3, 1, False)      # layer4 conv2
    bench(B, 28, 28, 128, 512, 1, 1, True)     # layer2 conv3
