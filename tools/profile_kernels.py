#!/usr/bin/env python
"""One launch of each representative hot kernel for an `ncu --set full` capture:
   3x3 conv (tensor-bound), 1x1 conv with residual, training conv with statistics, bn_train_apply (HBM-bound),
   segmented mean (HBM-bound), Cox scan kernels at 10 M.   python tools/profile_kernels.py"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench_conv
from multimodalbrainsurvival_b200 import aggregate, cox

bench_conv.bench(512, 14, 14, 256, 256, 3, 1, False, reps=1, warm=0)      # layer3 3x3
bench_conv.bench(512, 14, 14, 256, 1024, 1, 1, True, reps=1, warm=0)      # layer3 conv3 + residual
bench_conv.bench(512, 14, 14, 1024, 256, 1, 1, False, reps=1, warm=0)     # layer3 conv1
bench_conv.bench(128, 56, 56, 64, 256, 1, 1, False, stats=True, reps=1, warm=0)   # training conv with statistics
dev = "cuda"
n, d, g = 200_000, 2048, 2000
v = torch.randn(n, d, device=dev)
seg = (torch.arange(n, device=dev) % g).to(torch.int32)
aggregate.segmented_mean(v, seg, g)
s = torch.randn(10_000_000, device=dev, requires_grad=True)
t = torch.rand(10_000_000, device=dev) * 200
e = (torch.rand(10_000_000, device=dev) < 0.6).float()
cox.cox_loss(s, t, e).backward()
torch.cuda.synchronize()
