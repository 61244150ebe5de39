#!/usr/bin/env python
"""One launch of three conv shapes for an ncu pass that asks where the short-K 1x1 convs are bound (L2 -> SM bytes, L2 and
tensor-pipe utilisation):  layer3 conv1 (1024 -> 256), layer3 conv3 (256 -> 1024), layer3 3x3 (256 -> 256) as the control."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_conv

bench_conv.bench(512, 14, 14, 1024, 256, 1, 1, False, reps=1, warm=1)
bench_conv.bench(512, 14, 14, 256, 1024, 1, 1, False, reps=1, warm=1)
bench_conv.bench(512, 14, 14, 256, 256, 3, 1, False, reps=1, warm=1)
