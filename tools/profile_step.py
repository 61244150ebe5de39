#!/usr/bin/env python
"""One warm + one measured pass of each hot-path piece, for `ncu` launch lists.
    python tools/profile_step.py [resnet|cox|agg] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodalbrainsurvival_b200 import aggregate, cox, models, resnet  # noqa: E402
from oracle import resnet_oracle  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "resnet"
dev = torch.device("cuda:0")
if what == "resnet":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    net = resnet.resnet50()
    net.load_state_dict(resnet_oracle.init_state_dict(seed=1111))
    model = models.AggregationModel(net, models.Identity(), 2048, 2048, 1).to(dev).eval()
    x = torch.randn(B, 1, 3, 224, 224, device=dev)
    with torch.no_grad():
        for _ in range(2):
            f, _ = model.extract(x)
    torch.cuda.synchronize()
    print("features", tuple(f.shape), float(f.mean()))
elif what == "cox":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    s = torch.randn(n, device=dev, requires_grad=True)
    t = torch.rand(n, device=dev) * 200
    e = (torch.rand(n, device=dev) < 0.6).float()
    for _ in range(2):
        s.grad = None
        loss = cox.cox_loss(s, t, e)
        loss.backward()
    torch.cuda.synchronize()
    print("loss", float(loss.detach()))
    if os.environ.get("MMBS_TIME", "0") == "1":
        ts = []
        for _ in range(7):
            s.grad = None
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss = cox.cox_loss(s, t, e)
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            loss.backward()
            b.record()
            torch.cuda.synchronize()
            ts.append((a.elapsed_time(b), a.elapsed_time(m)))
        ts.sort()
        print("cox fwd+bwd ms (median, fwd part):", ts[len(ts) // 2], "=> GB/s on 112 B/sample:", n * 112 / ts[len(ts) // 2][0] / 1e6)
elif what == "agg":
    n, d, g = 100_000, 2048, 1000
    v = torch.randn(n, d, device=dev)
    seg = (torch.arange(n, device=dev) % g).to(torch.int32)
    for _ in range(2):
        m, c, l = aggregate.segmented_mean(v, seg, g)
    torch.cuda.synchronize()
    print("mean", float(m.mean()))
elif what == "train":
    # one fine-tuning step of the trunk (model.train(), fc + layer4 trainable): forward + backward
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    net = resnet.resnet50()
    net.load_state_dict(resnet_oracle.init_state_dict(seed=1111, bn3_gamma_scale=0.1))
    net = net.to(dev).train()
    for p in net.parameters():
        p.requires_grad = False
    for p in net.layer4.parameters():
        p.requires_grad = True
    x = torch.randn(B, 3, 224, 224, device=dev)
    gw = torch.randn(B, 2048, device=dev)
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        f = net.forward_extract(x)
        (f * gw).sum().backward()
    torch.cuda.synchronize()
    print("features", tuple(f.shape), float(f.mean()))
