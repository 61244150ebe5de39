#!/usr/bin/env python
"""Diagnostics for the training-mode trunk: per-tensor errors vs the fp32 oracle at a small batch, then
timing of forward+backward at the training batch size.   python tools/check_train.py [B_time]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multimodalbrainsurvival_b200 import resnet  # noqa: E402
from oracle import resnet_oracle  # noqa: E402


def model(sd):
    net = resnet.resnet50(pretrained=False)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    for p in net.parameters():
        p.requires_grad = False
    for m in (net.fc, net.layer4):
        for p in m.parameters():
            p.requires_grad = True
    return net


def main():
    sd = resnet_oracle.init_state_dict(seed=31, bn3_gamma_scale=0.1)
    net = model(sd)
    torch.manual_seed(5)
    x = torch.randn(6, 3, 224, 224)
    gw = torch.randn(6, 2048)
    f = net.forward_extract(x.cuda())
    (f * gw.cuda()).sum().backward()
    torch.cuda.synchronize()
    fo, go, so = resnet_oracle.train_step(sd, x, gw)
    fe, ge, se = resnet_oracle.train_step(sd, x, gw, emulate_bf16=True)
    print("features rel vs fp32", float((f.detach().cpu() - fo).norm() / fo.norm()),
          "vs bf16-emulating oracle", float((f.detach().cpu() - fe).norm() / fe.norm()))
    for name, p in net.layer4.named_parameters():
        a, b = p.grad.detach().cpu().double().flatten(), go["layer4." + name].double().flatten()
        c = ge["layer4." + name].double().flatten()
        print(f"{name:28s} vs fp32: rel {float((a - b).norm() / b.norm()):.4f} cos {float(torch.dot(a, b) / (a.norm() * b.norm())):.5f}"
              f"   vs emu: rel {float((a - c).norm() / c.norm()):.4f} cos {float(torch.dot(a, c) / (a.norm() * c.norm())):.5f}"
              f"   |ref| {float(b.norm()):.3e}")
    worst = max(float((net.get_buffer(k).cpu() - v).norm() / v.norm()) for k, v in so.items())
    print("running stats worst rel", worst)

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    del net
    torch.cuda.empty_cache()
    net = model(sd)
    x = torch.randn(B, 3, 224, 224, device="cuda")
    gw = torch.randn(B, 2048, device="cuda")
    for it in range(6):
        if it == 2:
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            t0 = time.time()
            fw = bw = 0.0
        net.zero_grad(set_to_none=True)
        if it >= 2:
            e0.record()
        f = net.forward_extract(x)
        if it >= 2:
            e1.record()
        (f * gw).sum().backward()
        if it >= 2:
            e2.record()
            torch.cuda.synchronize()
            fw += e0.elapsed_time(e1)
            bw += e1.elapsed_time(e2)
    n = 4
    print(f"B={B}: forward {fw / n:.2f} ms  backward {bw / n:.2f} ms  wall/step {(time.time() - t0) / n * 1e3:.2f} ms")
    os.environ["MMBS_RESNET_TRAIN"] = "0"
    torch.backends.cudnn.benchmark = True
    for it in range(5):
        if it == 2:
            torch.cuda.synchronize()
            t0 = time.time()
        net.zero_grad(set_to_none=True)
        f = net.forward_extract(x)
        (f * gw).sum().backward()
    torch.cuda.synchronize()
    print(f"B={B}: torch eager fp32 module graph {(time.time() - t0) / 3 * 1e3:.2f} ms/step")


if __name__ == "__main__":
    main()
