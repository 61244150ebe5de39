"""Experiment: where does the e2e loop lose time vs the resident loop?"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import aggregate, models, pipeline, resnet
from oracle import resnet_oracle
dev = torch.device("cuda:0")
net = resnet.resnet50(); net.load_state_dict(resnet_oracle.init_state_dict(seed=1111))
model = models.AggregationModel(net, models.Identity(), 2048, 2048, 1).to(dev).eval()
B = 512
host = [torch.randn(B, 1, 3, 224, 224).pin_memory() for _ in range(2)]
host_u8 = [torch.randint(0, 256, (B, 1, 3, 224, 224), dtype=torch.uint8).pin_memory() for _ in range(2)]
xs = [h.to(dev) for h in host]
host_out = torch.empty(B, 2048).pin_memory()
seg = (torch.arange(B, device=dev) // 100).to(torch.int32)

def timed(fn, n=10):
    fn(3); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); fn(n); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n

def resident(n, d2h=False, agg=True):
    for i in range(n):
        with torch.no_grad():
            f, _ = model.extract(xs[i % 2])
        if d2h: host_out.copy_(f, non_blocking=True)
        if agg: aggregate.segmented_mean(f, seg, 6)

def e2e(n, src=host, d2h=True, agg=True, depth=2):
    for x in pipeline.prefetch_to_device((src[i % 2] for i in range(n)), dev, depth=depth):
        with torch.no_grad():
            f, _ = model.extract(x)
        if d2h: host_out.copy_(f, non_blocking=True)
        if agg: aggregate.segmented_mean(f, seg, 6)

def h2d_only(n, src=host):
    for x in pipeline.prefetch_to_device((src[i % 2] for i in range(n)), dev, depth=2):
        pass

print("resident              ", timed(lambda n: resident(n)))
print("resident no agg       ", timed(lambda n: resident(n, agg=False)))
print("resident + d2h        ", timed(lambda n: resident(n, d2h=True)))
print("h2d only fp32         ", timed(lambda n: h2d_only(n)))
print("h2d only u8           ", timed(lambda n: h2d_only(n, host_u8)))
print("e2e fp32              ", timed(lambda n: e2e(n)))
print("e2e fp32 no d2h       ", timed(lambda n: e2e(n, d2h=False)))
print("e2e fp32 depth 3      ", timed(lambda n: e2e(n, depth=3)))
print("e2e u8                ", timed(lambda n: e2e(n, host_u8)))
print("e2e u8 no d2h no agg  ", timed(lambda n: e2e(n, host_u8, d2h=False, agg=False)))
