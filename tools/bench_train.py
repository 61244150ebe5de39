#!/usr/bin/env python
"""Training-step throughput of BASELINE configs 3 and 5:

  histo : AggregationModel(resnet50, Identity) in train mode, fc + layer4 trainable, Cox loss, Adam
          (/root/reference/1_HistoPathology/2_HistoPath_train.py:365-395, config_ffpe_train.json), B = 128 / GPU
  joint : BagHistopathologyRNAModel(resnet50, rna MLP, Dropout(0.8)+Linear head), Cox loss, Adam
          (/root/reference/5_JointFusion/1_JointFusion_train.py:314-325,196-230), B = 128 / GPU

One step = forward (batch-statistics BatchNorm in every layer) + global Cox loss (risk set all-gathered over the
ranks) + backward through layer4 / the MLPs + SUM all-reduce of the parameter gradients + the scripts' torch.optim.Adam
stepping through the fused multi-tensor kernel (optim.accelerate_optimizer).
Inputs are device resident for `value`; `e2e` feeds pinned host batches (H2D inside the timed region) and reads
the loss back every step.  Importable (bench.py `secondary`) or  python tools/bench_train.py [histo|joint] [steps]
(under torchrun for N > 1)."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GFLOP_FWD, GFLOP_BWD_L4 = 8.174, 2.82      # per patch, SURVEY.md §8d config 3
MLP_GFLOP_STEP = 33.3                       # RNA MLP fwd+bwd at B = 128


def build(kind, torch, dev):
    import torch.nn as nn
    from multimodalbrainsurvival_b200 import models, resnet
    torch.manual_seed(1111)
    net = resnet.randomize_batchnorm_(resnet.resnet50(), seed=1111, bn3_gamma_scale=0.1)
    for p in net.parameters():
        p.requires_grad = False
    for layer in (net.fc, net.layer4):       # n_layers_to_train = 2
        for p in layer.parameters():
            p.requires_grad = True
    torch.manual_seed(1111)
    if kind == "histo":
        model = models.AggregationModel(net, models.Identity(), 2048, 2048, 1)
    else:
        rna = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048))
        head = nn.Sequential(nn.Dropout(0.8), nn.Linear(4096, 1))
        model = models.BagHistopathologyRNAModel(net, rna, head)
    return model.to(dev).train()


def _adam(torch, params):
    """The scripts' torch.optim.Adam (lr / weight_decay of the example configs), stepping through the fused
    multi-tensor kernel (optim.accelerate_optimizer; MMBS_BENCH_STOCK_ADAM=1 times torch's own foreach step)."""
    from multimodalbrainsurvival_b200 import optim
    opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5)
    return opt if os.environ.get("MMBS_BENCH_STOCK_ADAM", "0") == "1" else optim.accelerate_optimizer(opt)


def run_mlp(kind, torch, dev, world=1, rank=0, steps=5, warmup=3, rows=None):
    """kind 'rna'  : BASELINE config 1 - RNAOnlyModel (12778 -> 4096 -> 2048 -> 1), B = 128, Cox, Adam
                     (/root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257,150-170)
       kind 'early': BASELINE config 4 - early-fusion MLP (4096 -> 2048 -> 200 -> 1) as one full-batch step over a
                     cohort shard of `rows` samples per GPU (bf16 features generated on the device), Cox loss over the
                     risk set all-gathered from every rank (/root/reference/3_EarlyFusion/2_EarlyFusion_train.py:242-253)."""
    import torch.distributed as dist
    import torch.nn as nn
    from multimodalbrainsurvival_b200 import dist as mdist
    # N > 1: wide MLP weights get their gradient from all-gathered (dz, h) factors instead of a 243 MB all-reduce
    mdist.enable_factored_mlp_gradients(world > 1 and os.environ.get("MMBS_FACTORED_GRADS", "1") == "1")
    from multimodalbrainsurvival_b200 import _lib, models
    torch.manual_seed(1111)
    if kind == "rna":
        rows = rows or 128
        rna = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048))
        model = models.RNAOnlyModel(rna, nn.Sequential(nn.Linear(2048, 1))).to(dev).train()
        width, mflop = 12778, 260.0      # fwd + wgrad + dgrad (no dgrad into the input)
        dtype = torch.float32
    else:
        rows = rows or 1_250_000
        model = models.accelerate(nn.Sequential(nn.Dropout(), nn.Linear(4096, 2048), nn.ReLU(), nn.Dropout(),
                                                nn.Linear(2048, 200), nn.ReLU(), nn.Dropout(), nn.Linear(200, 1))
                                  ).to(dev).train()
        width, mflop = 4096, 36.0
        dtype = torch.bfloat16
    params = list(model.parameters())
    opt = _adam(torch, params)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    x = torch.randn(rows, width, device=dev, generator=g, dtype=torch.float32 if rows <= 4096 else torch.bfloat16).to(dtype)
    times = torch.rand(rows, device=dev, generator=g) * 200 + torch.arange(rows, device=dev) * 2.0 ** -20
    status = (torch.rand(rows, device=dev, generator=g) < 0.6).float()

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(x)
        loss = mdist.global_cox_loss(out.view(-1), times, status, equal_sizes=True)
        loss.backward()
        mdist.allreduce_gradients(params)
        opt.step()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync()
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    sync()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"workload": f"{kind}_mlp_cox_train_step ({rows} samples per GPU, risk set of {rows * world})",
            "ms_per_step": ms, "steps_per_s": 1e3 / ms, "samples_per_s": world * rows * 1e3 / ms,
            "tflops_per_gpu": rows * mflop * 1e6 / (ms * 1e-3) / 1e12, "n_gpus": world,
            "gpu_launches_per_step": (_lib.launch_count() - l0) / steps, "loss": float(loss.detach())}


def run(kind, torch, dev, world=1, rank=0, steps=10, warmup=3, batch=128):
    import torch.distributed as dist
    from multimodalbrainsurvival_b200 import dist as mdist
    # N > 1: wide MLP weights get their gradient from all-gathered (dz, h) factors instead of a 243 MB all-reduce
    mdist.enable_factored_mlp_gradients(world > 1 and os.environ.get("MMBS_FACTORED_GRADS", "1") == "1")
    from multimodalbrainsurvival_b200 import _lib, engine
    model = build(kind, torch, dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = _adam(torch, params)
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    xs = [torch.randn(batch, 1, 3, 224, 224, device=dev, generator=g) for _ in range(2)]
    rna = torch.randn(batch, 12778, device=dev, generator=g)
    times = torch.rand(batch, device=dev, generator=g) * 200
    status = (torch.rand(batch, device=dev, generator=g) < 0.6).float()
    host_x = [x.cpu().pin_memory() for x in xs]
    host_rna = rna.cpu().pin_memory()

    def step(x, r):
        opt.zero_grad(set_to_none=True)
        out = model(x)[0] if kind == "histo" else model(x, r)
        loss = mdist.global_cox_loss(out.view(-1), times, status, equal_sizes=True)
        loss.backward()
        mdist.allreduce_gradients(params)
        opt.step()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync()
        l0 = _lib.launch_count() + engine.GRAPH_LAUNCHES
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            last = fn(i)
        b.record()
        sync()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n, _lib.launch_count() + engine.GRAPH_LAUNCHES - l0, last

    for i in range(warmup):
        step(xs[i % 2], rna)
    ms, launches, loss = timed(lambda i: step(xs[i % 2], rna), steps)

    from multimodalbrainsurvival_b200 import pipeline

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        """Pinned host batches staged by pipeline.prefetch_to_device (the H2D copy of batch i+1 overlaps the
        kernels of batch i); the loss of EVERY step is copied to pinned host memory and read there one step later (what a
        logging training loop does: a blocking .item() per step would idle the GPU between steps)."""
        last = None
        # joint fusion: the RNA batch rides in the same prefetch slot as its patches (a copy issued on the compute stream
        # would queue behind the NEXT batch's 77 MB on the DMA engine: +1.6 ms per step)
        src = ((host_x[i % 2], host_rna) for i in range(n)) if kind == "joint" else (host_x[i % 2] for i in range(n))
        for i, item in enumerate(pipeline.prefetch_to_device(src, dev, depth=2)):
            x, r = item if kind == "joint" else (item, None)
            loss = step(x, r).detach()
            loss_host[i % 2].copy_(loss.reshape(1), non_blocking=True)      # D2H of this step's loss
            loss_done[i % 2].record()
            if i > 0:                                                        # .. read on the host one step later
                loss_done[(i - 1) % 2].synchronize()
                last = float(loss_host[(i - 1) % 2])
        loss_done[(n - 1) % 2].synchronize()
        return float(loss_host[(n - 1) % 2])

    e2e_loop(2)
    n_e2e = max(steps, 20)   # (a 10-step loop is dominated by the fill of the prefetch ring: +-1 ms run to run)
    ms_e2e, _, _ = timed(lambda i: e2e_loop(n_e2e) if i == 0 else None, 1)
    ms_e2e /= n_e2e
    gflop = batch * (GFLOP_FWD + GFLOP_BWD_L4) + (MLP_GFLOP_STEP * batch / 128 if kind == "joint" else 0.0)
    return {"workload": f"{kind}_cox_finetune_step_b{batch}_per_gpu (fc + layer4 trainable, Adam)",
            "steps_per_s": world * 1e3 / ms / world, "samples_per_s": world * batch * 1e3 / ms, "ms_per_step": ms,
            "tflops_per_gpu": gflop / ms, "gflop_per_step_per_gpu": gflop, "n_gpus": world,
            "e2e": {"ms_per_step": ms_e2e, "samples_per_s": world * batch * 1e3 / ms_e2e,
                    "h2d_bytes_per_step": batch * 3 * 224 * 224 * 4 + (batch * 12778 * 4 if kind == "joint" else 0),
                    "d2h_bytes_per_step": 4},
            "gpu_launches_per_step": launches / steps, "loss": float(loss.detach()),
            "collectives": "all-gather of (score,time,status) triples + all-gather of the MLP weight-gradient factors (dz, h) + SUM all-reduce of the remaining gradients" if world > 1 else "none (N=1)"}


def main():
    import torch
    import torch.distributed as dist
    kind = sys.argv[1] if len(sys.argv) > 1 else "histo"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = run_mlp(kind, torch, dev, world, rank, steps) if kind in ("rna", "early") else run(kind, torch, dev, world, rank, steps)
    if rank == 0:
        import json
        print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
