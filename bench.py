#!/usr/bin/env python
"""Headline benchmark (BASELINE.json configs[1]): ResNet-50 patch feature extraction,
224x224 synthetic patches, batch 512 per GPU, bf16 storage / fp32 accumulate, followed by
the per-case aggregation of that batch's features.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one batch of 512 patches per GPU through AggregationModel.extract (the
reference's extract_features hot loop, 1_HistoPathology/4_HistoPath_extractfeatures.py:60-71)
plus the segmented per-case mean of the batch's features (:80-88).  Prints ONE JSON line.

* value      patches/s over all GPUs, inputs resident in HBM (CUDA events, max over ranks)
* e2e        same metric through the public API with HOST (pinned) inputs: H2D copy of the
             fp32 batch and D2H of the features inside the timed region
* roofline   the tcgen05 conv kernel: 8.174 GFLOP/patch over the summed
             device time of its launches, against MEASURED_PEAKS.json (sustained bf16)
* cpu_baseline / --impl reference: the oracle port of the reference's CPU path
             (oracle/resnet_oracle.py, fp32 torch on all host cores), bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_PATCH = 8.174          # 2*MAC over the 53 convs at 224x224
BATCH = 512
PATCHES_PER_CASE = 100
CPU_SAMPLE = 32                  # patches per CPU-baseline pass (bounded sample)
_OUT = sys.stdout
METRIC = "resnet50_patch_feature_extraction_throughput"
UNIT = "patches/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tensor": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "tensor_burst": float(d.get("bf16_tflops", 1590.0)),
                "hbm": float(d.get("hbm_gbs", 6650.0)), "source": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_state_dict():
    """CPU arm only: the oracle's seeded state_dict (reference key names) for the oracle's functional forward."""
    from oracle import resnet_oracle
    return resnet_oracle.init_state_dict(seed=1111)


# --------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(sd, x):
    from oracle import resnet_oracle
    return resnet_oracle.forward_extract(sd, x)


def time_cpu(steps, warmup, sample=CPU_SAMPLE):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict()
    x = torch.randn(sample, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        cpu_reference_pass(sd, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(sd, x)
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} passes of {sample} patches (fp32 torch CPU, oracle/resnet_oracle.py)",
            "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)   # every step = one pass over the bounded 32-patch sample
    r = time_cpu(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "resnet50_extract_224_b512_bf16 (CPU sample of %d patches per step)" % CPU_SAMPLE},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


# --------------------------------------------------------------------------- GPU arm
def cox_secondary(torch, dev, peaks):
    """Cox fwd+bwd on the 10 M-sample cohort (BASELINE config 4): ms and fraction of HBM peak
    on the 112 B/sample algorithmic traffic."""
    from multimodalbrainsurvival_b200 import cox
    n = 10_000_000
    g = torch.Generator(device=dev).manual_seed(1111)
    s = torch.randn(n, device=dev, generator=g).requires_grad_(True)
    t = torch.rand(n, device=dev, generator=g) * 200
    e = (torch.rand(n, device=dev, generator=g) < 0.6).float()
    times = []
    for i in range(6):
        s.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = cox.cox_loss(s, t, e)
        loss.backward()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(a.elapsed_time(b))
    ms = statistics.median(times)
    gbs = n * 112 / (ms * 1e-3) / 1e9
    # the same call back to back, no host synchronisation in between (what a training loop sees: the host-side work of
    # call i+1 hides behind the kernels of call i)
    reps = 8
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        s.grad = None
        cox.cox_loss(s, t, e).backward()
    b.record()
    torch.cuda.synchronize()
    ms_b2b = a.elapsed_time(b) / reps
    return {"workload": "cox_fwd_bwd_10M", "ms": ms, "alg_bytes_per_sample": 112, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / peaks["hbm"], "timing": "median of 4 calls, host synchronised before each",
            "ms_back_to_back": ms_b2b, "frac_of_hbm_peak_back_to_back": n * 112 / (ms_b2b * 1e-3) / 1e9 / peaks["hbm"],
            "loss": float(loss.detach())}


def aggregation_secondary(torch, dev, peaks):
    """Per-case mean of patch features (extract_features tail): 200 k patches x 2048 features, 100 patches per case.
    Algorithmic bytes N*(4D+4) + G*4D over the device time of mmbs_segmented_mean."""
    from multimodalbrainsurvival_b200 import aggregate
    n, d = 200_000, 2048
    g = n // PATCHES_PER_CASE
    v = torch.randn(n, d, device=dev)
    seg = (torch.arange(n, device=dev) // PATCHES_PER_CASE).to(torch.int32)
    times = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        aggregate.segmented_mean(v, seg, g)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(a.elapsed_time(b))
    ms = statistics.median(times)
    gbs = (n * (4 * d + 4) + g * 4 * d) / (ms * 1e-3) / 1e9
    return {"workload": "segmented_mean_200k_x_2048", "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm"]}


def eager_cox(scores, times, status):
    """The reference's loss written with stock torch ops (1_HistoPathology/models.py:90-111 restated: sort by
    descending time, shift by the max, exp, cumsum, log, event mask, NaN check with its host sync, mean) - the
    library competitor of csrc/cox.cu on the same GPU."""
    import torch
    order = torch.sort(-times)[1]
    s, d = scores[order], status[order]
    s = s - s.max()
    terms = -(s - (s.exp().cumsum(0) + 1e-5).log()) * d
    if bool((terms != terms).any()):
        raise FloatingPointError("NaN in the Cox terms")
    return terms.mean()


def gpu_eager_secondary(torch, dev, net, steps):
    """Same-box library baselines (the bar on the box is PyTorch eager over cuDNN / cuBLAS / CUB): the
    UNMODIFIED architecture through stock torch modules in bf16 channels_last, torch.sort + cumsum Cox at 10 M and an
    eager RNA training step.  None of libmmbs runs here (MMBS kernels are bypassed through the stock module graph)."""
    import copy
    import torch.nn as nn
    out = {}
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        eager = copy.deepcopy(net).to(dev).eval().to(memory_format=torch.channels_last).bfloat16()
        xs = [torch.randn(BATCH, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
              for _ in range(2)]
        with torch.no_grad():
            for i in range(3):
                eager._features_torch(xs[i % 2])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(steps):
                eager._features_torch(xs[i % 2])
            b.record()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out["extract"] = {"workload": "resnet50 forward_extract, stock torch modules, bf16 channels_last, cuDNN (benchmark mode), B=512",
                          "ms_per_step": ms, "patches_per_s": BATCH / (ms * 1e-3)}
        del eager, xs
        torch.cuda.empty_cache()
    except Exception as ex:
        out["extract"] = {"error": repr(ex)}
    finally:
        torch.backends.cudnn.benchmark = old_bench
    try:
        n = 10_000_000
        g = torch.Generator(device=dev).manual_seed(1111)
        s = torch.randn(n, device=dev, generator=g).requires_grad_(True)
        t = torch.rand(n, device=dev, generator=g) * 200
        e = (torch.rand(n, device=dev, generator=g) < 0.6).float()
        times = []
        for i in range(5):
            s.grad = None
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eager_cox(s, t, e).backward()
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(a.elapsed_time(b))
        out["cox"] = {"workload": "cox fwd+bwd 10M, torch.sort + gather + cumsum + autograd (fp32)", "ms": statistics.median(times)}
        del s, t, e
        torch.cuda.empty_cache()
    except Exception as ex:
        out["cox"] = {"error": repr(ex)}
    try:
        torch.manual_seed(3333)
        mlp = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048),
                            nn.Linear(2048, 1)).to(dev).train()
        opt = torch.optim.Adam(mlp.parameters(), lr=1e-6, weight_decay=1e-5)
        x = torch.randn(128, 12778, device=dev)
        t = torch.rand(128, device=dev) * 200
        e = (torch.rand(128, device=dev) < 0.6).float()

        def one():
            opt.zero_grad(set_to_none=True)
            eager_cox(mlp(x).view(-1), t, e).backward()
            opt.step()
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            one()
        b.record()
        torch.cuda.synchronize()
        out["rna_step"] = {"workload": "RNA MLP 12778-4096-2048-1 + Cox + torch.optim.Adam, fp32 eager, B=128",
                           "ms_per_step": a.elapsed_time(b) / 10}
    except Exception as ex:
        out["rna_step"] = {"error": repr(ex)}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from multimodalbrainsurvival_b200 import _lib, aggregate, engine, models, pipeline, resnet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    torch.manual_seed(1111)
    net = resnet.randomize_batchnorm_(resnet.resnet50(), seed=1111)   # random-init weights of the named architecture
    model = models.AggregationModel(net, models.Identity(), 2048, 2048, 1).to(dev).eval()

    B = BATCH
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # two resident input batches (2 x 308 MB > L2) alternated between steps
    xs = [torch.randn(B, 1, 3, 224, 224, device=dev, generator=gen) for _ in range(2)]
    n_cases = (B + PATCHES_PER_CASE - 1) // PATCHES_PER_CASE
    seg = (torch.arange(B, device=dev) // PATCHES_PER_CASE).to(torch.int32)
    host = [torch.randn(B, 1, 3, 224, 224).pin_memory() for _ in range(2)]
    host_out = torch.empty(B, 2048).pin_memory()

    def step(i):
        with torch.no_grad():
            feats, _ = model.extract(xs[i % 2])
        return aggregate.segmented_mean(feats, seg, n_cases)[0]

    host_u8 = [torch.randint(0, 256, (B, 1, 3, 224, 224), dtype=torch.uint8).pin_memory() for _ in range(2)]
    writer = pipeline.HostWriter(dev)

    def run_e2e(steps, src=None):
        """Public-API loop with HOST inputs: pinned fp32 batches are staged by
        pipeline.prefetch_to_device (H2D of batch i+1 overlaps the kernels of batch i), features
        are read back to pinned host memory every step."""
        outs = None
        src = host if src is None else src
        batches = (src[i % 2] for i in range(steps))
        for x in pipeline.prefetch_to_device(batches, dev, depth=2):
            with torch.no_grad():
                feats, _ = model.extract(x)
            writer.write(feats, host_out)          # D2H of every step's features, on a side stream
            outs = aggregate.segmented_mean(feats, seg, n_cases)[0]
        writer.wait()
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = _lib.launch_count() + engine.GRAPH_LAUNCHES
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            a.record()
            for i in range(steps):
                fn(i)
            b.record()
            barrier()
        ms = a.elapsed_time(b)
        launches = _lib.launch_count() + engine.GRAPH_LAUNCHES - l0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, cs.summary()

    warmup = max(args.warmup, 3)
    ms, launches, clocks = timed(step, args.steps, warmup)
    def time_e2e(src):
        """K steps of the host-fed loop, timed on the device; median of 3 repeats (a single host hiccup - page
        faults in the pinned staging path, a descheduled Python thread - can double one 70 ms measurement)."""
        run_e2e(2, src)
        reps = []
        for _ in range(3):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run_e2e(args.steps, src)
            b.record()
            barrier()
            t_ms = a.elapsed_time(b)
            if world > 1:
                t = torch.tensor([t_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_ms = float(t.item())
            reps.append(t_ms)
        return statistics.median(reps)

    ms_e2e = time_e2e(host)
    # same loop fed with raw uint8 pixels (normalisation fused into the input pack kernel): 4x fewer H2D bytes
    ms_e2e_u8 = time_e2e(host_u8)
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # training-step secondaries (BASELINE configs 3 and 5): every rank takes part (global Cox + gradient all-reduce)
    train_sec = {}
    if os.environ.get("MMBS_BENCH_TRAIN", "1") == "1":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_train
        net._engines.clear()
        torch.cuda.empty_cache()
        for kind in ("histo", "joint"):
            try:
                train_sec["train_" + kind] = bench_train.run(kind, torch, dev, world, rank, steps=20)   # ~0.1 s
            except Exception as ex:  # secondary metrics must never kill the headline line
                train_sec["train_" + kind] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        for kind in ("rna", "early"):
            try:
                # (the RNA step is < 1 ms: a 5-step loop doubles on one hiccup of the box - 50 steps; the 50 ms early step: 5)
                train_sec["train_" + kind] = bench_train.run_mlp(kind, torch, dev, world, rank,
                                                                 steps=50 if kind == "rna" else 5)
            except Exception as ex:
                train_sec["train_" + kind] = {"error": repr(ex)}
            torch.cuda.empty_cache()

    line = None
    if rank == 0:
        # per-kernel device time of one step (events around every launch of the conv kernel)
        conv_ms = conv_kernel_time(torch, model, xs[0])
        achieved = B * GFLOP_PER_PATCH / (conv_ms * 1e-3) / 1e3 if conv_ms else None  # TFLOP/s
        traffic = None   # dram bytes per launch of the conv kernel, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("conv_dram_bytes_per_launch")
        # SURVEY 8d formula: patches/s per GPU x 8.174 GFLOP over the MEASURED BURST bf16 peak, from the very timed
        # region `value` comes from (the whole step: pack, convs, pools, aggregation); the kernel-only figure
        # (summed per-launch CUDA-event time of the 53 conv launches) and the sustained-peak fractions sit beside it
        step_tflops = (value / world) * GFLOP_PER_PATCH / 1e3
        roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, 53 launches/step)",
                    "achieved": step_tflops, "peak": peaks["tensor_burst"], "unit": "TFLOP/s",
                    "frac": step_tflops / peaks["tensor_burst"],
                    "frac_of_sustained_peak": step_tflops / peaks["tensor"],
                    "kernel_only": {"achieved": achieved, "frac_of_burst": (achieved / peaks["tensor_burst"]) if achieved else None,
                                    "frac_of_sustained": (achieved / peaks["tensor"]) if achieved else None},
                    "traffic": traffic,
                    "traffic_note": "ncu dram read+write bytes per conv launch (mean of the 53 launches of one "
                                    "512-patch step, profiles/r01_conv_traffic.json); algorithmic activation I/O "
                                    "is 54.6 MB/patch = 527 MB per launch",
                    "flops_per_launch": B * GFLOP_PER_PATCH * 1e9 / 53,
                    "peak_source": peaks["source"] + " (burst bf16 %.1f; sustained %.1f)" % (peaks["tensor_burst"], peaks["tensor"]),
                    "conv_ms_per_step": conv_ms}
        # the CPU arm is timed on rank 0 at N=1 only (at N>1 the other ranks busy-wait on the host cores)
        cpu = time_cpu(2, 1) if world == 1 else None
        try:
            cox_sec = cox_secondary(torch, dev, peaks)
        except Exception as ex:  # secondary metric must never kill the headline line
            cox_sec = {"error": repr(ex)}
        try:
            train_sec["aggregation"] = aggregation_secondary(torch, dev, peaks)
        except Exception as ex:
            train_sec["aggregation"] = {"error": repr(ex)}
        if os.environ.get("MMBS_BENCH_EAGER", "1") == "1":
            net._engines.clear()
            torch.cuda.empty_cache()
            train_sec["gpu_eager"] = gpu_eager_secondary(torch, dev, net, max(3, min(args.steps, 10)))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "resnet50_extract_224_b512_bf16", "batch_per_gpu": B,
                           "patches_per_case": PATCHES_PER_CASE, "chunk": int(os.environ.get("MMBS_RESNET_CHUNK", 0)) or "default",
                           "l2": "two alternating 308 MB input batches (> L2)", "parallelism": f"dp{world} (patches sharded, no collective)"},
                "roofline": roofline,
                "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
                # headline e2e: the loader hands RAW uint8 pixels to model.extract (one-line loader change documented in
                # PILToTensor() instead of ToTensor()+Normalize(); the normalisation runs in the device
                # pack kernel).  The unmodified loader's fp32 tensors are timed beside it (e2e_fp32): 4x the H2D bytes.
                "e2e": {"value": world * B * args.steps / (ms_e2e_u8 * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": B * 3 * 224 * 224, "d2h_bytes_per_step": B * 2048 * 4,
                        "ms_per_step": ms_e2e_u8 / args.steps,
                        "input": "pinned host uint8 pixels (ToTensor+Normalize fused into the device pack kernel)",
                        "timing": "median of 3 repeats of the K-step loop"},
                "e2e_fp32": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * 224 * 224 * 4,
                             "d2h_bytes_per_step": B * 2048 * 4, "ms_per_step": ms_e2e / args.steps,
                             "input": "pinned host fp32 normalised patches (what the reference's unmodified loader "
                                      "hands to model.extract)"},
                "gpu_launches": int(launches), "clocks": clocks, "secondary": dict({"cox": cox_sec}, **train_sec)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=_OUT, flush=True)


def conv_kernel_time(torch, model, x):
    """Sum of the device time of every conv_gemm_kernel launch in one step (CUDA events on the
    launching stream around each launch; CUDA-graph replay disabled for this instrumented pass)."""
    from multimodalbrainsurvival_b200 import engine
    net = model.resnet
    evs = []
    orig = engine.ConvPlan.run

    def timed_run(self):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(self)
        b.record()
        evs.append((a, b))

    saved_engines = dict(net._engines)
    old_env = os.environ.get("MMBS_CUDA_GRAPH")
    net._engines.clear()
    os.environ["MMBS_CUDA_GRAPH"] = "0"
    engine.ConvPlan.run = timed_run
    try:
        with torch.no_grad():
            model.extract(x)          # builds instrumented engines, warm
            torch.cuda.synchronize()
            evs.clear()
            model.extract(x)
        torch.cuda.synchronize()
        total = sum(a.elapsed_time(b) for a, b in evs)
    finally:
        engine.ConvPlan.run = orig
        if old_env is None:
            os.environ.pop("MMBS_CUDA_GRAPH", None)
        else:
            os.environ["MMBS_CUDA_GRAPH"] = old_env
        net._engines.clear()
        net._engines.update(saved_engines)
    return total


def main():
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL prints its version banner
    # on the first communicator) are sent to stderr; the JSON goes to the original stdout
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _OUT
    _OUT = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
