"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in multimodalbrainsurvival_b200.dist:
the all-gathered global risk set, SUM gradient reduction, and distributed case aggregation.
The arithmetic kernels are replaced by oracle-grade torch/numpy stand-ins (injected), so these
tests check the plumbing, ordering and gradient semantics only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torch_cox(scores, times, status):
    _, idx = torch.sort(-times, stable=True)
    cs = scores[idx]
    cs = cs - torch.max(cs)
    return (-(cs - torch.log(torch.cumsum(torch.exp(cs), 0) + 1e-5)) * status[idx]).mean()


def _segmean_cpu(values, seg, n_seg):
    v = values.reshape(values.shape[0], -1).double()
    acc = torch.zeros(n_seg, v.shape[1], dtype=torch.float64).index_add_(0, seg.long(), v)
    cnt = torch.bincount(seg.long(), minlength=n_seg)
    return (acc / cnt[:, None]).float(), cnt.to(torch.int32), None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodalbrainsurvival_b200 import dist as mdist
        g = torch.Generator().manual_seed(0)
        n_all = 37
        s_all = torch.randn(n_all, generator=g)
        t_all = torch.randint(0, 10, (n_all,), generator=g).float()      # ties across ranks
        e_all = (torch.rand(n_all, generator=g) < 0.6).float()
        cut = 20                                                         # unequal shards: 20 + 17
        sl = slice(0, cut) if rank == 0 else slice(cut, n_all)
        w = torch.ones(3, requires_grad=True)                            # a shared "model" parameter
        feats = torch.stack([s_all[sl], s_all[sl] ** 2, torch.ones_like(s_all[sl])], 1)
        local_scores = feats @ w
        loss = mdist.global_cox_loss(local_scores, t_all[sl], e_all[sl], loss_fn=_torch_cox)
        loss.backward()
        mdist.allreduce_gradients([w])
        # single-process reference on the concatenated batch
        w2 = torch.ones(3, requires_grad=True)
        f_all = torch.stack([s_all, s_all ** 2, torch.ones_like(s_all)], 1)
        ref = _torch_cox(f_all @ w2, t_all, e_all)
        ref.backward()
        ok_loss = abs(float(loss) - float(ref)) < 1e-6
        ok_grad = bool(torch.allclose(w.grad, w2.grad, atol=1e-6))
        # distributed case aggregation with cases split across ranks
        cases_all = [f"c{i % 5}" for i in range(n_all)]
        fz = torch.arange(n_all * 4, dtype=torch.float32).view(n_all, 4)
        uniq, means = mdist.aggregate_case_features_distributed(fz[sl], cases_all[sl.start:sl.stop],
                                                                segmented_mean=_segmean_cpu)
        ref_means = np.stack([fz[[i for i, c in enumerate(cases_all) if c == u]].double().mean(0).numpy() for u in uniq])
        ok_agg = uniq == sorted(set(cases_all)) and np.allclose(means, ref_means)
        own = mdist.shard_cases(cases_all)
        q.put((rank, ok_loss, ok_grad, ok_agg, sorted({cases_all[i] for i in own})))
    finally:
        dist.destroy_process_group()


def test_global_risk_set_and_gradient_sum_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = []
    for rank, ok_loss, ok_grad, ok_agg, cases in res:
        assert ok_loss, f"rank {rank}: global loss differs from the single-process loss"
        assert ok_grad, f"rank {rank}: SUM-reduced gradient differs from the single-process gradient"
        assert ok_agg, f"rank {rank}: distributed case means differ"
        owned += cases
    assert sorted(owned) == [f"c{i}" for i in range(5)]      # every case owned by exactly one rank
