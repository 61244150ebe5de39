"""Multi-GPU host logic (one process per GPU, ``torch.distributed`` over NCCL/NVLink).

The reference is single-device; data parallelism is added here:

* training  - ``global_cox_loss``: the Cox risk set is made global by all-gathering the
  (time, event, score) triples before the sort+scan (the loss of reference
  ``cox_loss`` /root/reference/1_HistoPathology/models.py:90-111 over the concatenated batch);
  ``allreduce_gradients``: parameter gradients are SUM-reduced, because the loss is a mean over
  the GLOBAL batch and every rank back-propagates only its slice of the global gradient.
* extraction - patches shard by case; ``aggregate_case_features_distributed`` reduces the
  per-case (sum, count) accumulators once at the end (no collective on the data path).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


# ---- weight gradients of wide linear layers from all-gathered OPERANDS ----------------------------------------------
# dW = dz^T h is a sum over the batch, so SUM-all-reducing the (out x in) fp32 gradient of every rank (243 MB for the
# RNA model at batch 128 per GPU) can be replaced by all-gathering its two thin factors - dz [batch, out] and
# h [batch, in] in bf16 (1 + 3.3 MB per rank for the 12778 -> 4096 layer) - and running the weight-gradient GEMM over
# the GLOBAL batch on every rank: 8x the wgrad flops (still an HBM-bound outer product), ~50x fewer bytes on NVLink.
# mlp._MLPTrainEngine does this for every layer where it pays once enable_factored_mlp_gradients() was called; those
# weights come out of backward already summed over ranks and allreduce_gradients() skips them.
_FACTORED = {"enabled": False, "group": None}
GLOBAL_GRAD_ATTR = "_mmbs_grad_is_global"


def enable_factored_mlp_gradients(enabled=True, group=None):
    _FACTORED["enabled"], _FACTORED["group"] = bool(enabled), group


def factored_mlp_world():
    """(world size, group) when MLP weight gradients are to be built from all-gathered operands, else (1, None)."""
    if not _FACTORED["enabled"] or not (dist.is_available() and dist.is_initialized()):
        return 1, None
    return dist.get_world_size(_FACTORED["group"]), _FACTORED["group"]


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_risk_set(scores, times, status, group=None, equal_sizes=False):
    """All-gather the packed fp32 (score, time, status) triples of every rank.  Returns
    (all_scores, all_times, all_status, offset, n_local): ``all_scores[offset:offset+n_local]`` is
    this rank's live ``scores`` tensor (autograd flows into it), the rest are constants.
    ``equal_sizes=True`` (every rank holds the same number of samples, what a DistributedSampler
    guarantees) skips the size exchange and its host synchronisation: one collective, no sync."""
    world, rank = _world(group)
    s = scores.reshape(-1)
    t = times.reshape(-1).to(s.dtype)
    e = status.reshape(-1).to(s.dtype)
    n = s.numel()
    if world == 1:
        return s, t, e, 0, n
    if equal_sizes:
        sizes = [n] * world
    else:
        sizes = [torch.zeros(1, dtype=torch.int64, device=s.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=s.device), group=group)
        sizes = [int(x.item()) for x in sizes]
    nmax = max(sizes)
    packed = torch.zeros((nmax, 3), dtype=s.dtype, device=s.device)
    packed[:n, 0] = s.detach()
    packed[:n, 1] = t
    packed[:n, 2] = e
    out = torch.empty((world * nmax, 3), dtype=s.dtype, device=s.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, nmax, 3)
    parts_s, parts_t, parts_e = [], [], []
    for r in range(world):
        parts_s.append(s if r == rank else out[r, :sizes[r], 0])
        parts_t.append(out[r, :sizes[r], 1])
        parts_e.append(out[r, :sizes[r], 2])
    return (torch.cat(parts_s), torch.cat(parts_t), torch.cat(parts_e), sum(sizes[:rank]), n)


def global_cox_loss(scores, times, status, group=None, loss_fn=None, equal_sizes=False):
    """Cox loss over the risk set of ALL ranks.  Every rank returns the same global loss; its
    backward yields d(global loss)/d(local scores).  Ties across ranks are ordered by
    (rank, local index) = the order of the concatenated batch."""
    if loss_fn is None:
        from .cox import cox_loss as loss_fn
    all_s, all_t, all_e, _, _ = gather_risk_set(scores, times, status, group, equal_sizes=equal_sizes)
    return loss_fn(all_s, all_t, all_e)


def allreduce_gradients(params, group=None, bucket_bytes=256 << 20):
    """SUM-all-reduce ``.grad`` of ``params`` in flat fp32 buckets (NVSwitch: size buckets for
    launch latency, not link count)."""
    world, _ = _world(group)
    # (weights whose gradient was built from all-gathered operands are already summed over the ranks)
    grads = [p.grad for p in params if p.grad is not None and not getattr(p, GLOBAL_GRAD_ATTR, False)]
    if world == 1 or not grads:
        return
    if dist.get_backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
        # one NCCL group launch, in place on the gradient tensors: no flatten / copy-back passes
        try:
            with dist._coalescing_manager(group=group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            return
        except TypeError:   # private API signature moved (raised before any collective): bucketed path
            pass
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def shard_cases(case_list, group=None):
    """Indices of the rows this rank owns when patches are partitioned BY CASE (every case lives
    on one rank, so the per-case mean needs no collective)."""
    world, rank = _world(group)
    uniq = sorted(set(case_list))
    owner = {c: i % world for i, c in enumerate(uniq)}
    return [i for i, c in enumerate(case_list) if owner[c] == rank]


def aggregate_case_features_distributed(features, case_list, group=None, segmented_mean=None):
    """Per-case mean when the rows of a case may be spread over ranks: local (sum, count) per
    case, one SUM all-reduce of the accumulators, divide.  Returns (sorted case ids, means
    [n_cases, D] float64 numpy) on every rank."""
    world, _ = _world(group)
    if segmented_mean is None:
        from .aggregate import segmented_mean
    local_cases = list(case_list)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, sorted(set(local_cases)), group=group)
        uniq = sorted(set().union(*[set(g) for g in gathered]))
    else:
        uniq = sorted(set(local_cases))
    lut = {c: i for i, c in enumerate(uniq)}
    seg = torch.from_numpy(np.fromiter((lut[c] for c in local_cases), dtype=np.int32, count=len(local_cases)))
    seg = seg.to(features.device)
    d = features.reshape(features.shape[0], -1).shape[1]
    if len(local_cases):
        mean, counts, _ = segmented_mean(features, seg, len(uniq))
        cnt = counts.to(torch.float64)
        acc = torch.nan_to_num(mean.reshape(len(uniq), d).to(torch.float64)) * cnt[:, None]
    else:
        cnt = torch.zeros(len(uniq), dtype=torch.float64, device=features.device)
        acc = torch.zeros((len(uniq), d), dtype=torch.float64, device=features.device)
    if world > 1:
        buf = torch.cat([acc, cnt[:, None]], dim=1)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        acc, cnt = buf[:, :d], buf[:, d]
    return uniq, (acc / cnt[:, None]).cpu().numpy()
