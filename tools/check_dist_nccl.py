#!/usr/bin/env python
"""Multi-GPU check of multimodalbrainsurvival_b200.dist on real devices (NCCL over NVLink).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist_nccl.py
Every rank holds a shard of a synthetic cohort; the all-gathered global Cox loss and the
SUM-reduced gradient of a shared parameter must equal the single-GPU result on the whole cohort,
and the distributed per-case aggregation must equal the single-GPU aggregation."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from multimodalbrainsurvival_b200 import aggregate, cox  # noqa: E402
from multimodalbrainsurvival_b200 import dist as mdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for n_all in (1024, 200_003):
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(n_all, 4, generator=g)
    t_all = torch.randint(0, 300, (n_all,), generator=g).float()          # ties across ranks
    e_all = (torch.rand(n_all, generator=g) < 0.6).float()
    bounds = [n_all * r // world + (3 * r if r else 0) for r in range(world)] + [n_all]   # unequal shards
    sl = slice(bounds[rank], bounds[rank + 1])
    w = torch.full((4,), 0.3, device=dev, requires_grad=True)
    loss = mdist.global_cox_loss(x_all[sl].to(dev) @ w, t_all[sl].to(dev), e_all[sl].to(dev))
    loss.backward()
    mdist.allreduce_gradients([w])
    w1 = torch.full((4,), 0.3, device=dev, requires_grad=True)
    ref = cox.cox_loss(x_all.to(dev) @ w1, t_all.to(dev), e_all.to(dev))
    ref.backward()
    dl = abs(float(loss.detach()) - float(ref.detach()))
    dg = float((w.grad - w1.grad).abs().max() / w1.grad.abs().max())
    good = dl <= 1e-6 * abs(float(ref.detach())) + 1e-7 and dg <= 2e-5
    ok &= good
    print(f"[rank {rank}] n={n_all}: global loss {float(loss.detach()):.7f} ref {float(ref.detach()):.7f} "
          f"rel grad err {dg:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
# data-parallel MLP training: weight gradients built from all-gathered operands (dist.enable_factored_mlp_gradients)
# + SUM all-reduce of what is left must equal the single-GPU gradients of the whole batch
import torch.nn as nn  # noqa: E402
from multimodalbrainsurvival_b200 import mlp, models  # noqa: E402
mdist.enable_factored_mlp_gradients()
torch.manual_seed(11)                                   # same weights on every rank
def make():
    torch.manual_seed(11)
    return (nn.Sequential(nn.Dropout(0.0), nn.Linear(2048, 1024), nn.ReLU(), nn.Dropout(0.0), nn.Linear(1024, 256)),
            nn.Sequential(nn.Linear(256, 1)))
per = 128
g = torch.Generator().manual_seed(5)
x_all = torch.randn(per * world, 2048, generator=g)
t_all = torch.rand(per * world, generator=g) * 100
e_all = (torch.rand(per * world, generator=g) < 0.6).float()
rna, head = make()
model = models.RNAOnlyModel(rna, head).to(dev).train()
sl = slice(per * rank, per * (rank + 1))
out = model(x_all[sl].to(dev))
loss = mdist.global_cox_loss(out.view(-1), t_all[sl].to(dev), e_all[sl].to(dev), equal_sizes=True)
loss.backward()
n_global = sum(bool(getattr(p_, mdist.GLOBAL_GRAD_ATTR, False)) for p_ in model.parameters())
mdist.allreduce_gradients(list(model.parameters()))
rna1, head1 = make()
ref_model = nn.Sequential(rna1, head1).to(dev).train()   # stock torch modules, fp32, the whole batch on one GPU
ref_loss = cox.cox_loss(ref_model(x_all.to(dev)).view(-1), t_all.to(dev), e_all.to(dev))
ref_loss.backward()
good = n_global >= 1 and abs(float(loss.detach()) - float(ref_loss.detach())) <= 2e-2 * abs(float(ref_loss.detach()))
# (the Cox loss is shift invariant: the bias gradients of the last layers are ~0 - errors are measured against the
#  largest gradient norm of the model, not against those)
gmax = max(float(p2.grad.norm()) for p2 in ref_model.parameters())
for (n1, p1), (n2, p2) in zip(model.named_parameters(), ref_model.named_parameters()):
    rel = float((p1.grad - p2.grad).norm() / max(float(p2.grad.norm()), 1e-2 * gmax))
    good &= rel <= 6e-2   # bf16 operands through two layers (the all-reduced first-layer bias shows the same level)
    if rank == 0:
        print(f"[rank 0] factored-gradient MLP {n1}: rel err {rel:.2e} (global-batch wgrad: {getattr(p1, mdist.GLOBAL_GRAD_ATTR, False)})", flush=True)
ok &= good
print(f"[rank {rank}] data-parallel MLP gradients {'OK' if good else 'MISMATCH'}", flush=True)
mdist.enable_factored_mlp_gradients(False)

# distributed per-case aggregation (cases split across ranks)
n, d = 5000, 256
g = torch.Generator().manual_seed(3)
feats = torch.randn(n, d, generator=g)
cases = [f"case{int(i) % 37:02d}" for i in torch.randint(0, 1000, (n,), generator=g)]
sl = slice(n * rank // world, n * (rank + 1) // world)
uniq, means = mdist.aggregate_case_features_distributed(feats[sl].to(dev), cases[sl.start:sl.stop])
u1, m1 = aggregate.aggregate_case_features(feats.to(dev), cases, case_order=sorted(set(cases)))
good = uniq == u1 and np.allclose(means, m1, rtol=2e-5, atol=2e-6)
ok &= good
print(f"[rank {rank}] distributed aggregation {'OK' if good else 'MISMATCH'}", flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
