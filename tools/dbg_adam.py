import copy, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from multimodalbrainsurvival_b200 import optim
DEV = "cuda:0"
SHAPES = [(4096, 1277), (4096,), (2048, 4096), (2048,), (1, 2048), (1,), (7, 3, 3, 5), (333,), (1000003,)]
def make(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in SHAPES]
def groups(ps):
    return [{"params": ps[:4], "lr": 1e-3}, {"params": ps[4:6], "lr": 5e-2, "betas": (0.8, 0.95)},
            {"params": ps[6:], "lr": 3e-4, "weight_decay": 0.0, "eps": 1e-6}]
pa = make(3); pd = [a.detach().clone().requires_grad_(True) for a in pa]
oa = optim.accelerate_optimizer(torch.optim.Adam(groups(pa), weight_decay=1e-5))
od = torch.optim.Adam(groups(pd), weight_decay=1e-5)
g = torch.Generator(device=DEV).manual_seed(4)
def cmp(tag, xs, ys):
    print(tag, ["%.2e" % float((x.detach() - y.detach()).abs().max()) for x, y in zip(xs, ys)])
for it in range(3):
    for a, d in zip(pa, pd):
        gr = torch.randn(a.shape, device=DEV, generator=g); a.grad, d.grad = gr.clone(), gr.clone()
    oa.step(); od.step()
    cmp(f"step{it+1} fused-vs-stock", pa, pd)
sd = copy.deepcopy(oa.state_dict())
sdd = copy.deepcopy(od.state_dict())
for k in ("exp_avg", "exp_avg_sq"):
    print(k, ["%.2e" % float((sd["state"][i][k] - sdd["state"][i][k]).abs().max()) for i in range(len(SHAPES))])
print("steps", [float(sd["state"][i]["step"]) for i in range(len(SHAPES))], [float(sdd["state"][i]["step"]) for i in range(len(SHAPES))])
pb = [a.detach().clone().requires_grad_(True) for a in pa]
ob = torch.optim.Adam(groups(pb), weight_decay=1e-5); ob.load_state_dict(sd)
for it in range(2):
    for a, b, d in zip(pa, pb, pd):
        gr = torch.randn(a.shape, device=DEV, generator=g); a.grad, b.grad, d.grad = gr.clone(), gr.clone(), gr.clone()
    oa.step(); ob.step(); od.step()
    cmp(f"step{it+4} fused-vs-stock", pa, pd)
    cmp(f"step{it+4} fused-vs-resumed", pa, pb)
    cmp(f"step{it+4} stock-vs-resumed", pd, pb)
