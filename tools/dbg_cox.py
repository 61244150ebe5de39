import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from multimodalbrainsurvival_b200 import cox
n, distinct = 5000, 7
rng = np.random.default_rng(n + distinct)
s = (rng.standard_normal(n) * 2).astype(np.float32)
t = rng.integers(0, distinct, n).astype(np.float32)
e = (rng.uniform(size=n) < 0.5).astype(np.float32)
dev = "cuda:0"
sc = torch.tensor(s, device=dev, requires_grad=True)
tt = torch.tensor(t, device=dev); ee = torch.tensor(e, device=dev)
try:
    loss = cox.cox_loss(sc, tt, ee); torch.cuda.synchronize(); print("fwd ok", float(loss))
    loss.backward(); torch.cuda.synchronize(); print("bwd ok")
    p = cox.risk_order(tt); torch.cuda.synchronize(); print("order ok")
except Exception as ex:
    print("EXC", ex)
