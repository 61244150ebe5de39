import sys; sys.path.insert(0,'/root/repo')
import torch
from multimodalbrainsurvival_b200 import cox
dev='cuda'; n=10_000_000
g=torch.Generator(device=dev).manual_seed(1)
s=torch.randn(n,device=dev,generator=g).requires_grad_(True)
e=(torch.rand(n,device=dev,generator=g)<0.6).float()
for name,t in (('months', torch.floor(torch.rand(n,device=dev,generator=g)*240)), ('days', torch.floor(torch.rand(n,device=dev,generator=g)*6000)), ('uniform', torch.rand(n,device=dev,generator=g)*200)):
    ts=[]
    for i in range(6):
        s.grad=None
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); cox.cox_loss(s,t,e).backward(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, 'fwd+bwd ms', sorted(ts)[2], cox.pipeline_state(t))
