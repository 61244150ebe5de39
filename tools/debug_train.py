import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multimodalbrainsurvival_b200 import resnet, train_engine
from oracle import resnet_oracle

sd = resnet_oracle.init_state_dict(seed=31)
net = resnet.resnet50(pretrained=False); net.load_state_dict(sd); net = net.cuda().train()
for p in net.parameters(): p.requires_grad = False
torch.manual_seed(5)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
x = torch.randn(B, 3, 224, 224, device="cuda")
_orig = train_engine.ResNetTrainEngine._add_block
def _ab(self, blk, x, save):
    out = _orig(self, blk, x, True)
    return out
train_engine.ResNetTrainEngine._add_block = _ab
train_engine.ResNetTrainEngine._build_backward = lambda self: None
eng = train_engine.ResNetTrainEngine(net, B)
rm0 = {n: b.running_mean.clone() for n, b in net.named_modules() if isinstance(b, torch.nn.BatchNorm2d)}
f = eng.forward(x)
torch.cuda.synchronize()
def nchw(t): return t.permute(0, 3, 1, 2).float()
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
# stem
w = net.conv1.weight.to(torch.bfloat16).float()
raw_ref = F.conv2d(x.to(torch.bfloat16).float(), w, stride=2, padding=3)
raw0 = eng._keep[1]
print("stem raw", rel(nchw(raw0), raw_ref), raw0.shape)
st = eng.bns[0]
m_ref = nchw(raw0).mean((0, 2, 3)); v_ref = nchw(raw0).var((0, 2, 3), unbiased=False)
print("bn1 mean", rel(st.mean, m_ref), "invstd", rel(st.invstd, 1 / torch.sqrt(v_ref + 1e-5)))
print("stats sum", rel(st.stats[:64], nchw(raw0).sum((0,2,3))), "sq", rel(st.stats[64:128], (nchw(raw0)**2).sum((0,2,3))))
pool = eng._keep[2]
a0 = F.relu((nchw(raw0) - m_ref.view(1,-1,1,1)) / torch.sqrt(v_ref.view(1,-1,1,1) + 1e-5) * net.bn1.weight.view(1,-1,1,1) + net.bn1.bias.view(1,-1,1,1))
print("pool", rel(nchw(pool), F.max_pool2d(a0, 3, 2, 1)), pool.shape)
# per block (layer4 only keeps everything; check others through stats of their raw buffers)
for r in eng.l4:
    blk = r["blk"]
    print("block", tuple(r["x"].shape))
    xin = nchw(r["x"])
    def bnf(t, st):
        m = t.mean((0,2,3), keepdim=True); v = t.var((0,2,3), unbiased=False, keepdim=True)
        return (t - m) / torch.sqrt(v + 1e-5) * st.bn.weight.view(1,-1,1,1) + st.bn.bias.view(1,-1,1,1)
    bw = lambda c: c.weight.to(torch.bfloat16).float()
    raw1 = F.conv2d(xin, bw(blk.conv1)); print(" raw1", rel(nchw(r["raw1"]), raw1))
    a1 = F.relu(bnf(nchw(r["raw1"]), r["b1"])); print(" a1", rel(nchw(r["a1"]), a1))
    raw2 = F.conv2d(nchw(r["a1"]), bw(blk.conv2), stride=r["stride"], padding=1); print(" raw2", rel(nchw(r["raw2"]), raw2))
    a2 = F.relu(bnf(nchw(r["raw2"]), r["b2"])); print(" a2", rel(nchw(r["a2"]), a2))
    raw3 = F.conv2d(nchw(r["a2"]), bw(blk.conv3)); print(" raw3", rel(nchw(r["raw3"]), raw3))
    if r["rawd"] is not None:
        rawd = F.conv2d(xin, bw(blk.downsample[0]), stride=r["stride"]); print(" rawd", rel(nchw(r["rawd"]), rawd))
        res = bnf(nchw(r["rawd"]), r["bd"])
    else:
        res = xin
    out = F.relu(bnf(nchw(r["raw3"]), r["b3"]) + res); print(" out", rel(nchw(r["out"]), out))
