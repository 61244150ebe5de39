"""Oracle for the ResNet-50 patch encoder.  TEST INFRASTRUCTURE ONLY.

A functional fp32 torch restatement of ``ResNet.forward_extract``
(/root/reference/5_JointFusion/resnet.py:151-165; identical at
1_HistoPathology/resnet.py:151-165) with ``Bottleneck.forward`` (resnet.py:70-90)
and the layer plan of ``ResNet.__init__/_make_layer`` (resnet.py:94-132,
[3,4,6,3] blocks, stride on the 3x3 conv, resnet.py:61).  It walks a plain
``state_dict`` with the reference's key names instead of
owning modules, so the same weights can be pushed through the reference module,
this oracle and the CUDA path.

Modes:
  * eval BN (running stats, eps 1e-5)           - resnet.py:60,63,66,99,123 in eval()
  * ``emulate_bf16=True`` rounds the input, the weights and every stored
    activation to bf16 exactly where the CUDA path stores bf16 (fp32 accumulate,
    fp32 BN affine) - a tight checker for the kernels.

Pinned by tests/golden/resnet_*.npz generated from the reference module by
tools/make_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LAYER_PLAN = ((1, 64, 3, 1), (2, 128, 4, 2), (3, 256, 6, 2), (4, 512, 3, 2))


def init_state_dict(seed: int = 1111, randomize_bn: bool = True, num_classes: int = 1000,
                    bn3_gamma_scale: float = 1.0):
    """Seeded ResNet-50 state_dict with the reference's key names and He-normal
    conv init (resnet.py:109-115).  BN running stats are randomised so that the
    folded scale/shift is exercised.

    ``bn3_gamma_scale`` < 1 shrinks the gamma of every block's last BatchNorm (what trained ResNets look like:
    the residual branch is a small correction).  The training-mode parity cases use 0.1: with gamma ~ 1 a
    randomly initialised ResNet-50 under batch-statistics BatchNorm amplifies any rounding by ~1.3x per block
    (measured with ``train_step(emulate_bf16=True)``: 10 % feature error from bf16 storage alone), which tests
    the conditioning of the synthetic network rather than the kernels."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k):
        n = k * k * cout
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / n)

    def bn(name, c):
        if randomize_bn:
            sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
            sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(c)
            sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", 64, 3, 7)
    bn("bn1", 64)
    inplanes = 64
    for li, planes, blocks, stride in LAYER_PLAN:
        for b in range(blocks):
            p = f"layer{li}.{b}"
            conv(p + ".conv1", planes, inplanes, 1)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inplanes, 1)
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    if bn3_gamma_scale != 1.0:
        for k in sd:
            if k.endswith("bn3.weight"):
                sd[k] = sd[k] * bn3_gamma_scale
    bound = 1.0 / math.sqrt(2048)
    sd["fc.weight"] = (torch.rand(num_classes, 2048, generator=g) * 2 - 1) * bound
    sd["fc.bias"] = (torch.rand(num_classes, generator=g) * 2 - 1) * bound
    return sd


def _r(x, emulate):
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


def _bn(x, sd, name):
    scale = sd[name + ".weight"] / torch.sqrt(sd[name + ".running_var"] + BN_EPS)
    shift = sd[name + ".bias"] - sd[name + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


@torch.no_grad()
def forward_extract(sd, x, emulate_bf16: bool = False, prefix: str = ""):
    """x: (B,3,224,224) fp32 NCHW -> (B,2048) fp32."""
    e = emulate_bf16
    w = (lambda k: _r(sd[prefix + k + ".weight"].float(), e))
    sdp = {k[len(prefix):]: v.float() for k, v in sd.items() if k.startswith(prefix) and v.dtype.is_floating_point}
    x = _r(x.float(), e)
    x = F.conv2d(x, w("conv1"), stride=2, padding=3)
    x = _r(F.relu(_bn(x, sdp, "bn1")), e)
    x = F.max_pool2d(x, 3, 2, 1)
    for li, planes, blocks, stride in LAYER_PLAN:
        for b in range(blocks):
            p = f"layer{li}.{b}"
            s = stride if b == 0 else 1
            out = F.conv2d(x, w(p + ".conv1"))
            out = _r(F.relu(_bn(out, sdp, p + ".bn1")), e)
            out = F.conv2d(out, w(p + ".conv2"), stride=s, padding=1)
            out = _r(F.relu(_bn(out, sdp, p + ".bn2")), e)
            out = F.conv2d(out, w(p + ".conv3"))
            out = _bn(out, sdp, p + ".bn3")
            if b == 0:
                res = F.conv2d(x, w(p + ".downsample.0"), stride=s)
                res = _r(_bn(res, sdp, p + ".downsample.1"), e)
            else:
                res = x
            last = (li == 4 and b == blocks - 1)
            x = F.relu(out + res)
            if not last:
                x = _r(x, e)  # the final block's output is pooled in fp32
    return F.avg_pool2d(x, 7, 1).flatten(1)


def train_step(sd, x, grad_features, momentum: float = 0.1, prefix: str = "", emulate_bf16: bool = False):
    """Training-mode restatement (``model.train()``): every BatchNorm2d normalises with the batch mean and
    biased variance and updates running_mean / running_var (unbiased) with ``momentum``
    (nn.BatchNorm2d defaults used by resnet.py:60-66,99); autograd through layer4 only (the reference's
    fine-tuning set, /root/reference/1_HistoPathology/2_HistoPath_train.py:541-551, n_layers_to_train = 2).

    x (B,3,224,224) fp32, grad_features (B,2048) = dLoss/dfeatures, or a callable ``f(features) -> dLoss/dfeatures``
    evaluated on the detached features of this very forward pass (a head + loss on top of the trunk).
    ``emulate_bf16`` rounds the input, the weights, every raw convolution output (the statistics are taken
    from the rounded values) and every activation to bf16, where the CUDA path stores bf16.
    Returns (features, {layer4 parameter name: gradient}, {bn buffer name: updated value})."""
    e = emulate_bf16
    sdp = {k[len(prefix):]: v.detach().clone().float() for k, v in sd.items()
           if k.startswith(prefix) and v.dtype.is_floating_point}
    leaves = {k: v.requires_grad_(True) for k, v in sdp.items()
              if k.startswith("layer4.") and not k.endswith(("running_mean", "running_var"))}
    new_stats = {}

    def bn(t, name):
        t = _r(t, e)
        rm, rv = sdp[name + ".running_mean"].clone(), sdp[name + ".running_var"].clone()
        y = F.batch_norm(t, rm, rv, sdp[name + ".weight"], sdp[name + ".bias"], True, momentum, BN_EPS)
        new_stats[name + ".running_mean"], new_stats[name + ".running_var"] = rm, rv
        return y

    def w(k):
        t = sdp[k + ".weight"]
        return _r(t, e) if t.dim() == 4 else t

    t = F.conv2d(_r(x.float(), e), w("conv1"), stride=2, padding=3)
    t = F.max_pool2d(_r(F.relu(bn(t, "bn1")), e), 3, 2, 1)
    for li, planes, blocks, stride in LAYER_PLAN:
        for b in range(blocks):
            p = f"layer{li}.{b}"
            s = stride if b == 0 else 1
            out = _r(F.relu(bn(F.conv2d(t, w(p + ".conv1")), p + ".bn1")), e)
            out = _r(F.relu(bn(F.conv2d(out, w(p + ".conv2"), stride=s, padding=1), p + ".bn2")), e)
            out = bn(F.conv2d(out, w(p + ".conv3")), p + ".bn3")
            res = bn(F.conv2d(t, w(p + ".downsample.0"), stride=s), p + ".downsample.1") if b == 0 else t
            t = _r(F.relu(out + res), e)
    feats = F.avg_pool2d(t, 7, 1).flatten(1)
    names = sorted(leaves)
    if callable(grad_features):
        grad_features = grad_features(feats.detach())
    grads = torch.autograd.grad(feats, [leaves[k] for k in names], grad_outputs=grad_features.float())
    return feats.detach(), dict(zip(names, grads)), new_stats
