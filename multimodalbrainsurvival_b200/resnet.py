"""Drop-in for the reference's ``resnet.py`` (import surface, module tree and
``state_dict`` keys identical):

    from resnet import resnet50            /root/reference/1_HistoPathology/2_HistoPath_train.py:41
    model.forward(x), model.forward_extract(x), .conv1/.bn1/.layer1..4/.fc
        /root/reference/5_JointFusion/resnet.py:92-165 (ResNet), :54-90 (Bottleneck)

What differs is what runs underneath ``forward_extract``: for CUDA tensors in
eval mode without autograd the patches go through ``engine.ResNetEngine`` - NHWC
bf16 activations, tcgen05 implicit-GEMM convolutions with the eval BatchNorm,
ReLU and residual add fused into the epilogue (csrc/gemm_tcgen05.cu).  The
weights stay in the stock ``nn.Conv2d`` / ``nn.BatchNorm2d`` parameters; packed bf16
copies are caches rebuilt when a parameter's version counter moves.

Training mode (``model.train()``: batch-statistics BatchNorm in every layer, autograd through
``layer4``; the reference's fine-tuning configuration, 2_HistoPath_train.py:541-551) goes through
``train_engine.ResNetTrainEngine``: the same conv kernel with a statistics epilogue, BatchNorm
finalize/apply kernels and a layer4 backward made of dgrad/wgrad GEMMs.  Only configurations the
kernels do not cover (gradients into stem..layer3 or into the input, other ResNet depths, CPU tensors)
run the stock module graph below.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ['ResNet', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152']

model_urls = {
    'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth',
    'resnet34': 'https://download.pytorch.org/models/resnet34-333f7ec4.pth',
    'resnet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
    'resnet101': 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth',
    'resnet152': 'https://download.pytorch.org/models/resnet152-b121ed2d.pth',
}


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, 3, stride=stride, padding=1, bias=False)


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = conv3x3(inplanes, planes, stride), nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = conv3x3(planes, planes), nn.BatchNorm2d(planes)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + shortcut)


class Bottleneck(nn.Module):
    """1x1 -> 3x3 (carries the stride) -> 1x1 (x4), residual add, ReLU."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = _conv1x1(inplanes, planes), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = conv3x3(planes, planes, stride), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = _conv1x1(planes, planes * 4), nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + shortcut)


class _Trunk(nn.Module):
    """Shared body of ResNet / RNfour / RNone (they differ in the stem's input channels)."""
    _in_channels = 3

    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(self._in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():  # He-normal convs, unit BatchNorm (reference resnet.py:109-115)
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._engines = {}

    def _make_layer(self, block, planes, blocks, stride=1):
        width = planes * block.expansion
        downsample = None
        if stride != 1 or self.inplanes != width:
            downsample = nn.Sequential(_conv1x1(self.inplanes, width, stride), nn.BatchNorm2d(width))
        stages = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = width
        stages += [block(width, planes) for _ in range(1, blocks)]
        return nn.Sequential(*stages)

    # ---- module-graph path (training mode / autograd / non-accelerated variants)
    def _features_torch(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.avgpool(x)
        return x.view(x.size(0), -1)

    def forward(self, x):
        return self.fc(self.forward_extract(x))

    # uint8 inputs (raw pixels, 3-channel stems only) are normalised on the device with the reference's transform constants
    # (Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]), 4_HistoPath_extractfeatures.py:133-137)
    input_mean = (0.485, 0.456, 0.406)
    input_std = (0.229, 0.224, 0.225)

    def forward_extract(self, x):
        if self._can_accelerate_train(x):
            from . import train_engine
            with torch.cuda.device(x.device):   # the kernels launch on the CURRENT device's stream
                return train_engine.run_train(self, x, self._engines)
        if not x.is_cuda and os.environ.get("MMBS_DISABLE_KERNELS", "0") != "1":
            # north_star: no CPU fallback.  (MMBS_DISABLE_KERNELS=1 keeps the stock module graph reachable for
            # host-side tooling: checkpoint round-trips, oracle pinning.)
            raise RuntimeError("ResNet.forward_extract: input must be a CUDA tensor (this build has no CPU path; "
                               "move the model and the batch to the GPU)")
        if x.dtype == torch.uint8 and self._in_channels != 3:
            raise RuntimeError("uint8 pixels are only defined for the 3-channel stem (the reference normalises RGB patches)")
        if x.dtype == torch.uint8 and not self._can_accelerate(x):
            mean = torch.tensor(self.input_mean, device=x.device).view(1, 3, 1, 1)
            std = torch.tensor(self.input_std, device=x.device).view(1, 3, 1, 1)
            x = (x.float() / 255.0 - mean) / std
        if self._can_accelerate(x):
            return self._features_b200(x)
        return self._features_torch(x)

    def _can_accelerate(self, x):
        if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1":
            return False
        # every depth of the reference's resnet.py (18/34: BasicBlock, 50/101/152: Bottleneck; resnet.py:167-337)
        return (x.is_cuda and not self.training and not torch.is_grad_enabled()
                and x.dim() == 4 and tuple(x.shape[1:]) == (self._in_channels, 224, 224)
                and isinstance(self.layer1[0], (Bottleneck, BasicBlock))
                and self.fc.in_features == 512 * type(self.layer1[0]).expansion)

    def _can_accelerate_train(self, x):
        """model.train() on CUDA, fp32 (normalised) or uint8 (raw) 224x224 patches, ResNet-50, gradients (if any) confined
        to layer4."""
        if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1" or os.environ.get("MMBS_RESNET_TRAIN", "1") != "1":
            return False
        if self._in_channels != 3:   # the training engine is built for the RGB stem
            return False
        if not (self.training and x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.dim() == 4
                and tuple(x.shape[1:]) == (3, 224, 224) and isinstance(self.layer1[0], Bottleneck)
                and self.fc.in_features == 2048 and [len(l) for l in (self.layer1, self.layer2, self.layer3,
                                                                      self.layer4)] == [3, 4, 6, 3]):
            return False
        # a BatchNorm put back into eval mode by hand (frozen-statistics fine-tuning) must use its running statistics
        # (and exotic BatchNorm configurations - cumulative momentum, no affine / running statistics - stay on it too)
        if not all(m.training and m.momentum is not None and m.affine and m.track_running_stats
                   for m in self.modules() if isinstance(m, nn.BatchNorm2d)):
            return False
        # the BatchNorm / weight-pack kernels read parameters and buffers through raw float* pointers
        for t in list(self.parameters()) + [b for b in self.buffers() if b.dtype.is_floating_point]:
            if t.dtype != torch.float32 or t.device != x.device:
                return False
        if torch.is_grad_enabled():
            from . import train_engine
            if x.requires_grad or train_engine.trainable_outside_layer4(self):
                return False
        return True

    def _features_b200(self, x):
        with torch.cuda.device(x.device):       # the kernels launch on the CURRENT device's stream
            return self._features_b200_impl(x)

    def _features_b200_impl(self, x):
        from . import engine
        B = x.shape[0]
        x = x.contiguous() if x.dtype == torch.uint8 else x.float().contiguous()
        out = torch.empty((B, self.fc.in_features), dtype=torch.float32, device=x.device)
        done = 0
        while done < B:
            chunk = engine.default_chunk(B - done)
            key = ("eval", x.device.index, chunk)
            eng = self._engines.get(key)
            ver = engine.ResNetEngine.weights_version(self)
            if eng is None or eng._weights_version != ver:
                self._engines.pop(key, None)
                while sum(1 for k in self._engines if k[0] == "eval") >= 3:   # engines own GBs of buffers: keep a few
                    self._engines.pop(next(k for k in self._engines if k[0] == "eval"))
                eng = engine.ResNetEngine(self, chunk)
                self._engines[key] = eng
            eng.run_chunk(x[done:done + chunk], out[done:done + chunk], norm=(self.input_mean, self.input_std))
            done += chunk
        return out

    def __getstate__(self):  # engines hold ctypes handles: never pickle / deepcopy them
        d = self.__dict__.copy()
        d["_engines"] = {}
        return d


class ResNet(_Trunk):
    _in_channels = 3


class RNfour(_Trunk):
    """4-channel stem (reference resnet.py:167-250); eval-mode extraction runs the same kernels."""
    _in_channels = 4


class RNone(_Trunk):
    """1-channel stem (reference resnet.py:253-337)."""
    _in_channels = 1


def adapt_pretrained_stem(model, pretrained_dict):
    """Load 3-channel ImageNet weights into a 1- / 4-channel trunk the way the reference does
    (resnet50_4channel / resnet50_1channel, resnet.py:375-428): everything but conv1 as is; 4 channels: conv1 ~
    N(0, 0.001) with the RGB filters copied into the first three input channels; 1 channel: the mean of the RGB filters."""
    own = model.state_dict()
    own.update({k: v for k, v in pretrained_dict.items() if k != 'conv1.weight'})
    model.load_state_dict(own)
    w = pretrained_dict['conv1.weight']
    with torch.no_grad():
        if model._in_channels == 4:
            model.conv1.weight.normal_(0, 0.001)
            model.conv1.weight[:, :3] = w
        elif model._in_channels == 1:
            model.conv1.weight.copy_(w.mean(dim=1, keepdim=True))
        else:
            model.conv1.weight.copy_(w)
    return model


class ResNetProject(nn.Module):
    def __init__(self, resnet, hdim=200, input_dim=2048, dropout=.3):
        super().__init__()
        self.resnet = resnet
        self.hdim = hdim
        self.dropout = nn.Dropout(p=dropout)
        self.project = nn.Linear(input_dim, hdim)
        self.fc = nn.Linear(hdim, 1)

    def forward_extract(self, x):
        return self.dropout(torch.tanh(self.project(self.resnet.forward_extract(x))))

    def forward(self, x):
        return self.fc(self.forward_extract(x))


_CONFIGS = {
    'resnet18': (BasicBlock, [2, 2, 2, 2]), 'resnet34': (BasicBlock, [3, 4, 6, 3]),
    'resnet50': (Bottleneck, [3, 4, 6, 3]), 'resnet101': (Bottleneck, [3, 4, 23, 3]),
    'resnet152': (Bottleneck, [3, 8, 36, 3]),
}


def _build(name, pretrained, cls=ResNet, **kwargs):
    block, layers = _CONFIGS[name]
    model = cls(block, layers, **kwargs)
    if pretrained:
        import torch.utils.model_zoo as model_zoo
        state = model_zoo.load_url(model_urls[name])
        if cls is not ResNet:
            adapt_pretrained_stem(model, state)
        else:
            model.load_state_dict(state)
    return model


def randomize_batchnorm_(model, seed=1111, bn3_gamma_scale=1.0):
    """Synthetic stand-in for trained BatchNorm state (benchmarks / smoke runs have no checkpoint to load): seeded
    gamma in [0.5, 1.5], small beta, running_mean ~ N(0, 0.1), running_var in [0.5, 1.5], so that the folded
    scale/shift of the eval path is exercised; ``bn3_gamma_scale`` shrinks every block's last gamma (trained
    ResNets look like that; the training-mode cases use 0.1).  In place; returns the model."""
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            c = m.num_features
            with torch.no_grad():
                m.weight.copy_((0.5 + torch.rand(c, generator=g)) * (bn3_gamma_scale if name.endswith("bn3") else 1.0))
                m.bias.copy_(0.1 * torch.randn(c, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(c, generator=g))
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))
    return model


def resnet18(pretrained=False, **kwargs):
    return _build('resnet18', pretrained, **kwargs)


def resnet34(pretrained=False, **kwargs):
    return _build('resnet34', pretrained, **kwargs)


def resnet50(pretrained=False, **kwargs):
    """ResNet-50; pretrained=True downloads the ImageNet weights (needs network)."""
    return _build('resnet50', pretrained, **kwargs)


def resnet50_4channel(pretrained=False, **kwargs):
    return _build('resnet50', pretrained, cls=RNfour, **kwargs)


def resnet50_1channel(pretrained=False, **kwargs):
    return _build('resnet50', pretrained, cls=RNone, **kwargs)


def resnet101(pretrained=False, **kwargs):
    return _build('resnet101', pretrained, **kwargs)


def resnet152(pretrained=False, **kwargs):
    return _build('resnet152', pretrained, **kwargs)
