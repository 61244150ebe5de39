"""Drop-in superset of the reference's four ``models.py`` files (same names, signatures,
return conventions and ``state_dict`` keys):

  1_HistoPathology/models.py : Identity :13, TanhAttention :22, AggregationModel :35-57,
                               AggregationProjectModel :59, cox_loss :90, CoxLoss :113,
                               NLLSurvLoss :121, nll_loss :155, PatchBagDataset :234
  5_JointFusion/models.py    : BagHistopathologyRNAModel :87-104, HistopathologyRNAModel :106
  2_GeneExpression/models.py = 3_EarlyFusion/models.py : RNAOnlyModel :8-21, cox_loss :24

The wrappers are thin: the arithmetic lives in the kernels reached through
``resnet.forward_extract`` (engine.ResNetEngine), ``mlp.run_mlp`` (tcgen05 GEMMs reading the
stock ``nn.Linear`` parameters in place) and ``cox.cox_loss``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .cox import CoxLoss, cox_loss  # noqa: F401  (re-exported under the reference's names)
from .datasets import PatchBagDataset  # noqa: F401
from . import mlp as _mlp

__all__ = ["Identity", "TanhAttention", "AggregationModel", "AggregationProjectModel", "cox_loss", "CoxLoss",
           "NLLSurvLoss", "nll_loss", "PatchBagDataset", "BagHistopathologyRNAModel", "HistopathologyRNAModel",
           "RNAOnlyModel", "accelerate"]


class Identity(nn.Module):
    """Aggregator that keeps every patch feature; attention weights are all ones."""

    def forward(self, x):
        return x, torch.ones(x.shape[0], x.shape[1], device=x.device)


def _inference(module, *tensors):
    """CUDA fp32 tensors and nothing to differentiate: the fused kernels may run (they have no backward)."""
    import os
    if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1":
        return False
    if not all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 for t in tensors):
        return False
    if torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or
                                    any(p.requires_grad for p in module.parameters())):
        return False
    return True


def _linear(owner, name, x2d):
    """``getattr(owner, name)(x2d)`` for an nn.Linear child through the tcgen05 linear plan (mlp._MLPEngine reads the
    stock parameters in place).  The one-layer Sequential lives in __dict__ only: it must not appear in state_dict."""
    key = "_mmbs_seq_" + name
    seq = owner.__dict__.get(key)
    if seq is None or seq[0] is not getattr(owner, name):
        seq = nn.Sequential(getattr(owner, name))
        owner.__dict__[key] = seq
    seq.train(False)
    return _mlp.run_mlp(seq, x2d)


class TanhAttention(nn.Module):
    def __init__(self, dim=2048):
        super().__init__()
        self.dim = dim
        self.vector = torch.nn.Parameter(torch.zeros(dim))
        self.linear = nn.Linear(dim, dim, bias=False)

    def _fused(self, x, want_out):
        """csrc/attention.cu behind the linear layer's GEMM: -> (out | None, pooled, attention_weights)."""
        from . import _lib
        b, bag, dim = x.shape
        x = x.contiguous()
        h = _linear(self, "linear", x.view(b * bag, dim))
        attn = torch.empty((b, bag, 1), dtype=torch.float32, device=x.device)
        out = torch.empty_like(x) if want_out else None
        pooled = torch.empty((b, dim), dtype=torch.float32, device=x.device)
        vec = self.vector.detach().float().contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mmbs_attention_pool(_lib.ptr(x), _lib.ptr(h), _lib.ptr(vec), b, bag, dim, _lib.ptr(attn),
                                                      _lib.ptr(out), _lib.ptr(pooled), _lib.stream_ptr()),
                       "mmbs_attention_pool")
        return out, pooled, attn

    def _fusable(self, x):
        return (x.dim() == 3 and x.shape[2] == self.dim and 1 <= x.shape[1] <= 8192 and x.shape[0] >= 1
                and _inference(self, x))

    def forward(self, x):
        if self._fusable(x):
            out, _, attention_weights = self._fused(x, True)
            return out, attention_weights
        logits = torch.tanh(self.linear(x)).matmul(self.vector.unsqueeze(-1))
        attention_weights = F.softmax(logits, dim=1)
        return x * attention_weights * x.shape[1], attention_weights

    def pooled(self, x):
        """(mean over the bag of forward(x)[0], attention_weights) without the per-patch intermediate."""
        if self._fusable(x):
            _, pooled, attention_weights = self._fused(x, False)
            return pooled, attention_weights
        out, attention_weights = self.forward(x)
        return out.mean(dim=1), attention_weights


def _aggregate(aggregator, features):
    """aggregator(features) followed by the mean over the bag (reference AggregationModel.extract :53-56)."""
    if isinstance(aggregator, TanhAttention):
        return aggregator.pooled(features)
    features, attention_weights = aggregator(features)
    return features.mean(dim=1), attention_weights


def _head(owner, name, x):
    """An nn.Linear head on (B, features): the linear plan when nothing is differentiated, the module otherwise."""
    if x.dim() == 2 and _inference(owner, x):
        return _linear(owner, name, x)
    return getattr(owner, name)(x)


def _bag_features(resnet, x, resnet_dim):
    """(B, bag, C, H, W) -> per-patch features (B, bag, resnet_dim) through forward_extract."""
    batch_size, bag_size = x.shape[0], x.shape[1]
    feats = resnet.forward_extract(x.reshape(-1, *x.shape[2:]))
    return feats.view(batch_size, bag_size, resnet_dim)


class AggregationModel(nn.Module):
    def __init__(self, resnet, aggregator, aggregator_dim, resnet_dim=2048, out_features=1):
        super().__init__()
        self.resnet = resnet
        self.aggregator = aggregator
        self.fc = nn.Linear(aggregator_dim, out_features)
        self.aggregator_dim = aggregator_dim
        self.resnet_dim = resnet_dim

    def extract(self, x):
        return _aggregate(self.aggregator, _bag_features(self.resnet, x, self.resnet_dim))

    def forward(self, x):
        features, attention_weights = self.extract(x)
        return _head(self, "fc", features), attention_weights


class AggregationProjectModel(nn.Module):
    def __init__(self, resnet, aggregator, aggregator_dim, resnet_dim=2048, out_features=1, hdim=200, dropout=.3):
        super().__init__()
        self.resnet = resnet
        self.aggregator = aggregator
        self.aggregator_dim = aggregator_dim
        self.resnet_dim = resnet_dim
        self.hdim = hdim
        self.dropout = nn.Dropout(p=dropout)
        self.project = nn.Linear(aggregator_dim, hdim)
        self.fc = nn.Linear(hdim, out_features)

    def extract(self, x):
        features, attention_weights = _aggregate(self.aggregator, _bag_features(self.resnet, x, self.resnet_dim))
        if features.dim() == 2 and not self.training and _inference(self, features):
            from . import _lib
            features = _linear(self, "project", features)   # tcgen05 GEMM + bias; F.tanh in place; eval dropout = identity
            with torch.cuda.device(features.device):
                _lib.check(_lib.lib().mmbs_tanh_inplace_f32(_lib.ptr(features), features.numel(), _lib.stream_ptr()),
                           "mmbs_tanh_inplace_f32")
            return features, attention_weights
        features = self.dropout(torch.tanh(self.project(features)))
        return features, attention_weights

    def forward(self, x):
        features, attention_weights = self.extract(x)
        return _head(self, "fc", features), attention_weights


class BagHistopathologyRNAModel(nn.Module):
    def __init__(self, resnet, rna_mlp, final_mlp):
        super().__init__()
        self.resnet = resnet
        self.rna_mlp = rna_mlp
        self.final_mlp = final_mlp

    def forward(self, patch_bag, rna, resnet_dim=2048):
        image_features = _bag_features(self.resnet, patch_bag, resnet_dim).mean(dim=1)
        rna_features = _mlp.run_mlp(self.rna_mlp, rna)
        return _mlp.run_mlp(self.final_mlp, torch.cat([image_features, rna_features], dim=1))


class HistopathologyRNAModel(nn.Module):
    def __init__(self, resnet, rna_mlp, final_mlp):
        super().__init__()
        self.resnet = resnet
        self.rna_mlp = rna_mlp
        self.final_mlp = final_mlp

    def forward(self, patch, rna):
        image_features = self.resnet.forward_extract(patch)
        rna_features = _mlp.run_mlp(self.rna_mlp, rna)
        return _mlp.run_mlp(self.final_mlp, torch.cat([image_features, rna_features], dim=1))


class RNAOnlyModel(nn.Module):
    def __init__(self, rna_mlp, final_mlp):
        super().__init__()
        self.rna_mlp = rna_mlp
        self.final_mlp = final_mlp

    def forward(self, rna):
        return _mlp.run_mlp(self.final_mlp, _mlp.run_mlp(self.rna_mlp, rna))

    def extract(self, rna):
        return _mlp.run_mlp(self.rna_mlp, rna)


def accelerate(module: nn.Module) -> nn.Module:
    """Route a bare ``nn.Sequential`` MLP (the early-fusion script uses one as its whole
    model, 3_EarlyFusion/2_EarlyFusion_train.py:242-253) through the fused kernels.
    Parameters, ``state_dict`` keys and optimizer bindings are untouched."""
    return _mlp.AcceleratedSequential.wrap(module)


# --------------------------------------------------------------------------- NLL survival loss
def _nll_loss_torch(h, y, c, alpha, eps, reduction):
    """The reference's formula evaluated by torch (CPU tensors / dtypes the kernel does not take)."""
    n = len(y)
    y = y.view(n, 1)
    c = c.view(n, 1).float()
    hazards = torch.sigmoid(h)
    surv = torch.cat([torch.ones_like(c), torch.cumprod(1 - hazards, dim=1)], 1)  # S(-1) = 1
    log_s_prev = torch.log(torch.gather(surv, 1, y).clamp(min=eps))
    log_h_this = torch.log(torch.gather(hazards, 1, y).clamp(min=eps))
    log_s_this = torch.log(torch.gather(surv, 1, y + 1).clamp(min=eps))
    loss = (1 - alpha) * (-c * log_s_this) - (1 - c) * (log_s_prev + log_h_this)
    return loss.mean() if reduction == 'mean' else loss.sum()


class _NLLSurvFn(torch.autograd.Function):
    """csrc/nll.cu: one pass writes the loss and d loss_i / d h; backward only scales."""

    @staticmethod
    def forward(ctx, h, y, c, alpha, eps, mean):
        from . import _lib
        n, k = h.shape
        hc = h.detach().contiguous()
        yc = y.detach().reshape(-1).to(torch.int64).contiguous()
        cc = c.detach().reshape(-1).float().contiguous()
        buf = torch.empty(n * (k + 1) + 2, dtype=torch.float32, device=h.device)
        loss_i, grad_unit, out = buf[:n], buf[n:n * (k + 1)], buf[n * (k + 1):]
        with torch.cuda.device(h.device):
            _lib.check(_lib.lib().mmbs_nll_surv_forward(_lib.ptr(hc), _lib.ptr(yc), _lib.ptr(cc), n, k, float(alpha),
                                                        float(eps), int(mean), _lib.ptr(loss_i), _lib.ptr(grad_unit),
                                                        _lib.ptr(out), _lib.ptr(out[1:]), _lib.stream_ptr()),
                       "mmbs_nll_surv_forward")
        ctx.save_for_backward(grad_unit)
        ctx.shape, ctx.mean = (n, k), bool(mean)
        return out[0].clone().reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        from . import _lib
        (grad_unit,) = ctx.saved_tensors
        n, k = ctx.shape
        g = grad_loss.detach().reshape(1).float().contiguous()
        grad_h = torch.empty((n, k), dtype=torch.float32, device=grad_unit.device)
        with torch.cuda.device(grad_unit.device):
            _lib.check(_lib.lib().mmbs_nll_surv_backward(_lib.ptr(grad_unit), _lib.ptr(g), n, k, int(ctx.mean),
                                                         _lib.ptr(grad_h), _lib.stream_ptr()), "mmbs_nll_surv_backward")
        return grad_h, None, None, None, None, None


def nll_loss(h, y, c, alpha=0.0, eps=1e-7, reduction='mean'):
    """Discrete-time survival negative log-likelihood (Zadeh & Schmid 2020), the ``survival_bin`` task of the
    histopathology scripts (/root/reference/1_HistoPathology/models.py:155-232).
    h: (n, n_bins) logits; y: (n, 1) int64 bin; c: (n, 1) censoring indicator (1 = censored).
    CUDA fp32 logits run csrc/nll.cu (forward + gradient in one pass); anything else evaluates the same formula in torch."""
    if reduction not in ('mean', 'sum'):
        raise ValueError("Bad input for reduction: {}".format(reduction))
    if h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and h.shape[0] >= 1 and len(y) == h.shape[0]:
        return _NLLSurvFn.apply(h, y, c, alpha, eps, reduction == 'mean')
    return _nll_loss_torch(h, y, c, alpha, eps, reduction)


class NLLSurvLoss(nn.Module):
    def __init__(self, alpha=0.0, eps=1e-7, reduction='mean'):
        super().__init__()
        self.alpha = alpha
        self.eps = eps
        self.reduction = reduction

    def __call__(self, h, y, c):
        return nll_loss(h=h, y=y.unsqueeze(dim=1), c=c.unsqueeze(dim=1), alpha=self.alpha, eps=self.eps,
                        reduction=self.reduction)
