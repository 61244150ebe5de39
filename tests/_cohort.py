"""Synthetic survival cohort + the reference-semantics (fp32, CPU) fine-tuning loop used by the
c-index parity tests (north_star: "the downstream c-index must agree within 0.005").

Reference semantics restated here (all through ``oracle/``, never the product path):
  * one fine-tuning step = ``model.train()`` forward of ``AggregationModel`` (batch-statistics BatchNorm in every
    layer), ``CoxLoss`` on the batch, backward through ``fc`` + ``layer4``, ``Adam(lr, weight_decay)``
    (/root/reference/1_HistoPathology/2_HistoPath_train.py:296-337, :541-558);
  * scoring = eval-mode forward over the cohort, per-case mean of the patch scores, Harrell's C of
    ``(survival_months, -score, vital_status)`` (/root/reference/1_HistoPathology/3_HistoPath_savescore.py:126-152).
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import cindex_oracle, cox_oracle, resnet_oracle


def make_cohort(n_cases: int, patches_per_case: int, seed: int):
    """Patches whose statistics differ per case (scale + colour offset), so that the patch scores of a random
    network spread over the cases instead of collapsing onto one value."""
    g = torch.Generator().manual_seed(seed)
    scale = 0.5 + 1.5 * torch.rand(n_cases, generator=g)
    off = 0.5 * torch.randn(n_cases, 3, 1, 1, generator=g)
    n = n_cases * patches_per_case
    x = torch.randn(n, 3, 224, 224, generator=g)
    x = x * scale.repeat_interleave(patches_per_case).view(-1, 1, 1, 1) + off.repeat_interleave(patches_per_case, 0)
    case = np.repeat(np.arange(n_cases), patches_per_case)
    return x, case


def survival_from_scores(case_scores: np.ndarray, seed: int):
    """Survival times anti-correlated with the risk score (c-index ~ 0.75) + 65 % observed events."""
    rng = np.random.default_rng(seed)
    z = (case_scores - case_scores.mean()) / (case_scores.std() + 1e-12)
    t = (np.exp(-0.8 * z + 0.6 * rng.standard_normal(z.shape[0])) * 30).astype(np.float32)
    e = (rng.uniform(size=z.shape[0]) < 0.65).astype(np.float32)
    return t, e


def case_mean(patch_scores: np.ndarray, case: np.ndarray) -> np.ndarray:
    n_cases = int(case.max()) + 1
    return np.array([patch_scores[case == c].astype(np.float32).mean() for c in range(n_cases)], dtype=np.float32)


def cindex(case_scores, t, e) -> float:
    return cindex_oracle.concordance_index(t, -np.asarray(case_scores, dtype=np.float64), e)


def oracle_scores(sd, fc_w, fc_b, x, emulate_bf16=False, chunk=32) -> np.ndarray:
    outs = []
    for i in range(0, x.shape[0], chunk):
        f = resnet_oracle.forward_extract(sd, x[i:i + chunk], emulate_bf16=emulate_bf16)
        outs.append((f @ fc_w.t() + fc_b).view(-1))
    return torch.cat(outs).numpy()


def trainable_names(sd):
    return sorted(k for k, v in sd.items() if k.startswith("layer4.") and v.dtype.is_floating_point
                  and not k.endswith(("running_mean", "running_var")))


def oracle_finetune(sd, fc_w, fc_b, batches, lr, weight_decay, emulate_bf16=False, ranks=1):
    """K fine-tuning steps with the reference's semantics on CPU.  ``batches`` = [(x, times, events), ...].
    ``ranks`` > 1 restates data parallelism WITHOUT synchronised BatchNorm: the batch is split into ``ranks`` equal
    slices that are normalised with their own batch statistics, the Cox risk set is the whole batch and the
    parameter gradients are summed (used to bound the documented per-rank-BN deviation).
    Returns (updated state dict, fc_w, fc_b, [loss per step])."""
    sd = {k: v.clone() for k, v in sd.items()}
    fc_w, fc_b = fc_w.clone(), fc_b.clone()
    names = trainable_names(sd)
    params = [sd[k] for k in names] + [fc_w, fc_b]
    opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
    losses = []
    for x, t, e in batches:
        B = x.shape[0]
        per = B // ranks
        slices = [slice(r * per, (r + 1) * per) for r in range(ranks)]
        # forward of every slice (features only), then the global Cox gradient, then the backward of every slice
        feats = [resnet_oracle.train_step(sd, x[s], torch.zeros(per, 2048), emulate_bf16=emulate_bf16)[0] for s in slices]
        f_all = torch.cat(feats)
        scores = (f_all @ fc_w.t() + fc_b).view(-1)
        loss, _ = cox_oracle.cox_forward(scores.numpy(), t, e, np.float32)
        g = torch.tensor(cox_oracle.cox_backward(scores.numpy(), t, e), dtype=torch.float32)   # dL/dscores
        losses.append(float(loss))
        gf = g[:, None] * fc_w                      # dL/dfeatures
        grads = {k: torch.zeros_like(sd[k]) for k in names}
        stats = None
        for r, s in enumerate(slices):
            _, gr, st = resnet_oracle.train_step(sd, x[s], gf[s], emulate_bf16=emulate_bf16)
            for k in names:
                grads[k] += gr[k]
            if r == 0:
                stats = st                          # rank 0's running statistics are the ones that get saved
        for k in names:
            sd[k].grad = grads[k]
        fc_w.grad = g[None, :] @ f_all
        fc_b.grad = g.sum().view(1)
        opt.step()
        sd.update({k: v for k, v in stats.items()})
    for k in names:
        sd[k].grad = None
    return sd, fc_w.detach(), fc_b.detach(), losses
