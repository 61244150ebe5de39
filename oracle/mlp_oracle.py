"""Oracle for the gene-expression / fusion MLPs.  TEST INFRASTRUCTURE ONLY.

Functional fp32 restatement (eval mode: Dropout = identity) of the
``nn.Sequential`` stacks the reference builds inline:

  RNA   : Dropout-Linear(12778,4096)-ReLU-Dropout-Linear(4096,2048) (+ Linear(2048,1))
          /root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257
  early : Dropout-Linear(4096,2048)-ReLU-Dropout-Linear(2048,200)-ReLU-Dropout-Linear(200,1)
          /root/reference/3_EarlyFusion/2_EarlyFusion_train.py:242-253
  joint : rna_mlp as RNA; head Dropout(0.8)-Linear(4096,1) on cat([img, rna])
          /root/reference/5_JointFusion/1_JointFusion_train.py:314-325,
          5_JointFusion/models.py:94-104

and of their gradients (plain autograd on the same functional graph).
``emulate_bf16`` rounds inputs/weights/hidden activations to bf16 where the
CUDA path stores bf16 (fp32 accumulation and bias).
Pinned by tests/golden/mlp_*.npz (tools/make_golden.py, reference modules).
"""
from __future__ import annotations

import torch


def _r(x, e):
    return x.to(torch.bfloat16).to(torch.float32) if e else x


def mlp_forward(x, layers, emulate_bf16=False):
    """layers: list of (weight (out,in), bias (out,), relu: bool).  Returns the
    list of every layer's output (post-activation)."""
    outs = []
    h = x.float()
    for (w, b, relu) in layers:
        h = _r(h, emulate_bf16) @ _r(w.float(), emulate_bf16).t() + b.float()
        if relu:
            h = torch.relu(h)
        outs.append(h)
    return outs


def rna_layers(sd, prefix="rna_mlp."):
    return [(sd[prefix + "1.weight"], sd[prefix + "1.bias"], True),
            (sd[prefix + "4.weight"], sd[prefix + "4.bias"], False)]


def early_layers(sd, prefix=""):
    return [(sd[prefix + "1.weight"], sd[prefix + "1.bias"], True),
            (sd[prefix + "4.weight"], sd[prefix + "4.bias"], True),
            (sd[prefix + "7.weight"], sd[prefix + "7.bias"], False)]
