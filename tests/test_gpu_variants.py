"""GPU parity of the remaining model variants against goldens produced by the reference's own
classes (tools/make_golden.py gen_variants): TanhAttention, AggregationModel / AggregationProjectModel around it
(5_JointFusion/models.py:22-88) and the 1- / 4-channel ResNet-50 stems (resnet.py:167-337, 375-428).
Tolerances: attention weights / pooled features go through one bf16 tcgen05 GEMM (fp32 accumulation): 2e-2 relative
L2 like the MLP tests; ResNet features 1e-2 relative L2 (north_star)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import det_input

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "variants_reference.npz")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


class _FakeResnet(nn.Module):
    def forward_extract(self, p):
        return p.flatten(1)


def _attention(g, dev):
    from multimodalbrainsurvival_b200 import models
    att = models.TanhAttention(g["att_vector"].shape[0])
    with torch.no_grad():
        att.vector.copy_(torch.tensor(g["att_vector"]))
        att.linear.weight.copy_(torch.tensor(g["att_linear"]))
    return att.to(dev).eval()


def test_tanh_attention_matches_reference():
    from multimodalbrainsurvival_b200 import _lib
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    att = _attention(g, dev)
    x = torch.tensor(g["att_x"], device=dev)
    before = _lib.lib().mmbs_launch_count()
    with torch.no_grad():
        out, w = att(x)
        pooled, w2 = att.pooled(x)
    assert _lib.lib().mmbs_launch_count() > before, "TanhAttention did not run the kernels"
    assert tuple(out.shape) == g["att_out"].shape and tuple(w.shape) == g["att_weights"].shape
    assert _rel(w.cpu().numpy(), g["att_weights"]) < 2e-2
    assert _rel(out.cpu().numpy(), g["att_out"]) < 2e-2
    assert _rel(pooled.cpu().numpy(), g["att_out"].mean(axis=1)) < 2e-2
    assert torch.equal(w, w2)
    assert abs(float(w.sum(dim=1).mean()) - 1.0) < 1e-5
    # autograd active on the parameters: the module graph (no kernel backward for the aggregator), same numbers
    out_t, w_t = att(x)
    assert out_t.requires_grad and _rel(out_t.detach().cpu().numpy(), g["att_out"]) < 1e-4


def test_aggregation_models_with_attention_match_reference():
    from multimodalbrainsurvival_b200 import models
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    dim = g["att_vector"].shape[0]
    bag = torch.tensor(g["att_x"], device=dev).view(3, 5, 1, 16, 16)
    agg = models.AggregationModel(_FakeResnet(), _attention(g, dev), dim, resnet_dim=dim)
    proj = models.AggregationProjectModel(_FakeResnet(), _attention(g, dev), dim, resnet_dim=dim,
                                          hdim=g["proj_b"].shape[0])
    with torch.no_grad():
        agg.fc.weight.copy_(torch.tensor(g["agg_fc_w"]))
        agg.fc.bias.copy_(torch.tensor(g["agg_fc_b"]))
        proj.project.weight.copy_(torch.tensor(g["proj_w"]))
        proj.project.bias.copy_(torch.tensor(g["proj_b"]))
        proj.fc.weight.copy_(torch.tensor(g["proj_fc_w"]))
        proj.fc.bias.copy_(torch.tensor(g["proj_fc_b"]))
    agg, proj = agg.to(dev).eval(), proj.to(dev).eval()
    with torch.no_grad():
        y_agg, w = agg(bag)
        f_agg, _ = agg.extract(bag)
        y_proj, _ = proj(bag)
        f_proj, _ = proj.extract(bag)
    assert tuple(w.shape) == (3, 5, 1)
    assert _rel(f_agg.cpu().numpy(), g["agg_feat"]) < 2e-2
    assert _rel(f_proj.cpu().numpy(), g["proj_feat"]) < 2e-2
    # scalar heads: absolute error against the scale of the features that feed them
    assert np.abs(y_agg.cpu().numpy() - g["agg_out"]).max() < 2e-2 * max(1.0, np.abs(g["agg_out"]).max())
    assert np.abs(y_proj.cpu().numpy() - g["proj_out"]).max() < 2e-2 * max(1.0, np.abs(g["proj_out"]).max())
    # state_dict keys are the reference's (the one-layer Sequentials of the kernel path are not registered)
    assert sorted(proj.state_dict().keys()) == sorted(
        ["aggregator.vector", "aggregator.linear.weight", "project.weight", "project.bias", "fc.weight", "fc.bias"])


@pytest.mark.parametrize("channels", [4, 1])
def test_one_and_four_channel_stems_match_reference(channels):
    """RNfour / RNone run the same eval engine (generic-channel space-to-depth pack); golden from the reference's
    resnet50_4channel / resnet50_1channel with a seeded state_dict."""
    from multimodalbrainsurvival_b200 import resnet
    from oracle import resnet_oracle
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    net = (resnet.resnet50_4channel if channels == 4 else resnet.resnet50_1channel)()
    sd = net.state_dict()
    sd.update({k: v for k, v in resnet_oracle.init_state_dict(seed=77).items() if k != "conv1.weight"})
    gen = torch.Generator().manual_seed(100 + channels)
    sd["conv1.weight"] = torch.randn(64, channels, 7, 7, generator=gen) * (2.0 / (49 * 64)) ** 0.5
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    x = torch.tensor(det_input((2, channels, 224, 224), a=0.013 * channels), device=dev)
    with torch.no_grad():
        f = net.forward_extract(x)
    assert net._engines, "the eval engine did not run"
    ref = g["stem%d_feat" % channels]
    assert _rel(f.cpu().numpy(), ref) < 1e-2, _rel(f.cpu().numpy(), ref)


def test_adapt_pretrained_stem_follows_the_reference_surgery():
    from multimodalbrainsurvival_b200 import resnet
    from oracle import resnet_oracle
    sd = resnet_oracle.init_state_dict(seed=5)
    four = resnet.adapt_pretrained_stem(resnet.resnet50_4channel(), sd)
    one = resnet.adapt_pretrained_stem(resnet.resnet50_1channel(), sd)
    assert torch.equal(four.conv1.weight[:, :3], sd["conv1.weight"])
    assert float(four.conv1.weight[:, 3].detach().std()) < 0.01
    assert torch.allclose(one.conv1.weight, sd["conv1.weight"].mean(dim=1, keepdim=True))
    assert torch.equal(one.layer4[2].conv3.weight, sd["layer4.2.conv3.weight"])
