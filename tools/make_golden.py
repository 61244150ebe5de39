#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code on CPU.

Runs only where /root/reference is mounted (the build container).  The GPU box
has no reference tree, so the vectors produced here are committed and the tests
read only them.  Nothing is copied out of the reference: its modules are
imported by path and called.

Stubs: lifelines / tensorboardX / matplotlib / sksurv are absent from the image
and only needed for the scripts' top-level imports.

    python tools/make_golden.py            # writes tests/golden/*.npz
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MMBS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub_third_party():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("lifelines")
    mod("lifelines.utils", concordance_index=lambda *a, **k: float("nan"))
    mod("sksurv")
    mod("sksurv.metrics", concordance_index_censored=lambda *a, **k: (float("nan"),))
    mod("tensorboardX", SummaryWriter=object)
    plt = mod("matplotlib.pyplot", switch_backend=lambda *a, **k: None)
    mod("matplotlib", pyplot=plt)


def load_ref(relpath, name):
    """Import a reference file by path with its own directory first on sys.path
    (the scripts import ``resnet`` / ``models`` by bare name)."""
    path = os.path.join(REF, relpath)
    d = os.path.dirname(path)
    for k in ("resnet", "models", "datasets"):
        sys.modules.pop(k, None)
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.path.remove(d)
    return m


def det_input(shape, a=0.37, b=1.3):
    """RNG-free deterministic tensor (reproducible on any box)."""
    n = int(np.prod(shape))
    i = np.arange(n, dtype=np.float64)
    return (np.sin(i * a) + 0.5 * np.cos(i * b * 0.01)).astype(np.float32).reshape(shape)


# ----------------------------------------------------------------------------- cox
def cox_cases():
    rng = np.random.default_rng(20261018)
    cases = {}
    cases["kat"] = (np.array([0.5, -1, 2, 0, 1.5, -0.5, 0.25, -2], np.float32),
                    np.array([5, 3, 5, 1, 3, 5, 0, 0], np.float32),
                    np.array([1, 0, 1, 1, 0, 1, 1, 0], np.float32))
    cases["n1_event"] = (np.array([0.3], np.float32), np.array([2.0], np.float32), np.array([1.0], np.float32))
    cases["n2"] = (np.array([0.3, -0.2], np.float32), np.array([2.0, 7.5], np.float32), np.array([1.0, 1.0], np.float32))
    for n in (22, 128, 1000, 4099):
        s = rng.standard_normal(n).astype(np.float32)
        t = rng.uniform(0, 200, n).astype(np.float32)
        e = (rng.uniform(size=n) < 0.6).astype(np.float32)
        cases[f"rand_{n}"] = (s, t, e)
    n = 1000
    cases["ties_1000"] = (rng.standard_normal(n).astype(np.float32) * 3,
                          rng.integers(0, 50, n).astype(np.float32),
                          (rng.uniform(size=n) < 0.6).astype(np.float32))
    cases["all_censored"] = (rng.standard_normal(64).astype(np.float32),
                             rng.uniform(0, 200, 64).astype(np.float32), np.zeros(64, np.float32))
    s = rng.standard_normal(300).astype(np.float32)
    s[[3, 77, 200]] = s.max() + 1.0  # tied maximum: max-backward splits evenly
    cases["tied_max"] = (s, rng.integers(0, 20, 300).astype(np.float32),
                         (rng.uniform(size=300) < 0.7).astype(np.float32))
    cases["wide_scores"] = (rng.standard_normal(512).astype(np.float32) * 30,
                            rng.uniform(0, 200, 512).astype(np.float32),
                            (rng.uniform(size=512) < 0.6).astype(np.float32))
    cases["neg_zero_times"] = (rng.standard_normal(16).astype(np.float32),
                               np.array([0.0, -0.0, 1, 0.0, -0.0, 2, 1, 0, 3, -0.0, 0, 1, 2, 3, 0, -0.0], np.float32),
                               np.ones(16, np.float32))
    return cases


def gen_cox():
    ref = load_ref("5_JointFusion/models.py", "ref_joint_models")
    real_sort = torch.sort
    out = {}
    for name, (s, t, e) in cox_cases().items():
        # the reference calls torch.sort(-times) (unstable); the only well-defined
        # tie order is the stable one (SURVEY.md §7 hard part 4) -> force stable.
        torch.sort = lambda x, *a, **k: real_sort(x, *a, stable=True, **k)
        try:
            sc = torch.tensor(s, requires_grad=True)
            loss = ref.cox_loss(sc, torch.tensor(t), torch.tensor(e))
            loss.backward()
            perm = real_sort(-torch.tensor(t), stable=True)[1].numpy()
        finally:
            torch.sort = real_sort
        out[name + "/scores"] = s
        out[name + "/times"] = t
        out[name + "/status"] = e
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/grad"] = sc.grad.numpy()
        out[name + "/perm"] = perm.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "cox_reference.npz"), **out)
    print("cox:", len(cox_cases()), "cases")


# ----------------------------------------------------------------------- aggregation
def gen_aggregate():
    _stub_third_party()
    rng = np.random.default_rng(7)
    ext = load_ref("1_HistoPathology/4_HistoPath_extractfeatures.py", "ref_extract")
    sav = load_ref("1_HistoPathology/3_HistoPath_savescore.py", "ref_savescore")

    n, d, n_case = 57, 96, 9
    cases = [f"TCGA-{rng.integers(0, n_case):02d}" for _ in range(n)]
    feats = rng.standard_normal((n, d)).astype(np.float32)

    class FakeModel:
        def eval(self):
            pass

        def extract(self, x):
            return x, None

    bs = 8
    loader = [{"patch_bag": torch.tensor(feats[i:i + bs]), "WSI": cases[i:i + bs], "case": cases[i:i + bs]}
              for i in range(0, n, bs)]
    case_uniques, features_final = ext.extract_features(FakeModel(), loader, torch.device("cpu"))

    outputs = rng.standard_normal((n, 1)).astype(np.float32)
    surv = rng.uniform(1, 100, n).astype(np.float32)
    vital = (rng.uniform(size=n) < 0.5).astype(np.float32)
    _, df = sav.get_survival_CI(outputs, cases, surv, vital)
    np.savez_compressed(
        os.path.join(OUT, "aggregate_reference.npz"),
        cases=np.array(cases), features=feats,
        case_uniques=np.array(case_uniques), features_final=np.asarray(features_final),
        outputs=outputs, survival=surv, vital=vital,
        ci_ids=np.array(df["id"]).astype(str), ci_score=np.array(df["score"]),
        ci_survival=np.array(df["survival_months"]), ci_vital=np.array(df["vital_status"]))
    print("aggregate: features_final", np.asarray(features_final).shape, np.asarray(features_final).dtype,
          "score dtype", np.array(df["score"]).dtype)


# ---------------------------------------------------------------------------- resnet
def gen_resnet():
    from oracle import resnet_oracle
    refnet = load_ref("5_JointFusion/resnet.py", "ref_resnet")
    sd = resnet_oracle.init_state_dict(seed=1111)
    net = refnet.resnet50(pretrained=False)
    missing = net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.tensor(det_input((2, 3, 224, 224)))
    with torch.no_grad():
        f = net.forward_extract(x)
    fp = {k: float(v.double().abs().sum()) for k, v in sd.items() if k in
          ("conv1.weight", "layer2.1.conv2.weight", "layer4.2.bn3.running_var", "fc.weight")}
    np.savez_compressed(os.path.join(OUT, "resnet_reference.npz"), features=f.numpy(),
                        fp_keys=np.array(list(fp.keys())), fp_vals=np.array(list(fp.values())))
    print("resnet: features", tuple(f.shape), "mean", float(f.mean()), missing)


# ---------------------------------------------------------------------- resnet, training mode
TRAIN_SAMPLE_STRIDE = 997   # conv-weight gradients are stored as every 997th element (the full set is 60 MB)


def gen_resnet_train():
    """One model.train() step of the reference ResNet-50 with fc + layer4 trainable
    (2_HistoPath_train.py:541-551 with n_layers_to_train = 2): features, layer4 gradients for a fixed
    dLoss/dfeatures, running statistics after the step."""
    from oracle import resnet_oracle
    refnet = load_ref("5_JointFusion/resnet.py", "ref_resnet_train")
    sd = resnet_oracle.init_state_dict(seed=2222, bn3_gamma_scale=0.1)   # see init_state_dict: conditioning
    net = refnet.resnet50(pretrained=False)
    net.load_state_dict(sd, strict=True)
    net.train()
    for p_ in net.parameters():
        p_.requires_grad = False
    for layer in [net.fc, net.layer4]:
        for p_ in layer.parameters():
            p_.requires_grad = True
    x = torch.tensor(det_input((4, 3, 224, 224), a=0.7))
    gw = torch.tensor(det_input((4, 2048), a=1.3))
    f = net.forward_extract(x)
    (f * gw).sum().backward()
    out = {"features": f.detach().numpy()}
    for name, p_ in net.layer4.named_parameters():
        g = p_.grad.detach().flatten()
        out["grad/layer4." + name] = (g if g.numel() <= 4096 else g[::TRAIN_SAMPLE_STRIDE]).numpy()
        out["gnorm/layer4." + name] = np.array(float(p_.grad.double().norm()))
    for name in ("bn1", "layer1.0.bn1", "layer2.0.downsample.1", "layer3.5.bn3", "layer4.0.bn2", "layer4.2.bn3"):
        bn = net.get_submodule(name)
        out["stat/" + name + ".running_mean"] = bn.running_mean.numpy()
        out["stat/" + name + ".running_var"] = bn.running_var.numpy()
        out["stat/" + name + ".num_batches_tracked"] = bn.num_batches_tracked.numpy()
    # the oracle restatement must agree with the reference module to fp32 round-off
    fo, go, so = resnet_oracle.train_step(sd, x, gw)
    err_f = float((fo - f.detach()).norm() / f.detach().norm())
    err_g = max(float((go[k] - p_.grad).norm() / (p_.grad.norm() + 1e-30))
                for k, p_ in (("layer4." + n, q) for n, q in net.layer4.named_parameters()))
    err_s = float((so["layer4.2.bn3.running_var"] - net.layer4[2].bn3.running_var).abs().max())
    print("resnet_train: oracle vs reference  features", err_f, " max grad rel", err_g, " running_var abs", err_s)
    assert err_f < 1e-5 and err_g < 1e-3 and err_s < 1e-5
    np.savez_compressed(os.path.join(OUT, "resnet_train_reference.npz"), **out)
    print("resnet_train: features", tuple(f.shape), "keys", len(out))


# ------------------------------------------------------------------------------- mlp
def gen_mlp():
    import torch.nn as nn
    rna_models = load_ref("2_GeneExpression/models.py", "ref_rna_models")
    joint_models = load_ref("5_JointFusion/models.py", "ref_joint_models2")
    torch.manual_seed(1111)
    model_rna = nn.Sequential(nn.Dropout(), nn.Linear(12778, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 2048))
    head = nn.Sequential(nn.Linear(2048, 1))
    model = rna_models.RNAOnlyModel(model_rna, head).eval()
    x = torch.tensor(det_input((6, 12778), a=0.11))
    with torch.no_grad():
        y = model(x)
        feat = model.extract(x)
    torch.manual_seed(2222)
    early = nn.Sequential(nn.Dropout(), nn.Linear(4096, 2048), nn.ReLU(), nn.Dropout(), nn.Linear(2048, 200),
                          nn.ReLU(), nn.Dropout(), nn.Linear(200, 1)).eval()
    xe = torch.tensor(det_input((5, 4096), a=0.23))
    with torch.no_grad():
        ye = early(xe)
    # joint head on cat([img, rna]) with a stand-in "resnet" that returns given features
    torch.manual_seed(3333)

    class FakeResnet(nn.Module):
        def forward_extract(self, p):
            return p.flatten(1)[:, :2048]

    jhead = nn.Sequential(nn.Dropout(0.8), nn.Linear(4096, 1))
    jm = joint_models.BagHistopathologyRNAModel(FakeResnet(), model_rna, jhead).eval()
    bag = torch.tensor(det_input((6, 2, 1, 32, 64), a=0.05))  # 2048 values per patch
    with torch.no_grad():
        yj = jm(bag, x)
    np.savez_compressed(os.path.join(OUT, "mlp_reference.npz"),
                        rna_out=y.numpy(), rna_feat_head=feat[:, :64].numpy(), rna_feat_sum=feat.sum(1).numpy(),
                        early_out=ye.numpy(), joint_out=yj.numpy())
    print("mlp: rna", tuple(y.shape), "early", tuple(ye.shape), "joint", tuple(yj.shape))


# ------------------------------------------------------------------------- rna script
def gen_rna_script():
    """Runs the UNMODIFIED 2_GeneExpression/1_GeneExpress_train.py on CPU (synthetic CSVs of tests/_rna_script.py, the
    shipped config values, 2 epochs) with a concordance stub that records its arguments, then the call-sequence mirror
    of tests/_rna_script.py with the reference's own `models` module, and requires both recordings to be identical."""
    import json
    import runpy
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _rna_script as R
    _stub_third_party()
    recorded = []

    def recorder(months, neg_scores, vital):
        recorded.append((np.array(months), np.array(neg_scores), np.array(vital)))
        return 0.5

    sys.modules["lifelines.utils"].concordance_index = recorder
    tmp = tempfile.mkdtemp(prefix="rna_script_")
    paths = R.write_csvs(tmp)
    cfg = dict(R.CONFIG, use_cuda=False, train_csv_path=paths["train"], val_csv_path=paths["val"],
               test_csv_path=paths["test"], checkpoint_path=os.path.join(tmp, "out"),
               summary_path=os.path.join(tmp, "out", "summary"))
    cfg_path = os.path.join(tmp, "config_rna_train.json")
    with open(cfg_path, "w") as f:
        json.dump(cfg, f)
    script_dir = os.path.join(REF, "2_GeneExpression")
    for k in ("resnet", "models", "datasets"):
        sys.modules.pop(k, None)
    sys.path.insert(0, script_dir)
    argv = sys.argv
    sys.argv = ["1_GeneExpress_train.py", "--config", cfg_path]
    try:
        runpy.run_path(os.path.join(script_dir, "1_GeneExpress_train.py"), run_name="__main__")
    finally:
        sys.argv = argv
        sys.path.remove(script_dir)
    script_rec = list(recorded)
    ref_models = load_ref("2_GeneExpression/models.py", "ref_rna_models_script")
    mirror_rec, train_losses, last_state = R.run_like_script(ref_models, torch.device("cpu"))
    assert len(script_rec) == len(mirror_rec) == 2 * R.CONFIG["num_epochs"] + 3, (len(script_rec), len(mirror_rec))
    for a, b in zip(script_rec, mirror_rec):
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y)), "the mirror does not reproduce the script"
    out = {"train_losses": np.array(train_losses), "n_calls": np.array(len(script_rec)),
           "final_head_weight": last_state["final_mlp.0.weight"].numpy()}
    for i, (m, s, v) in enumerate(script_rec):
        out[f"call{i}/months"], out[f"call{i}/neg_score"], out[f"call{i}/vital"] = m, s, v
    # the deterministic variant (dropout_p = 0): same call sequence, same weights / sampler order
    nd_rec, nd_losses, nd_state = R.run_like_script(ref_models, torch.device("cpu"), dropout_p=0.0)
    out["nodrop/train_losses"] = np.array(nd_losses)
    out["nodrop/final_head_weight"] = nd_state["final_mlp.0.weight"].numpy()
    for i, (m, s, v) in enumerate(nd_rec):
        assert np.array_equal(m, script_rec[i][0]) and np.array_equal(v, script_rec[i][2])
        out[f"nodrop/call{i}/neg_score"] = s
    np.savez_compressed(os.path.join(OUT, "rna_script_reference.npz"), **out)
    print("rna_script: mirror == unmodified script on", len(script_rec), "evaluate calls; TRAIN losses", train_losses)


# ------------------------------------------------------------------------------- nll
def gen_nll():
    """nll_loss / NLLSurvLoss of 1_HistoPathology/models.py:120-232 on seeded inputs: loss and d loss / d h (autograd),
    mean and sum reductions, alpha 0 and 0.4, including logits extreme enough to hit the eps clamps."""
    _stub_third_party()
    m = load_ref("1_HistoPathology/models.py", "ref_histo_models_nll")
    rng = np.random.default_rng(21)
    out = {}
    for ci, (n, k, alpha, reduction, scale) in enumerate([(128, 4, 0.0, "mean", 1.0), (37, 4, 0.4, "sum", 1.0),
                                                          (64, 8, 0.0, "mean", 12.0), (5, 1, 0.0, "mean", 1.0)]):
        h = torch.tensor((rng.standard_normal((n, k)) * scale).astype(np.float32), requires_grad=True)
        y = torch.tensor(rng.integers(0, k, n).astype(np.int64))
        c = torch.tensor((rng.uniform(size=n) < 0.4).astype(np.float32))
        loss = m.NLLSurvLoss(alpha=alpha, reduction=reduction)(h, y, c)
        loss.backward()
        out[f"case{ci}/h"], out[f"case{ci}/y"], out[f"case{ci}/c"] = h.detach().numpy(), y.numpy(), c.numpy()
        out[f"case{ci}/alpha"], out[f"case{ci}/mean"] = np.float32(alpha), np.int32(reduction == "mean")
        out[f"case{ci}/loss"], out[f"case{ci}/grad"] = loss.detach().numpy(), h.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "nll_reference.npz"), **out)
    print("nll: 4 cases")


# ------------------------------------------------------------------------- attention / project / stems
def gen_variants():
    """TanhAttention, AggregationModel / AggregationProjectModel around it, and the 1- / 4-channel stems, all run through the
    reference's own classes (5_JointFusion/models.py, resnet.py) with seeded parameters that the test rebuilds."""
    import torch.nn as nn
    joint_models = load_ref("5_JointFusion/models.py", "ref_joint_models3")
    ref_resnet = load_ref("5_JointFusion/resnet.py", "ref_resnet3")

    class FakeResnet(nn.Module):
        def forward_extract(self, p):
            return p.flatten(1)

    out = {}
    torch.manual_seed(4444)
    dim, hdim = 256, 64
    att = joint_models.TanhAttention(dim)
    with torch.no_grad():
        att.vector.normal_(0, 0.5)
    x = torch.tensor(det_input((3, 5, dim), a=0.31))
    with torch.no_grad():
        o, a = att(x)
    out.update(att_x=x.numpy(), att_vector=att.vector.detach().numpy(), att_linear=att.linear.weight.detach().numpy(),
               att_out=o.numpy(), att_weights=a.numpy())
    agg = joint_models.AggregationModel(FakeResnet(), att, dim, resnet_dim=dim).eval()
    proj = joint_models.AggregationProjectModel(FakeResnet(), att, dim, resnet_dim=dim, hdim=hdim).eval()
    bag = x.view(3, 5, 1, 16, 16)
    with torch.no_grad():
        y_agg, _ = agg(bag)
        f_agg, _ = agg.extract(bag)
        y_proj, _ = proj(bag)
        f_proj, _ = proj.extract(bag)
    out.update(agg_fc_w=agg.fc.weight.detach().numpy(), agg_fc_b=agg.fc.bias.detach().numpy(), agg_out=y_agg.numpy(),
               agg_feat=f_agg.numpy(), proj_w=proj.project.weight.detach().numpy(), proj_b=proj.project.bias.detach().numpy(),
               proj_fc_w=proj.fc.weight.detach().numpy(), proj_fc_b=proj.fc.bias.detach().numpy(), proj_out=y_proj.numpy(),
               proj_feat=f_proj.numpy())
    # 1- / 4-channel stems: seeded trunk = the 3-channel seeded state_dict of the resnet oracle + a seeded conv1
    from oracle import resnet_oracle
    sd = resnet_oracle.init_state_dict(seed=77)
    for c, ctor in ((4, ref_resnet.resnet50_4channel), (1, ref_resnet.resnet50_1channel)):
        net = ctor(pretrained=False).eval()
        own = net.state_dict()
        own.update({k: v for k, v in sd.items() if k != "conv1.weight"})
        g = torch.Generator().manual_seed(100 + c)
        own["conv1.weight"] = torch.randn(64, c, 7, 7, generator=g) * (2.0 / (49 * 64)) ** 0.5
        net.load_state_dict(own)
        xi = torch.tensor(det_input((2, c, 224, 224), a=0.013 * c))
        with torch.no_grad():
            f = net.forward_extract(xi)
        out["stem%d_feat" % c] = f.numpy()
    np.savez_compressed(os.path.join(OUT, "variants_reference.npz"), **out)
    print("variants:", {k: v.shape for k, v in out.items() if k.endswith(("out", "feat"))})


# ------------------------------------------------------------------------- augmentation
def gen_augment():
    """Patches through the reference's training transforms (2_HistoPath_train.py:474-488) run by torchvision on PIL
    images, with the per-image parameters torchvision drew (replayed by pipeline.sample_augment_params under the same
    seed): inputs, parameter rows and outputs for the device kernel's parity test."""
    from PIL import Image
    import torchvision.transforms as T
    from multimodalbrainsurvival_b200 import pipeline
    size, n = 48, 8
    tf = T.Compose([T.Resize(size), T.RandomHorizontalFlip(), T.RandomVerticalFlip(),
                    T.ColorJitter(64.0 / 255, 0.75, 0.25, 0.04)])
    rng = np.random.default_rng(20261018)
    imgs = rng.integers(0, 256, (n, size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    imgs[1] = np.stack([(yy * 5) % 256, (xx * 5) % 256, (yy + xx) * 2 % 256], -1)           # smooth ramps
    imgs[2] = (np.array([200, 180, 190]) + rng.integers(-6, 6, (size, size, 3))).clip(0, 255)  # pale H&E-like patch
    imgs[3] = 128                                                                            # flat grey
    torch.manual_seed(4242)
    outs = np.stack([np.asarray(tf(Image.fromarray(im, "RGB"))) for im in imgs])
    torch.manual_seed(4242)
    rows = pipeline.sample_augment_params(n).numpy()
    # transforms.Resize on non-square, non-224 patches (smaller edge -> 48; down- and up-scaling)
    rs_in = rng.integers(0, 256, (2, 70, 90, 3), dtype=np.uint8)
    rs_out = np.stack([np.asarray(T.Resize(48)(Image.fromarray(im, "RGB"))) for im in rs_in])
    up_in = rng.integers(0, 256, (2, 20, 20, 3), dtype=np.uint8)
    up_out = np.stack([np.asarray(T.Resize(48)(Image.fromarray(im, "RGB"))) for im in up_in])
    np.savez_compressed(os.path.join(OUT, "augment_reference.npz"), imgs=imgs, params=rows, outs=outs,
                        resize_in=rs_in, resize_out=rs_out, upscale_in=up_in, upscale_out=up_out)
    print("augment:", imgs.shape, rows.shape, outs.shape, rs_out.shape, up_out.shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["cox", "aggregate", "resnet", "resnet_train", "mlp", "nll", "rna_script", "variants", "augment"]
    for w in which:
        globals()["gen_" + w]()
