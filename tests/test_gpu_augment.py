"""GPU parity of csrc/augment.cu (flips + ColorJitter on uint8 patches) - bit-exact, it is byte arithmetic:
(1) against outputs torchvision itself produced on PIL images (tests/golden/augment_reference.npz, tools/make_golden.py
gen_augment; reference transforms: 1_HistoPathology/2_HistoPath_train.py:474-488), (2) against the pinned oracle on
224 x 224 patches with every operation order, (3) decode -> augment -> forward_extract runs end to end."""
import itertools
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(imgs, rows):
    from multimodalbrainsurvival_b200 import pipeline
    out = pipeline.augment(torch.tensor(imgs, device=DEV), torch.tensor(rows))
    torch.cuda.synchronize()
    return out.permute(0, 2, 3, 1).cpu().numpy()   # back to HWC


def test_matches_torchvision_golden(golden):
    g = golden("augment_reference.npz")
    got = _run(g["imgs"], g["params"])
    assert got.shape == g["outs"].shape
    assert np.array_equal(got, g["outs"]), f"{int((got != g['outs']).sum())} bytes differ from torchvision's output"


def test_every_operation_order_matches_the_oracle():
    from multimodalbrainsurvival_b200 import pipeline
    rng = np.random.default_rng(11)
    orders = list(itertools.permutations(range(4)))
    n = len(orders)
    imgs = rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
    imgs[::3] = (rng.integers(0, 256, (len(imgs[::3]), 1, 1, 3)) + rng.integers(-25, 25, (len(imgs[::3]), 224, 224, 3))).clip(0, 255)
    torch.manual_seed(5)
    rows = pipeline.sample_augment_params(n).numpy()
    rows[:, 2:6] = np.array(orders, dtype=np.int32)
    rows[5, 2:6] = (-1, 3, -1, 0)          # operations switched off
    rows[7, 9] = 250                        # a negative hue factor wraps around
    got = _run(imgs, rows)
    for i in range(n):
        f = rows[i, 6:9].view(np.float32).astype(np.float64)
        x = imgs[i][:, ::-1] if rows[i, 0] else imgs[i]
        x = np.ascontiguousarray(x[::-1] if rows[i, 1] else x)
        for op in rows[i, 2:6]:
            if op == 3:
                hsv = ao.rgb_to_hsv(x)
                hsv[..., 0] = (hsv[..., 0].astype(np.uint32) + np.uint32(rows[i, 9])).astype(np.uint8)
                x = ao.hsv_to_rgb(hsv)
            elif op >= 0:
                x = (ao.adjust_brightness, ao.adjust_contrast, ao.adjust_saturation)[op](x, float(f[op]))
        assert np.array_equal(got[i], x), (i, rows[i].tolist(), int((got[i] != x).sum()))


def test_resize_matches_torchvision_golden_and_oracle(golden):
    from multimodalbrainsurvival_b200 import pipeline
    g = golden("augment_reference.npz")
    for key in ("resize", "upscale"):
        got = pipeline.resize(torch.tensor(g[key + "_in"], device=DEV), 48).cpu().numpy()
        assert got.shape == g[key + "_out"].shape and np.array_equal(got, g[key + "_out"]), key
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, (3, 256, 256, 3), dtype=np.uint8)
    got = pipeline.resize(torch.tensor(x, device=DEV), 224).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], ao.resize_bilinear(x[i], 224, 224))
    same = torch.tensor(x[:, :224, :224].copy(), device=DEV)
    assert pipeline.resize(same, 224) is same   # the reference's 224 x 224 patches: Resize(224) is the identity


def test_decode_augment_extract_pipeline(tmp_path):
    """PNG files -> host decoder -> device augmentation -> forward_extract on raw pixels: what a training loader feeds."""
    Image = pytest.importorskip("PIL.Image")
    from multimodalbrainsurvival_b200 import pipeline, resnet
    rng = np.random.default_rng(2)
    paths = []
    for i in range(4):
        f = os.path.join(tmp_path, f"WSI_patch_{i}.png")
        Image.fromarray(rng.integers(0, 256, (224, 224, 3), dtype=np.uint8), "RGB").save(f)
        paths.append(f)
    host = pipeline.decode_png_files(paths)
    rows = pipeline.sample_augment_params(4)
    rows[:, 0:2] = 0
    rows[:, 2:6] = -1                       # identity parameters: the kernel must reproduce the decoded pixels
    x = pipeline.augment(host.to(DEV, non_blocking=True), rows)
    assert torch.equal(x.cpu(), host.permute(0, 3, 1, 2))
    net = resnet.randomize_batchnorm_(resnet.resnet50()).to(DEV).eval()
    with torch.no_grad():
        f = net.forward_extract(pipeline.augment(host.to(DEV), pipeline.sample_augment_params(4)))
    assert tuple(f.shape) == (4, 2048) and bool(torch.isfinite(f).all())


def test_training_engine_takes_raw_pixels():
    """model.train(): uint8 pixels (the augmentation kernel's output) normalised by the stem pack kernel give the
    features of the fp32 tensor ToTensor + Normalize would have produced.  The two inputs differ by fp32 rounding of the
    normalisation before the bf16 cast; batch-statistics BatchNorm over 8 patches amplifies that like any other bf16
    rounding (training-mode features are within 5e-3 of the fp32 reference) - same bound here."""
    from multimodalbrainsurvival_b200 import resnet
    from oracle import resnet_oracle
    net = resnet.resnet50()
    net.load_state_dict(resnet_oracle.init_state_dict(seed=9, bn3_gamma_scale=0.1))
    net = net.to(DEV).train()
    for p in net.parameters():
        p.requires_grad = False
    for p in net.layer4.parameters():
        p.requires_grad = True
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (8, 3, 224, 224), dtype=torch.uint8, generator=g).to(DEV)
    mean = torch.tensor(net.input_mean, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(net.input_std, device=DEV).view(1, 3, 1, 1)
    f_u8 = net.forward_extract(u8)
    assert net._engines and f_u8.requires_grad
    f_u8.sum().backward()
    assert net.layer4[2].conv3.weight.grad is not None
    f_fp = net.forward_extract((u8.float() / 255.0 - mean) / std)
    rel = float((f_u8.detach() - f_fp.detach()).norm() / f_fp.detach().norm())
    assert rel < 5e-3, rel
