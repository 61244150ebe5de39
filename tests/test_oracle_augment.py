"""CPU: (1) oracle/augment_oracle.py pinned against the installed torchvision / Pillow - exhaustively for the colour-space
conversions, over every byte pair for the blend, end to end on images and through transforms.Compose with a shared seed;
(2) the library's host PNG decoder against PIL, bit-exact.  (Reference: the transforms of 2_HistoPath_train.py:474-488 and
the decode of PatchBagDataset.__getitem__, 1_HistoPathology/models.py:280-284.)"""
import os

import numpy as np
import pytest
import torch

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

from oracle import augment_oracle as ao  # noqa: E402


def _all_triples():
    r, g, b = np.meshgrid(*(np.arange(256, dtype=np.uint8),) * 3, indexing="ij")
    return np.stack([r, g, b], -1).reshape(4096, 4096, 3)


def test_colour_conversions_match_pillow_exhaustively():
    rgb = _all_triples()
    img = Image.fromarray(rgb, "RGB")
    assert np.array_equal(ao.rgb_to_l(rgb), np.asarray(img.convert("L")))
    assert np.array_equal(ao.rgb_to_hsv(rgb), np.asarray(img.convert("HSV")))
    assert np.array_equal(ao.hsv_to_rgb(rgb), np.asarray(Image.fromarray(rgb, "HSV").convert("RGB")))


@pytest.mark.parametrize("alpha", [0.0, 0.2509804, 0.75, 1.0, 1.2490196, 1.75, 0.7490196078431373])
def test_blend_matches_pillow_for_every_byte_pair(alpha):
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    ref = np.asarray(Image.blend(Image.fromarray(a, "L"), Image.fromarray(b, "L"), alpha))
    assert np.array_equal(ao.blend(a, b, alpha), ref)


def test_chain_matches_torchvision_functional():
    TF = pytest.importorskip("torchvision.transforms.functional")
    rng = np.random.default_rng(0)
    ops = [TF.adjust_brightness, TF.adjust_contrast, TF.adjust_saturation, TF.adjust_hue]
    for trial in range(12):
        img = rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)
        if trial % 2:   # smooth image: small saturation, the grey branch of the HSV conversion
            img = (rng.integers(0, 256, (1, 1, 3)) + rng.integers(-30, 30, (64, 48, 3))).clip(0, 255).astype(np.uint8)
        f = [rng.uniform(1 - 64 / 255, 1 + 64 / 255), rng.uniform(0.25, 1.75), rng.uniform(0.75, 1.25), rng.uniform(-0.04, 0.04)]
        order = rng.permutation(4)
        hf, vf = bool(rng.integers(2)), bool(rng.integers(2))
        p = Image.fromarray(img, "RGB")
        p = TF.hflip(p) if hf else p
        p = TF.vflip(p) if vf else p
        for op in order:
            p = ops[op](p, f[op])
        assert np.array_equal(ao.augment(img, hf, vf, order, f), np.asarray(p)), trial


def test_parameter_sampler_follows_torchvision_compose():
    """Same torch seed -> the reference's Compose on a PIL image == our sampled parameters through the oracle."""
    T = pytest.importorskip("torchvision.transforms")
    from multimodalbrainsurvival_b200 import pipeline
    tf = T.Compose([T.Resize(32), T.RandomHorizontalFlip(), T.RandomVerticalFlip(),
                    T.ColorJitter(64.0 / 255, 0.75, 0.25, 0.04)])
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (5, 32, 32, 3), dtype=np.uint8)
    torch.manual_seed(77)
    ref = [np.asarray(tf(Image.fromarray(im, "RGB"))) for im in imgs]
    torch.manual_seed(77)
    rows = pipeline.sample_augment_params(5).numpy()
    for im, row, want in zip(imgs, rows, ref):
        f = row[6:9].view(np.float32).astype(np.float64)
        # (the hue factor reaches the kernel as the integer shift torchvision derives from it: replay that)
        full = im[:, ::-1] if row[0] else im
        full = full[::-1] if row[1] else full
        full = np.ascontiguousarray(full)
        for op in row[2:6]:
            if op == 3:
                hsv = ao.rgb_to_hsv(full)
                hsv[..., 0] = (hsv[..., 0].astype(np.uint32) + np.uint32(row[9])).astype(np.uint8)
                full = ao.hsv_to_rgb(hsv)
            else:
                full = (ao.adjust_brightness, ao.adjust_contrast, ao.adjust_saturation)[op](full, float(f[op]))
        assert np.array_equal(full, want)


@pytest.mark.parametrize("shape", [(256, 256, 224, 224), (512, 512, 224, 224), (224, 224, 299, 299), (300, 200, 224, 149),
                                   (97, 131, 224, 224), (1000, 1000, 224, 224)])
def test_resize_matches_pillow(shape):
    """Image.resize(BILINEAR) = transforms.Resize on PIL images (antialiased triangle filter, two uint8 passes); the
    library's host coefficient routine (csrc/augment.cu mmbs_resample_coeffs) must produce the oracle's rows."""
    from multimodalbrainsurvival_b200 import pipeline
    h, w, oh, ow = shape
    img = np.random.default_rng(h + ow).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(ao.resize_bilinear(img, oh, ow), ref)
    for n_in, n_out in ((w, ow), (h, oh)):
        bo, ko = ao.resample_coeffs(n_in, n_out)
        bl, kl = pipeline.resample_coeffs(n_in, n_out)
        assert np.array_equal(bo, bl) and np.array_equal(ko, kl)


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "P", "LA"])
def test_png_decoder_matches_pillow(tmp_path, mode):
    from multimodalbrainsurvival_b200 import pipeline
    rng = np.random.default_rng(len(mode))
    paths, want = [], []
    for i in range(6):
        base = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
        if i % 2:   # smooth content: the encoder picks the Sub / Up / Average / Paeth filters
            yy, xx = np.mgrid[0:224, 0:224]
            base = np.stack([(yy + xx) // 2, (yy * 2) % 256, (xx * 3 + i) % 256], -1).astype(np.uint8)
        im = Image.fromarray(base, "RGB")
        if mode == "P":
            im = im.quantize(64)
        elif mode != "RGB":
            im = im.convert(mode)
        f = os.path.join(tmp_path, f"{mode}_{i}.png")
        im.save(f, optimize=bool(i % 3 == 0))
        paths.append(f)
        want.append(np.asarray(Image.open(f).convert("RGB")))
    got = pipeline.decode_png_files(paths, size=224, threads=3).numpy()
    assert np.array_equal(got, np.stack(want))


def test_png_decoder_reports_bad_files(tmp_path):
    from multimodalbrainsurvival_b200 import pipeline
    f = os.path.join(tmp_path, "small.png")
    Image.fromarray(np.zeros((10, 10, 3), np.uint8), "RGB").save(f)
    with pytest.raises(RuntimeError, match="patch size"):
        pipeline.decode_png_files([f], size=224)
    with pytest.raises(RuntimeError, match="cannot read"):
        pipeline.decode_png_files([os.path.join(tmp_path, "missing.png")], size=224)
