"""TEST INFRASTRUCTURE ONLY - numpy restatement of the uint8 arithmetic behind the reference's training transforms
(`transforms.RandomHorizontalFlip / RandomVerticalFlip / ColorJitter(64/255, 0.75, 0.25, 0.04)` on PIL images,
/root/reference/1_HistoPathology/2_HistoPath_train.py:474-488; the PIL image comes from PatchBagDataset.__getitem__,
1_HistoPathology/models.py:280-286).

The arithmetic itself lives in two third-party packages that are not vendored by the reference and not pinned by it:
torchvision (`transforms/_functional_pil.py`: adjust_brightness / contrast / saturation / hue, hflip, vflip) and Pillow
(`ImageEnhance`, `Image.blend` = libImaging/Blend.c, `convert("L" | "HSV" | "RGB")` = libImaging/Convert.c).  The image
used here has torchvision 0.26.0 and Pillow 12.2.0; this file restates their published algorithms and is PINNED
against those installed packages by tests/test_oracle_augment.py: exhaustively for the colour-space conversions (all 2^24
RGB and HSV triples), for every (a, b) byte pair of the blend at the factors the tests use, and end to end on images.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
from __future__ import annotations

import numpy as np

_F32 = np.float32


def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """Pillow Convert.c rgb2l: ITU-R 601-2 luma with 16-bit fixed-point weights, rounded."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(in1: np.ndarray, in2: np.ndarray, alpha: float) -> np.ndarray:
    """Pillow Blend.c ImagingBlend: out = in1 + alpha * (in2 - in1) in single precision (separate multiply and add),
    truncated; clipped to [0, 255] only when alpha is outside [0, 1] (inside, the result cannot leave the range)."""
    a = _F32(alpha)
    d = in2.astype(np.int32) - in1.astype(np.int32)
    t = in1.astype(_F32) + a * d.astype(_F32)       # numpy float32 ops round after each step, like the C code
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    out = np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32)))
    return out.astype(np.uint8)


def adjust_brightness(img: np.ndarray, factor: float) -> np.ndarray:
    """ImageEnhance.Brightness: blend(black, image, factor)."""
    return blend(np.zeros_like(img), img, factor)


def adjust_contrast(img: np.ndarray, factor: float) -> np.ndarray:
    """ImageEnhance.Contrast: blend(uniform grey of int(mean(L) + 0.5), image, factor)."""
    l = rgb_to_l(img)
    mean = int(float(l.astype(np.float64).sum()) / l.size + 0.5)
    return blend(np.full_like(img, mean), img, factor)


def adjust_saturation(img: np.ndarray, factor: float) -> np.ndarray:
    """ImageEnhance.Color: blend(L image replicated to three channels, image, factor)."""
    l = rgb_to_l(img)
    return blend(np.repeat(l[..., None], 3, axis=-1), img, factor)


def rgb_to_hsv(rgb: np.ndarray) -> np.ndarray:
    """Pillow Convert.c rgb2hsv_row (after colorsys.py): float ratios, the hue expressions evaluated in double and stored
    to a float, truncation to bytes."""
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    grey = maxc == minc
    cr = np.where(grey, 1, maxc - minc).astype(_F32)
    s = cr / np.where(maxc == 0, 1, maxc).astype(_F32)
    rc = (maxc - r).astype(_F32) / cr
    gc = (maxc - g).astype(_F32) / cr
    bc = (maxc - b).astype(_F32) / cr
    # `float h`: every assignment rounds the double expression on its right to single precision
    h = np.where(r == maxc, bc - gc,
                 np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(_F32),
                          (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(_F32))).astype(_F32)
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(_F32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    out = np.stack([np.where(grey, 0, uh), np.where(grey, 0, us), maxc], axis=-1)
    return out.astype(np.uint8)


def hsv_to_rgb(hsv: np.ndarray) -> np.ndarray:
    """Pillow Convert.c hsv2rgb (after colorsys.py)."""
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    hf = h.astype(_F32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(_F32).astype(np.float64)).astype(_F32)
    fs = (s.astype(_F32).astype(np.float64) / 255.0).astype(_F32)
    vf = v.astype(_F32).astype(np.float64)
    fsd, fd = fs.astype(np.float64), f.astype(np.float64)
    p = np.clip(np.round(vf * (1.0 - fsd)), 0, 255).astype(np.int32)
    q = np.clip(np.round(vf * (1.0 - fsd * fd)), 0, 255).astype(np.int32)
    t = np.clip(np.round(vf * (1.0 - fsd * (1.0 - fd))), 0, 255).astype(np.int32)
    k = i % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    grey = s == 0
    return np.stack([np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)], axis=-1).astype(np.uint8)


def adjust_hue(img: np.ndarray, factor: float) -> np.ndarray:
    """torchvision _functional_pil.adjust_hue: H channel + uint8(int32(factor * 255)) with wrap-around."""
    hsv = rgb_to_hsv(img)
    shift = np.int32(factor * 255).astype(np.uint8)
    hsv[..., 0] = (hsv[..., 0].astype(np.uint32) + np.uint32(shift)).astype(np.uint8)
    return hsv_to_rgb(hsv)


_OPS = (adjust_brightness, adjust_contrast, adjust_saturation, adjust_hue)


def augment(img_hwc: np.ndarray, hflip: bool, vflip: bool, order, factors) -> np.ndarray:
    """One patch through the reference's training transform chain up to (not including) ToTensor/Normalize:
    RandomHorizontalFlip -> RandomVerticalFlip -> ColorJitter with the four operations applied in `order` (a permutation
    of 0 = brightness, 1 = contrast, 2 = saturation, 3 = hue; an index of -1 skips) and their factors.  uint8 [H, W, 3]."""
    out = img_hwc
    if hflip:
        out = out[:, ::-1]
    if vflip:
        out = out[::-1]
    out = np.ascontiguousarray(out)
    for op in order:
        if op >= 0:
            out = _OPS[op](out, float(factors[op]))
    return out


# ----------------------------------------------------------------------------- Resize (Pillow libImaging/Resample.c)
PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter over the whole axis:
    -> (bounds int32 [out, 2] = (first input index, count), coefficients int32 [out, ksize] in 22-bit fixed point)."""
    scale = float(in_size) / out_size            # (double)(in1 - in0) / outSize
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                  # bilinear support = 1
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            v = 1.0 - t if t < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`Image.resize((out_w, out_h), Image.BILINEAR)` = what transforms.Resize does to a PIL image (antialiased triangle
    filter, horizontal pass then vertical pass, each rounded to uint8)."""
    h, w = img_hwc.shape[:2]
    out = img_hwc
    if out_w != w:
        out = _resample_axis(out, *resample_coeffs(w, out_w), axis=1)
    if out_h != h:
        out = _resample_axis(out, *resample_coeffs(h, out_h), axis=0)
    return out
