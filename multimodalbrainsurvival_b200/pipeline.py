"""Host <-> device staging helpers for the extract / savescore loops.

The reference moves every batch with a blocking ``.to(device)`` and reads features back with
``.cpu()`` (/root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:60-71).
``prefetch_to_device`` overlaps the H2D copy of batch i+1 with the kernels of batch i on a
side stream, through a fixed ring of device staging buffers (no allocation, and therefore no
implicit synchronisation, inside the loop)."""
from __future__ import annotations

import torch


def prefetch_to_device(batches, device, depth: int = 2):
    """Iterate over host tensors - or tuples / lists of host tensors (e.g. (patch_bag, rna) of the joint-fusion loader:
    copies that share the stream cannot queue behind each other's batches on the DMA engine) - yielding device tensors
    (tuples) whose copy was issued ``depth`` steps ahead.  A yielded item is only valid until the next one is requested
    (its buffers are reused)."""
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = depth + 1
    ring = [None] * slots          # device staging buffers (a list per slot)
    ready = [None] * slots         # copy finished (recorded on the copy stream)
    released = [None] * slots      # consumer finished with the buffers (recorded on its stream)
    it = iter(batches)
    head = 0                       # next slot to fill
    pending = []

    def issue():
        nonlocal head
        try:
            item = next(it)
        except StopIteration:
            return False
        single = isinstance(item, torch.Tensor)
        hosts = [item] if single else list(item)
        hosts = [h if h.is_pinned() else h.pin_memory() for h in hosts]
        k = head
        head = (head + 1) % slots
        if ring[k] is None or len(ring[k]) != len(hosts) or any(
                r.shape != h.shape or r.dtype != h.dtype for r, h in zip(ring[k], hosts)):
            ring[k] = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in hosts]
        with torch.cuda.stream(copy_stream):
            if released[k] is not None:
                copy_stream.wait_event(released[k])
            for r, h in zip(ring[k], hosts):
                r.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        ready[k] = ev
        pending.append((k, hosts, single))   # keep the pinned sources alive until their copies are consumed
        return True

    for _ in range(depth):
        if not issue():
            break
    while pending:
        k, _hosts, single = pending.pop(0)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[k])
        issue()
        yield ring[k][0] if single else tuple(ring[k])
        rel = torch.cuda.Event()
        rel.record(torch.cuda.current_stream(dev))
        released[k] = rel


class HostWriter:
    """Device -> pinned-host copies on a side stream: the D2H read-back of step i overlaps the kernels of step i+1
    instead of sitting in the compute stream (a 4 MB feature block at PCIe speed stalls it for ~0.15 ms).
    ``write`` orders the copy after everything enqueued so far on the current stream; ``wait`` makes the host
    wait for all pending copies (call it before reading the host buffers)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def write(self, src: torch.Tensor, dst_host: torch.Tensor) -> None:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            dst_host.copy_(src, non_blocking=True)
        src.record_stream(self.stream)     # the caching allocator must not recycle src before the copy ran

    def wait(self) -> None:
        self.stream.synchronize()


# ----------------------------------------------------------------------------- patch decode + augmentation
AUG_PARAM_WORDS = 10   # include/mmbs.h mmbs_aug_params: hflip, vflip, order[4], factor[3] (fp32 bits), hue_shift


def decode_png_files(paths, size: int = 224, out: torch.Tensor | None = None, threads: int = 0) -> torch.Tensor:
    """Decode patch files the way ``Image.open(f).convert('RGB')`` does (reference PatchBagDataset.__getitem__,
    1_HistoPathology/models.py:280-284) with the library's multi-threaded host decoder (csrc/png.cu):
    -> uint8 [len(paths), size, size, 3], pinned when CUDA is available so that it can be copied asynchronously."""
    import ctypes
    from . import _lib
    n = len(paths)
    if out is None:
        out = torch.empty((n, size, size, 3), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    if tuple(out.shape) != (n, size, size, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or out.is_cuda:
        raise ValueError("decode_png_files: `out` must be a contiguous host uint8 tensor [n, size, size, 3]")
    arr = (ctypes.c_char_p * max(n, 1))(*[str(p).encode() for p in paths])
    _lib.check(_lib.lib().mmbs_png_decode_files(arr, n, ctypes.c_void_p(out.data_ptr()), size, size, int(threads)),
               "mmbs_png_decode_files")
    return out


def sample_augment_params(batch: int, brightness: float = 64.0 / 255, contrast: float = 0.75, saturation: float = 0.25,
                          hue: float = 0.04, p_flip: float = 0.5, generator: torch.Generator | None = None) -> torch.Tensor:
    """Per-image parameters of RandomHorizontalFlip -> RandomVerticalFlip -> ColorJitter(brightness, contrast, saturation,
    hue) (reference 2_HistoPath_train.py:474-488), drawn from the torch CPU generator in torchvision's own order: one
    ``rand(1)`` per flip, ``randperm(4)``, then one ``uniform_`` per factor.  -> int32 [batch, 10] (mmbs_aug_params rows)."""
    import numpy as np
    rows = np.zeros((batch, AUG_PARAM_WORDS), dtype=np.int32)
    fview = rows.view(np.float32)
    ranges = ((max(0.0, 1.0 - brightness), 1.0 + brightness), (max(0.0, 1.0 - contrast), 1.0 + contrast),
              (max(0.0, 1.0 - saturation), 1.0 + saturation), (-hue, hue))
    for i in range(batch):
        rows[i, 0] = int(torch.rand(1, generator=generator) < p_flip)
        rows[i, 1] = int(torch.rand(1, generator=generator) < p_flip)
        rows[i, 2:6] = torch.randperm(4, generator=generator).numpy()
        f = [float(torch.empty(1).uniform_(lo, hi, generator=generator)) for lo, hi in ranges]
        fview[i, 6:9] = f[:3]
        rows[i, 9] = int(np.int32(f[3] * 255).astype(np.uint8))   # torchvision adjust_hue's wrap-around shift
    return torch.from_numpy(rows)


def augment(patches_u8_hwc: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
    """Flips + ColorJitter on the device (csrc/augment.cu), bit-exact with torchvision on PIL images:
    uint8 [B, H, W, 3] (CUDA) + int32 [B, 10] parameter rows -> uint8 [B, 3, H, W], what ``forward_extract`` takes as raw
    pixels (ToTensor + Normalize happen in its stem pack kernel)."""
    from . import _lib
    x = patches_u8_hwc
    if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[3] == 3):
        raise RuntimeError("augment: patches must be a CUDA uint8 tensor [B, H, W, 3] (no CPU path in this build)")
    b, h, w, _ = x.shape
    if tuple(params.shape) != (b, AUG_PARAM_WORDS) or params.dtype != torch.int32:
        raise ValueError("augment: params must be int32 [B, 10] (sample_augment_params)")
    x = x.contiguous()
    p = params.to(x.device, non_blocking=True).contiguous()
    out = torch.empty((b, 3, h, w), dtype=torch.uint8, device=x.device)
    ws = torch.empty(b, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmbs_augment_u8(_lib.ptr(x), _lib.ptr(out), b, h, w, _lib.ptr(p), _lib.ptr(ws),
                                              _lib.stream_ptr()), "mmbs_augment_u8")
    return out


_RESAMPLE_TABLES = {}   # (in_size, out_size, device index) -> (bounds, coefficients, ksize) on the device


def resample_coeffs(in_size: int, out_size: int):
    """Pillow's bilinear coefficient rows for one axis (host code in the library, csrc/augment.cu):
    -> (bounds int32 [out, 2], coefficients int32 [out, ksize])."""
    import ctypes
    import numpy as np
    from . import _lib
    L = _lib.lib()
    ksize = L.mmbs_resample_coeffs(in_size, out_size, None, None, 0)
    if ksize < 1:
        raise ValueError(f"resample_coeffs: bad sizes {in_size} -> {out_size}")
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    rc = L.mmbs_resample_coeffs(in_size, out_size, ctypes.c_void_p(bounds.ctypes.data), ctypes.c_void_p(kk.ctypes.data), kk.size)
    if rc != ksize:
        raise RuntimeError("mmbs_resample_coeffs failed")
    return bounds, kk


def resize(patches_u8_hwc: torch.Tensor, size) -> torch.Tensor:
    """``transforms.Resize(size)`` of the reference's transform chains (2_HistoPath_train.py:476,484) on a batch of decoded
    patches: uint8 [B, H, W, 3] (CUDA) -> uint8 [B, h, w, 3], bit-exact with PIL's antialiased bilinear resize.  An int
    `size` scales the smaller edge to it (torchvision's rule); patches already of that size are returned as they are."""
    from . import _lib
    x = patches_u8_hwc
    if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[3] == 3):
        raise RuntimeError("resize: patches must be a CUDA uint8 tensor [B, H, W, 3] (no CPU path in this build)")
    b, h, w, _ = x.shape
    if isinstance(size, int):
        if h <= w:
            oh, ow = size, int(size * w / h)
        else:
            oh, ow = int(size * h / w), size
    else:
        oh, ow = int(size[0]), int(size[1])
    if (oh, ow) == (h, w):
        return x
    x = x.contiguous()
    tabs = []
    for n_in, n_out in ((w, ow), (h, oh)):
        key = (n_in, n_out, x.device.index)
        t = _RESAMPLE_TABLES.get(key)
        if t is None:
            bounds, kk = resample_coeffs(n_in, n_out)
            t = (torch.from_numpy(bounds).to(x.device), torch.from_numpy(kk).to(x.device), kk.shape[1])
            if len(_RESAMPLE_TABLES) > 64:
                _RESAMPLE_TABLES.clear()
            _RESAMPLE_TABLES[key] = t
        tabs.append(t)
    out = torch.empty((b, oh, ow, 3), dtype=torch.uint8, device=x.device)
    tmp = torch.empty((b, h, ow, 3), dtype=torch.uint8, device=x.device)
    (bw, kw, ksw), (bh, kh, ksh) = tabs
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().mmbs_resize_bilinear_u8(_lib.ptr(x), _lib.ptr(out), _lib.ptr(tmp), b, h, w, oh, ow, _lib.ptr(bw),
                                                      _lib.ptr(kw), ksw, _lib.ptr(bh), _lib.ptr(kh), ksh, _lib.stream_ptr()),
                   "mmbs_resize_bilinear_u8")
    return out
