"""ctypes binding of libmmbs.so (the C ABI declared in include/mmbs.h).

The shared library is built in-tree (``multimodalbrainsurvival_b200/libmmbs.so``)
by ``csrc/Makefile`` for sm_100a only.  There is no CPU fallback: if the library
is missing, or no sm_100 device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmbs.so")
CSRC = os.path.join(_HERE, "csrc")

_lock = threading.Lock()
_lib = None

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_size = ctypes.c_size_t
c_float = ctypes.c_float


class ConvDesc(ctypes.Structure):
    """mmbs_conv_desc (include/mmbs.h)."""
    _fields_ = [
        ("batch", c_i32), ("in_h", c_i32), ("in_w", c_i32), ("c_in", c_i32), ("c_out", c_i32),
        ("ksize", c_i32), ("stride", c_i32), ("relu", c_i32), ("out_f32", c_i32), ("flags", c_i32),
        ("in_", c_void_p), ("weight", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
        ("residual", c_void_p), ("out", c_void_p), ("stats", c_void_p),
    ]


class BnTrainDesc(ctypes.Structure):
    """mmbs_bn_train_desc (include/mmbs.h)."""
    _fields_ = [
        ("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p),
        ("running_var", c_void_p), ("scale_out", c_void_p), ("shift_out", c_void_p), ("mean_out", c_void_p),
        ("invstd_out", c_void_p), ("eps", c_float), ("momentum", c_float), ("count", c_i64), ("c", c_i64),
    ]


class AdamGroup(ctypes.Structure):
    """mmbs_adam_group (include/mmbs.h)."""
    _fields_ = [("step_size", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("weight_decay", c_float), ("bias_correction2_sqrt", c_float),
                ("one_minus_beta1", c_float), ("one_minus_beta2", c_float)]


class AdamTensor(ctypes.Structure):
    """mmbs_adam_tensor (include/mmbs.h)."""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_i64),
                ("group", c_i32), ("reserved", c_i32)]


# name -> (restype, argtypes); every symbol include/mmbs.h declares
SIGNATURES = {
    "mmbs_last_error": (ctypes.c_char_p, []),
    "mmbs_version": (ctypes.c_int, []),
    "mmbs_launch_count": (c_i64, []),
    "mmbs_device_check": (ctypes.c_int, []),
    "mmbs_cox_workspace_bytes": (c_size, [c_i64]),
    "mmbs_cox_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_size, c_void_p]),
    "mmbs_cox_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                         c_void_p, c_void_p, c_size, c_void_p]),
    "mmbs_risk_order": (ctypes.c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_size, c_void_p]),
    "mmbs_cox_debug_state": (ctypes.c_int, [c_void_p, c_size, c_i64, c_void_p]),
    "mmbs_segmented_mean_workspace_bytes": (c_size, [c_i64, c_i64]),
    "mmbs_segmented_mean": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_size, c_void_p]),
    "mmbs_conv_plan_create": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_void_p)]),
    "mmbs_conv_plan_destroy": (None, [c_void_p]),
    "mmbs_conv_run": (ctypes.c_int, [c_void_p, c_void_p]),
    "mmbs_linear_plan_create": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64,
                                               c_i32, c_i32, ctypes.POINTER(c_void_p)]),
    "mmbs_linear_tn_plan_create": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64,
                                                  ctypes.POINTER(c_void_p)]),
    "mmbs_linear_nn_plan_create": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64,
                                                  c_i32, c_i32, ctypes.POINTER(c_void_p)]),
    "mmbs_conv_wgrad_plan_create": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                                   c_i64, ctypes.POINTER(c_void_p)]),
    "mmbs_stem_pack_input": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "mmbs_stem_pack_input_u8": (ctypes.c_int, [c_void_p, c_void_p, c_i64, ctypes.POINTER(c_float),
                                               ctypes.POINTER(c_float), c_void_p]),
    "mmbs_stem_pack_weight": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "mmbs_stem_pack_input_c": (ctypes.c_int, [c_void_p, c_void_p, c_i64, ctypes.c_int, c_void_p]),
    "mmbs_stem_pack_weight_c": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, c_void_p]),
    "mmbs_concordance_dominance": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int, c_void_p, c_void_p,
                                                  c_void_p, c_i64, c_void_p, c_void_p]),
    "mmbs_attention_pool": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int, ctypes.c_int, c_void_p,
                                           c_void_p, c_void_p, c_void_p]),
    "mmbs_tanh_inplace_f32": (ctypes.c_int, [c_void_p, c_i64, c_void_p]),
    "mmbs_png_decode": (ctypes.c_int, [c_void_p, c_size, c_void_p, ctypes.c_int, ctypes.c_int]),
    "mmbs_png_decode_files": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), c_i64, c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int]),
    "mmbs_resample_coeffs": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, ctypes.c_int]),
    "mmbs_resize_bilinear_u8": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p,
                                               ctypes.c_int, c_void_p]),
    "mmbs_augment_u8": (ctypes.c_int, [c_void_p, c_void_p, c_i64, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "mmbs_pack_conv_weight": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_bn_fold": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_i64, c_void_p,
                                    c_void_p, c_void_p]),
    "mmbs_maxpool_3x3s2": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_avgpool_global": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_avgpool_global_f32": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_cast_pad_bf16": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_dropout_cast_bf16": (ctypes.c_int, [c_void_p, c_i32, c_i64, c_void_p, c_i64, c_i64, c_i64, c_float,
                                              ctypes.c_uint64, ctypes.c_uint32, c_void_p]),
    "mmbs_mlp_bwd_elementwise": (ctypes.c_int, [c_void_p, c_i32, c_i64, c_void_p, c_i64, c_i32, c_float,
                                                ctypes.c_uint64, ctypes.c_uint32, c_i64, c_i64, c_i64, c_i64,
                                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmbs_transpose_bf16": (ctypes.c_int, [c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "mmbs_cast_transpose_pad_bf16": (ctypes.c_int, [c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "mmbs_bn_finalize": (ctypes.c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmbs_bn_apply": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p,
                                     c_i64, c_i64, c_void_p]),
    "mmbs_bn_relu_maxpool_3x3s2": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64,
                                                  c_void_p]),
    "mmbs_avgpool_global_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_bn_bwd_reduce": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64,
                                          c_void_p]),
    "mmbs_bn_bwd_apply": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_i64, c_i64, c_void_p]),
    "mmbs_im2col_t": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i32,
                                     c_void_p]),
    "mmbs_pack_conv_weight_dgrad": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_unpack_conv_wgrad": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_scatter_stride2": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "mmbs_add_relu_mask": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "mmbs_bn_train_apply": (ctypes.c_int, [ctypes.POINTER(BnTrainDesc), c_void_p, c_void_p, ctypes.POINTER(BnTrainDesc),
                                           c_i32, c_void_p, c_i64, c_void_p]),
    "mmbs_bn_train_relu_maxpool_3x3s2": (ctypes.c_int, [ctypes.POINTER(BnTrainDesc), c_void_p, c_void_p, c_i64, c_i64,
                                                        c_i64, c_void_p]),
    "mmbs_concordance_counts": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "mmbs_plan_set_dropout": (ctypes.c_int, [c_void_p, c_float, ctypes.c_uint64, ctypes.c_uint32]),
    "mmbs_nll_surv_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_float, c_float, c_i32,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmbs_nll_surv_backward": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, c_void_p, c_void_p]),
    "mmbs_write_matrix_csv": (ctypes.c_int, [c_void_p, c_i64, c_i64, ctypes.c_char_p, c_i32]),
    "mmbs_adam_step": (ctypes.c_int, [ctypes.POINTER(AdamTensor), c_i32, ctypes.POINTER(AdamGroup), c_i32, c_void_p]),
}


def build(verbose: bool = False) -> str:
    """Compile libmmbs.so in-tree with nvcc (sm_100a).  Cross-compiles without a GPU."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmmbs.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C multimodalbrainsurvival_b200/csrc` (there is no CPU/PyTorch fallback)")
            l = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


class MMBSError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mmbs_last_error().decode("utf-8", "replace")
        raise MMBSError(f"libmmbs {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().mmbs_launch_count())


def stream_ptr():
    """The current CUDA stream of the current device as an integer handle (every launch passes one: the raw C call is
    ~20x cheaper than building a torch.cuda.Stream object - 19 launches per RNA training step made that 10 % of the step)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:   # private names moved: the public (slower) route
        return torch.cuda.current_stream().cuda_stream


class on_device:
    """``with torch.cuda.device(dev)`` that costs nothing when `dev` is already current (the usual case)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        import torch
        idx = getattr(device, "index", device)
        self.idx = torch._C._cuda_getDevice() if idx is None else int(idx)
        self.prev = -1

    def __enter__(self):
        import torch
        cur = torch._C._cuda_getDevice()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            import torch
            torch.cuda.set_device(self.prev)
        return False


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
