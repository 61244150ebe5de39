"""Host-side execution engine over libmmbs: conv/linear plans and the ResNet-50
inference pipeline (NHWC bf16 activations, folded eval BatchNorm, fused ReLU/residual).

Replaces the arithmetic of ``ResNet.forward_extract``
(/root/reference/5_JointFusion/resnet.py:151-165 == 1_HistoPathology/resnet.py:151-165)
while reading the weights from the stock ``nn.Conv2d`` / ``nn.BatchNorm2d`` parameters
(state_dict stays byte-compatible).  PyTorch is used for device
memory and streams only.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import ConvDesc, c_void_p


class ConvPlan:
    """Owns one mmbs_conv_plan (TMA descriptors bound to fixed device buffers)."""

    def __init__(self, handle, keep):
        self._h = handle
        self._keep = keep  # tensors whose storage the descriptors point at

    def run(self):
        _lib.check(_lib.lib().mmbs_conv_run(self._h, _lib.stream_ptr()), "mmbs_conv_run")

    def __del__(self):
        try:
            if self._h:
                _lib.lib().mmbs_conv_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def conv_plan(x, w, out, *, ksize, stride, c_in, scale=None, shift=None, residual=None, relu=False,
              in_hw=None, halo_weights=False, stats=None, fwd_weights=False):
    """x: bf16 NHWC [B,H,W,Cin] (ksize 4 = stem: the [B,116,116,16] space-to-depth buffer);
    w: bf16 [Cout, k*k*Cin]; out: NHWC bf16 or fp32.
    fwd_weights=True: a data-gradient convolution on the FORWARD conv's packed weights (w is the forward
    [Cin_of_this_conv, k*k*Cout_of_this_conv] matrix; taps flipped in the kernel, MN-major B operand)."""
    B = x.shape[0]
    H, W = (x.shape[1], x.shape[2]) if in_hw is None else in_hw
    d = ConvDesc()
    c_out = w.shape[1] // (ksize * ksize) if fwd_weights else w.shape[0]
    d.batch, d.in_h, d.in_w, d.c_in, d.c_out = B, H, W, c_in, c_out
    d.ksize, d.stride, d.relu = ksize, stride, int(relu)
    d.out_f32 = int(out.dtype == torch.float32)
    d.flags = (1 if halo_weights else 0) | (2 if fwd_weights else 0)
    d.in_ = x.data_ptr()
    d.weight = w.data_ptr()
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.out = out.data_ptr()
    d.stats = stats.data_ptr() if stats is not None else None
    h = c_void_p()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmbs_conv_plan_create(ctypes.byref(d), ctypes.byref(h)), "mmbs_conv_plan_create")
    return ConvPlan(h, (x, w, out, scale, shift, residual, stats))


def linear_plan(x, w, bias, out, relu=False):
    """x bf16 [M,K], w bf16 [N,K] (K % 64 == 0, N % 32 == 0), out bf16/fp32 [M,N]."""
    M, K = x.shape
    N = w.shape[0]
    h = c_void_p()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmbs_linear_plan_create(
            _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), M, N, K, int(relu),
            int(out.dtype == torch.float32), ctypes.byref(h)), "mmbs_linear_plan_create")
    return ConvPlan(h, (x, w, bias, out))


def linear_nn_plan(x, w_kn, bias, out, relu=False):
    """out [M, N] = x[M, K] @ w_kn[K, N] (+ bias, ReLU); w_kn bf16 row-major (K outermost), K % 64 == 0, N % 64 == 0."""
    M, K = x.shape
    N = w_kn.shape[1]
    assert w_kn.shape[0] == K and tuple(out.shape) == (M, N)
    h = c_void_p()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmbs_linear_nn_plan_create(
            _lib.ptr(x), _lib.ptr(w_kn), _lib.ptr(bias), _lib.ptr(out), M, N, K, int(relu),
            int(out.dtype == torch.float32), ctypes.byref(h)), "mmbs_linear_nn_plan_create")
    return ConvPlan(h, (x, w_kn, bias, out))


def conv_wgrad_plan(dy, x, dw, *, ksize, stride):
    """dw fp32 OIHW [Cout, Cin, k, k] = weight gradient of conv(x, w, stride, pad=k//2) given dy; dy / x bf16 NHWC."""
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    assert tuple(dw.shape) == (Cout, Cin, ksize, ksize) and dw.dtype == torch.float32 and dw.is_contiguous()
    h = c_void_p()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmbs_conv_wgrad_plan_create(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), B, H, W, Cin, Cout,
                                                          ksize, stride, ctypes.byref(h)), "mmbs_conv_wgrad_plan_create")
    return ConvPlan(h, (dy, x, dw))


def linear_tn_plan(a_km, b_kn, out):
    """out fp32 [M, N] = a_km[K, M]^T @ b_kn[K, N]; a / b bf16 row-major (K outermost), M % 8 == 0, N % 64 == 0."""
    K, M = a_km.shape
    N = b_kn.shape[1]
    assert b_kn.shape[0] == K and tuple(out.shape) == (M, N) and out.dtype == torch.float32
    h = c_void_p()
    with torch.cuda.device(a_km.device):
        _lib.check(_lib.lib().mmbs_linear_tn_plan_create(_lib.ptr(a_km), _lib.ptr(b_kn), _lib.ptr(out), M, N, K,
                                                         ctypes.byref(h)), "mmbs_linear_tn_plan_create")
    return ConvPlan(h, (a_km, b_kn, out))


# ------------------------------------------------------------------ small wrappers
def pack_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    O, I, k, _ = w_oihw.shape
    w = w_oihw.detach().float().contiguous()
    out = torch.empty((O, k * k * I), dtype=torch.bfloat16, device=w.device)
    _lib.check(_lib.lib().mmbs_pack_conv_weight(_lib.ptr(w), _lib.ptr(out), O, I, k, _lib.stream_ptr()),
               "mmbs_pack_conv_weight")
    return out


def halo_eligible(conv: torch.nn.Conv2d) -> bool:
    """3x3 / stride 1 / 64 output channels with <= 72 KB of bf16 weights: the kernel keeps the
    weights resident in shared memory and reads each input row once per kw (halo box)."""
    return (conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.out_channels == 64
            and conv.in_channels % 64 == 0 and 9 * conv.in_channels * 64 * 2 <= 72 * 1024)


def pack_conv_weight_halo(w_oihw: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> bf16 [O][kw][I/64][kh][64] (the K order the halo variant iterates in)."""
    O, I, kh, kw = w_oihw.shape
    w = w_oihw.detach().float().reshape(O, I // 64, 64, kh, kw).permute(0, 4, 1, 3, 2).contiguous()
    return w.reshape(O, kh * kw * I).to(torch.bfloat16).contiguous()


def pack_stem_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    c = int(w_oihw.shape[1])
    assert tuple(w_oihw.shape) == (64, c, 7, 7) and c in (1, 3, 4), "stem conv must be Conv2d(1|3|4, 64, 7, stride 2, pad 3)"
    w = w_oihw.detach().float().contiguous()
    out = torch.empty((64, 256), dtype=torch.bfloat16, device=w.device)
    if c == 3:
        _lib.check(_lib.lib().mmbs_stem_pack_weight(_lib.ptr(w), _lib.ptr(out), _lib.stream_ptr()),
                   "mmbs_stem_pack_weight")
    else:   # RNone / RNfour
        _lib.check(_lib.lib().mmbs_stem_pack_weight_c(_lib.ptr(w), _lib.ptr(out), c, _lib.stream_ptr()),
                   "mmbs_stem_pack_weight_c")
    return out


def bn_fold(bn: torch.nn.BatchNorm2d):
    c = bn.num_features
    dev = bn.running_mean.device
    g = (bn.weight if bn.weight is not None else torch.ones(c, device=dev)).detach().float().contiguous()
    b = (bn.bias if bn.bias is not None else torch.zeros(c, device=dev)).detach().float().contiguous()
    m = bn.running_mean.detach().float().contiguous()
    v = bn.running_var.detach().float().contiguous()
    out = torch.empty((2, c), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().mmbs_bn_fold(_lib.ptr(g), _lib.ptr(b), _lib.ptr(m), _lib.ptr(v), float(bn.eps), c,
                                       _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.stream_ptr()), "mmbs_bn_fold")
    return out[0], out[1]


def cast_pad_bf16(x: torch.Tensor, cols_padded: int, out: torch.Tensor | None = None) -> torch.Tensor:
    x = x.detach().float().contiguous()
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols_padded), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().mmbs_cast_pad_bf16(_lib.ptr(x), _lib.ptr(out), rows, cols, cols_padded,
                                             _lib.stream_ptr()), "mmbs_cast_pad_bf16")
    return out


# ------------------------------------------------------------------ ResNet-50 pipeline
def front_chunk(chunk: int) -> int:
    """Sub-chunk for stem .. layer2 (big activations).  Small sub-chunks keep a bottleneck's
    tensors inside the 126 MB L2, but measured on B200 (profiles/r01_chunk_sweep.md) the extra
    launches cost more than the saved HBM traffic, so the default is the whole chunk."""
    f = int(os.environ.get("MMBS_RESNET_FRONT_CHUNK", "0")) or chunk
    f = max(1, min(f, chunk))
    return f if chunk % f == 0 else chunk


class ResNetEngine:
    """Inference pipeline for one (ResNet module, chunk size).

    A chunk of `chunk` patches is packed once, then stem / maxpool / layer1 / layer2 run in
    sub-chunks of `front` patches on reused scratch buffers (L2-resident), writing layer2's output
    for the whole chunk; layer3 / layer4 (small activations, compute-bound) run once on the whole
    chunk so their grids fill the 148 SMs.  Everything between the input pack and the final
    average pool is one CUDA graph."""

    def __init__(self, resnet, chunk: int):
        p = next(resnet.parameters())
        if not p.is_cuda:
            raise RuntimeError("ResNetEngine: the module must live on a CUDA device (no CPU fallback)")
        self.device = p.device
        self.chunk = int(chunk)
        self.front = front_chunk(self.chunk)
        self._steps = []
        self._keep = []
        self._weights_version = None
        self._graph = None
        self._runs = 0
        self.use_graph = os.environ.get("MMBS_CUDA_GRAPH", "1") == "1"
        self._build(resnet)

    @staticmethod
    def weights_version(resnet):
        return (tuple((t.data_ptr(), t._version) for t in list(resnet.parameters()) + list(resnet.buffers()))
                + (getattr(resnet, "_mmbs_train_steps", 0),))

    def _buf(self, *shape, dtype=torch.bfloat16):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def _plan(self, *a, **k):
        pl = conv_plan(*a, **k)
        self._keep.append(pl)
        return pl

    @staticmethod
    def _is_basic(blk):
        return not hasattr(blk, "conv3")   # BasicBlock (ResNet-18/34): two 3x3 convs; Bottleneck: 1x1 -> 3x3 -> 1x1

    @classmethod
    def _out_channels(cls, blk):
        return blk.conv2.out_channels if cls._is_basic(blk) else blk.conv3.out_channels

    @classmethod
    def _block_stride(cls, blk):
        return blk.conv1.stride[0] if cls._is_basic(blk) else blk.conv2.stride[0]

    def _block_steps(self, blk, w, x, out):
        """Steps of one residual block on input buffer x writing `out` (scratch allocated here)."""
        B, H, W, Cin = x.shape
        planes = blk.conv1.out_channels
        s = self._block_stride(blk)
        Ho, Wo = H // s, W // s
        steps = []
        if blk.downsample is not None:
            res = self._buf(B, Ho, Wo, self._out_channels(blk))
            steps.append(self._plan(x, w["wd"], res, ksize=1, stride=blk.downsample[0].stride[0], c_in=Cin,
                                    scale=w["scd"], shift=w["shd"], relu=False).run)
        else:
            res = x
        if self._is_basic(blk):
            # BasicBlock.forward (reference resnet.py:23-51): 3x3 (stride) -> BN -> ReLU -> 3x3 -> BN -> (+shortcut) -> ReLU
            t1 = self._buf(B, Ho, Wo, planes)
            steps.append(self._plan(x, w["w1"], t1, ksize=3, stride=s, c_in=Cin, scale=w["sc1"], shift=w["sh1"],
                                    relu=True, halo_weights=w["halo1"]).run)
            steps.append(self._plan(t1, w["w2"], out, ksize=3, stride=1, c_in=planes, scale=w["sc2"], shift=w["sh2"],
                                    residual=res, relu=True, halo_weights=w["halo2"]).run)
            return steps
        t1 = self._buf(B, H, W, planes)
        t2 = self._buf(B, Ho, Wo, planes)
        steps.append(self._plan(x, w["w1"], t1, ksize=1, stride=1, c_in=Cin, scale=w["sc1"], shift=w["sh1"],
                                relu=True).run)
        steps.append(self._plan(t1, w["w2"], t2, ksize=3, stride=s, c_in=planes, scale=w["sc2"], shift=w["sh2"],
                                relu=True, halo_weights=w["halo2"]).run)
        steps.append(self._plan(t2, w["w3"], out, ksize=1, stride=1, c_in=planes, scale=w["sc3"], shift=w["sh3"],
                                residual=res, relu=True).run)
        return steps

    @classmethod
    def _block_weights(cls, blk):
        use_halo = os.environ.get("MMBS_CONV_HALO", "1") == "1"
        w = {}
        if cls._is_basic(blk):
            # (the residual epilogue of the weights-resident halo variant is not instantiated: conv2 streams its weights)
            w["halo1"] = halo_eligible(blk.conv1) and use_halo
            w["halo2"] = False
            w["w1"] = pack_conv_weight_halo(blk.conv1.weight) if w["halo1"] else pack_conv_weight(blk.conv1.weight)
            w["w2"] = pack_conv_weight(blk.conv2.weight)
        else:
            w["halo2"] = halo_eligible(blk.conv2) and use_halo
            w["w1"] = pack_conv_weight(blk.conv1.weight)
            w["w2"] = pack_conv_weight_halo(blk.conv2.weight) if w["halo2"] else pack_conv_weight(blk.conv2.weight)
            w["w3"] = pack_conv_weight(blk.conv3.weight)
            w["sc3"], w["sh3"] = bn_fold(blk.bn3)
        w["sc1"], w["sh1"] = bn_fold(blk.bn1)
        w["sc2"], w["sh2"] = bn_fold(blk.bn2)
        if blk.downsample is not None:
            w["wd"] = pack_conv_weight(blk.downsample[0].weight)
            w["scd"], w["shd"] = bn_fold(blk.downsample[1])
        return w

    def _build(self, net):
        B, F = self.chunk, self.front
        L = _lib.lib()
        with torch.cuda.device(self.device):
            w_stem = pack_stem_weight(net.conv1.weight)
            sc0, sh0 = bn_fold(net.bn1)
            front_blocks = list(net.layer1) + list(net.layer2)
            back_blocks = list(net.layer3) + list(net.layer4)
            fw = [self._block_weights(b) for b in front_blocks]
            bw = [self._block_weights(b) for b in back_blocks]
            self._keep += [w_stem, sc0, sh0, fw, bw]

            self.x_s2d = self._buf(B, 116, 116, 16)
            mid = self._buf(B, 28, 28, self._out_channels(front_blocks[-1]))  # layer2 output, whole chunk
            # ---- front: per sub-chunk, on scratch buffers shared by all sub-chunks
            stem_out = self._buf(F, 112, 112, 64)
            pool_out = self._buf(F, 56, 56, 64)
            scratch_steps = None
            for j in range(B // F):
                xs = self.x_s2d[j * F:(j + 1) * F]
                self._steps.append(self._plan(xs, w_stem, stem_out, ksize=4, stride=1, c_in=16, scale=sc0, shift=sh0,
                                              relu=True, in_hw=(116, 116)).run)
                self._steps.append(lambda: _lib.check(L.mmbs_maxpool_3x3s2(
                    _lib.ptr(stem_out), _lib.ptr(pool_out), F, 112, 112, 64, _lib.stream_ptr()), "mmbs_maxpool_3x3s2"))
                if scratch_steps is None:  # plans between pool_out and the last front block are sub-chunk independent
                    scratch_steps = []
                    x = pool_out
                    for blk, w in zip(front_blocks[:-1], fw[:-1]):
                        s = self._block_stride(blk)
                        out = self._buf(F, x.shape[1] // s, x.shape[2] // s, self._out_channels(blk))
                        scratch_steps += self._block_steps(blk, w, x, out)
                        x = out
                    self._front_last_in = x
                self._steps += scratch_steps
                self._steps += self._block_steps(front_blocks[-1], fw[-1], self._front_last_in, mid[j * F:(j + 1) * F])
            # ---- back: whole chunk
            x = mid
            for bi, (blk, w) in enumerate(zip(back_blocks, bw)):
                last = bi == len(back_blocks) - 1
                s = self._block_stride(blk)
                # the final block's output stays bf16 like every other activation (TMA-store epilogue); the
                # average pool accumulates the 49 positions in fp32 (MMBS_RESNET_FINAL_F32=1: fp32 output)
                final_f32 = last and os.environ.get("MMBS_RESNET_FINAL_F32", "0") == "1"
                out = self._buf(B, x.shape[1] // s, x.shape[2] // s, self._out_channels(blk),
                                dtype=torch.float32 if final_f32 else torch.bfloat16)
                self._steps += self._block_steps(blk, w, x, out)
                x = out
            self.final = x  # [B,7,7,2048] (512 for the BasicBlock depths)
        self._weights_version = self.weights_version(net)
        self.n_kernels = len(self._steps) + 2

    def run_chunk(self, x_nchw: torch.Tensor, out: torch.Tensor,
                  norm=((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        """x_nchw: fp32 (already normalised) or uint8 (raw pixels) [chunk,3,224,224] contiguous;
        out: fp32 [chunk,2048]."""
        L = _lib.lib()
        B = self.chunk
        if x_nchw.dtype == torch.uint8:   # raw pixels: ToTensor()+Normalize() fused into the pack kernel
            m = (ctypes.c_float * 3)(*norm[0])
            sd = (ctypes.c_float * 3)(*norm[1])
            _lib.check(L.mmbs_stem_pack_input_u8(_lib.ptr(x_nchw), _lib.ptr(self.x_s2d), B, m, sd, _lib.stream_ptr()),
                       "mmbs_stem_pack_input_u8")
        elif x_nchw.shape[1] != 3:   # RNone / RNfour stems
            _lib.check(L.mmbs_stem_pack_input_c(_lib.ptr(x_nchw), _lib.ptr(self.x_s2d), B, int(x_nchw.shape[1]),
                                                _lib.stream_ptr()), "mmbs_stem_pack_input_c")
        else:
            _lib.check(L.mmbs_stem_pack_input(_lib.ptr(x_nchw), _lib.ptr(self.x_s2d), B, _lib.stream_ptr()),
                       "mmbs_stem_pack_input")
        self._run_body()
        if self.final.dtype == torch.float32:
            _lib.check(L.mmbs_avgpool_global_f32(_lib.ptr(self.final), _lib.ptr(out), B, 49, self.final.shape[3],
                                                 _lib.stream_ptr()), "mmbs_avgpool_global_f32")
        else:
            _lib.check(L.mmbs_avgpool_global(_lib.ptr(self.final), _lib.ptr(out), B, 49, self.final.shape[3],
                                             _lib.stream_ptr()), "mmbs_avgpool_global")


GRAPH_LAUNCHES = 0  # kernels launched through CUDA-graph replays (not seen by mmbs_launch_count)


def _engine_run_body(self):
    """Everything between the input pack and the final pool: eager on the first call (kernel
    attributes are set lazily), captured into a CUDA graph on the second, replayed afterwards."""
    self._runs += 1
    if not self.use_graph or self._runs == 1:
        for step in self._steps:
            step()
        return
    global GRAPH_LAUNCHES
    GRAPH_LAUNCHES += len(self._steps)
    if self._graph is None:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for step in self._steps:
                step()
        self._graph = g
    self._graph.replay()


ResNetEngine._run_body = _engine_run_body


def default_chunk(batch: int) -> int:
    env = os.environ.get("MMBS_RESNET_CHUNK")
    if env:
        return max(1, min(int(env), batch))
    return min(batch, 512)
