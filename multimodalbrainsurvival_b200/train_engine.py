"""Training-mode ResNet-50 trunk on libmmbs: batch-statistics BatchNorm forward for the whole
network and the backward pass through ``layer4`` (the configuration the reference fine-tunes:
``n_layers_to_train = 2`` -> ``fc`` + ``layer4``, /root/reference/1_HistoPathology/2_HistoPath_train.py:541-551,
/root/reference/5_JointFusion/1_JointFusion_train.py:380-392).

Reference arithmetic replaced: ``ResNet.forward_extract`` under ``model.train()``
(/root/reference/5_JointFusion/resnet.py:151-165, ``Bottleneck.forward`` :70-90) and its autograd graph.

Forward, per convolution:   tcgen05 implicit GEMM writing the raw bf16 output and the per-channel
sum / sum of squares from its epilogue -> ``mmbs_bn_train_apply`` (finalize fused into the normalise + residual +
ReLU pass: scale/shift derived in-kernel, saved mean/invstd, running statistics updated in place like
``nn.BatchNorm2d``).
Backward, per layer4 conv:  BatchNorm backward (reduce + apply), data gradient = the same conv kernel
on flipped/transposed weights, weight gradient = the same GEMM kernel on K-major (pixel-major) operands
with fp32 accumulation and output.  PyTorch provides device memory, streams and the autograd hook only.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import torch

from . import _lib, engine
from ._lib import BnTrainDesc


def _pad64(n):
    return (n + 63) // 64 * 64


def _ck(rc, what):
    _lib.check(rc, what)


class _BNState:
    """Per-BatchNorm device state: epilogue sums, folded scale/shift, saved mean / invstd."""

    def __init__(self, bn, arena, off):
        c = bn.num_features
        self.bn, self.c = bn, c
        self.stats = arena[off:off + 2 * c]
        self.scale = arena[off + 2 * c:off + 3 * c]
        self.shift = arena[off + 3 * c:off + 4 * c]
        self.mean = arena[off + 4 * c:off + 5 * c]
        self.invstd = arena[off + 5 * c:off + 6 * c]

    @staticmethod
    def floats(bn):
        return 6 * bn.num_features


class _Conv:
    """One convolution of the trunk: packed bf16 weights [Cout][kh][kw][Cin], repacked in place when the parameter
    changes.  The layer4 data-gradient convolutions read the same matrix (MN-major B operand, taps flipped in the
    kernel), so no transposed copy exists."""

    def __init__(self, conv, kind):
        self.conv, self.kind = conv, kind   # kind: 'stem' | 'plain' | 'halo'
        O, I, k, _ = conv.weight.shape
        dev = conv.weight.device
        cols = 256 if kind == 'stem' else k * k * I
        self.w = torch.empty((O, cols), dtype=torch.bfloat16, device=dev)
        self.version = None

    def refresh(self):
        cw = self.conv.weight
        ver = (cw.data_ptr(), cw._version)
        if ver == self.version:
            return
        L = _lib.lib()
        w = cw.detach().float().contiguous()
        O, I, k, _ = w.shape
        if self.kind == 'stem':
            _ck(L.mmbs_stem_pack_weight(_lib.ptr(w), _lib.ptr(self.w), _lib.stream_ptr()), "mmbs_stem_pack_weight")
        elif self.kind == 'halo':
            self.w.copy_(engine.pack_conv_weight_halo(w))
        else:
            _ck(L.mmbs_pack_conv_weight(_lib.ptr(w), _lib.ptr(self.w), O, I, k, _lib.stream_ptr()),
                "mmbs_pack_conv_weight")
        self.version = ver


class ResNetTrainEngine:
    """Buffers + plans of one (ResNet-50 module, batch size) pair for training-mode steps."""

    def __init__(self, net, batch: int):
        p = next(net.parameters())
        if not p.is_cuda:
            raise RuntimeError("ResNetTrainEngine: the module must live on a CUDA device (no CPU fallback)")
        self._net_ref = weakref.ref(net)   # the module owns the engine (net._engines): no strong cycle
        self.B, self.device = int(batch), p.device
        self._keep = []
        self.fwd_steps = []
        self.convs = []
        self.bns = []
        self.step = 0
        self.use_graph = os.environ.get("MMBS_CUDA_GRAPH", "1") == "1"
        self._graphs = {}
        self._hot_convs = []
        self._grads = {}
        self.l4 = []          # per layer4 block: dict of saved tensors / states / backward plans
        with torch.cuda.device(self.device):
            self._build()

    # ------------------------------------------------------------------ construction helpers
    def _buf(self, *shape, dtype=torch.bfloat16, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def _conv(self, conv, kind='plain'):
        c = _Conv(conv, kind)
        self.convs.append(c)
        return c

    def _bn(self, bn):
        if bn.momentum is None or not bn.track_running_stats or not bn.affine:
            raise RuntimeError("ResNetTrainEngine: BatchNorm2d needs affine=True, track_running_stats=True and a "
                               "fixed momentum")
        st = _BNState(bn, self.arena, self._arena_off)
        self._arena_off += _BNState.floats(bn)
        self.bns.append(st)
        return st

    def _add_conv_bn(self, cv, st, x, raw, *, ksize, stride, c_in, in_hw=None, halo=False):
        """The convolution writing its raw bf16 output and the per-channel sums of its BatchNorm."""
        plan = engine.conv_plan(x, cv.w, raw, ksize=ksize, stride=stride, c_in=c_in, in_hw=in_hw, halo_weights=halo,
                                stats=st.stats)
        self._keep.append(plan)
        st.count = raw.shape[0] * raw.shape[1] * raw.shape[2]
        self.fwd_steps.append(plan.run)

    @staticmethod
    def _bn_desc(st):
        """mmbs_bn_train_desc of one BatchNorm (built per call: parameters may have been re-allocated)."""
        bn = st.bn
        d = BnTrainDesc()
        d.stats, d.gamma, d.beta = st.stats.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr()
        d.running_mean, d.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        d.scale_out, d.shift_out = st.scale.data_ptr(), st.shift.data_ptr()
        d.mean_out, d.invstd_out = st.mean.data_ptr(), st.invstd.data_ptr()
        d.eps, d.momentum, d.count, d.c = float(bn.eps), float(bn.momentum), st.count, st.c
        return d

    def _add_apply(self, raw, st, out, relu=True, res=None, res_st=None):
        """Fused BatchNorm finalize + normalise (+ residual) (+ ReLU): mmbs_bn_train_apply."""
        L = _lib.lib()
        rows = raw.numel() // raw.shape[-1]

        def apply():
            d = self._bn_desc(st)
            rd = self._bn_desc(res_st) if res_st is not None else None
            _ck(L.mmbs_bn_train_apply(ctypes.byref(d), _lib.ptr(raw), _lib.ptr(res),
                                      ctypes.byref(rd) if rd is not None else None, int(relu), _lib.ptr(out), rows,
                                      _lib.stream_ptr()), "mmbs_bn_train_apply")
        self.fwd_steps.append(apply)

    def _build(self):
        net, B = self._net_ref(), self.B
        L = _lib.lib()
        all_bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        self.arena = torch.zeros(sum(_BNState.floats(b) for b in all_bns), dtype=torch.float32, device=self.device)
        self._arena_off = 0
        # ---- stem
        self.x_s2d = self._buf(B, 116, 116, 16)
        cv0, st0 = self._conv(net.conv1, 'stem'), self._bn(net.bn1)
        raw0 = self._buf(B, 112, 112, 64)
        pool = self._buf(B, 56, 56, 64)
        self._add_conv_bn(cv0, st0, self.x_s2d, raw0, ksize=4, stride=1, c_in=16, in_hw=(116, 116))
        self.fwd_steps.append(lambda: _ck(L.mmbs_bn_train_relu_maxpool_3x3s2(
            ctypes.byref(self._bn_desc(st0)), _lib.ptr(raw0), _lib.ptr(pool), B, 112, 112, _lib.stream_ptr()),
            "mmbs_bn_train_relu_maxpool_3x3s2"))
        # ---- bottlenecks
        x = pool
        for li, layer in enumerate([net.layer1, net.layer2, net.layer3, net.layer4]):
            for blk in layer:
                x = self._add_block(blk, x, save=(li == 3))
        self.final = x   # [B,7,7,2048]
        self.feats = self._buf(B, 2048, dtype=torch.float32)
        self.dfeat = self._buf(B, 2048, dtype=torch.float32)
        self.g_final = self._buf(*self.final.shape)
        self._build_backward()

    def _add_block(self, blk, x, save):
        B, H, W, Cin = x.shape
        planes = blk.conv1.out_channels
        s = blk.conv2.stride[0]
        Ho, Wo = H // s, W // s
        halo = False   # the halo variant's 8x16 pixel box overhangs a 56x56 image: its tile rows would pollute the sums
        c1, c2, c3 = self._conv(blk.conv1), self._conv(blk.conv2, 'halo' if halo else 'plain'), self._conv(blk.conv3)
        b1, b2, b3 = self._bn(blk.bn1), self._bn(blk.bn2), self._bn(blk.bn3)
        raw1 = self._buf(B, H, W, planes)
        raw2 = self._buf(B, Ho, Wo, planes)
        raw3 = self._buf(B, Ho, Wo, planes * 4)
        # frozen layers normalise in place; layer4 keeps raw and activated tensors for the backward pass
        a1 = self._buf(B, H, W, planes) if save else raw1
        a2 = self._buf(B, Ho, Wo, planes) if save else raw2
        out = self._buf(B, Ho, Wo, planes * 4) if save else raw3
        rec = {"blk": blk, "x": x, "raw1": raw1, "a1": a1, "raw2": raw2, "a2": a2, "raw3": raw3, "out": out,
               "c1": c1, "c2": c2, "c3": c3, "b1": b1, "b2": b2, "b3": b3, "stride": s, "rawd": None}
        if blk.downsample is not None:
            cd, bd = self._conv(blk.downsample[0]), self._bn(blk.downsample[1])
            rawd = self._buf(B, Ho, Wo, planes * 4)
            self._add_conv_bn(cd, bd, x, rawd, ksize=1, stride=blk.downsample[0].stride[0], c_in=Cin)
            rec.update(cd=cd, bd=bd, rawd=rawd)
        self._add_conv_bn(c1, b1, x, raw1, ksize=1, stride=1, c_in=Cin)
        self._add_apply(raw1, b1, a1)
        self._add_conv_bn(c2, b2, a1, raw2, ksize=3, stride=s, c_in=planes, halo=halo)
        self._add_apply(raw2, b2, a2)
        self._add_conv_bn(c3, b3, a2, raw3, ksize=1, stride=1, c_in=planes)
        if blk.downsample is not None:
            self._add_apply(raw3, b3, out, res=rec["rawd"], res_st=rec["bd"])
        else:
            self._add_apply(raw3, b3, out, res=x)
        if save:
            self.l4.append(rec)
        return out

    # ------------------------------------------------------------------ backward construction
    def _wgrad_tn(self, dy2d, x2d):
        """dW[co, ci] = sum_p dY[p, co] X[p, ci] of a 1x1 convolution: a TN GEMM on the NHWC tensors themselves
        (MN-major tcgen05 operands, no transposed copies), split-K, fp32 output."""
        dw = self._buf(dy2d.shape[1], x2d.shape[1], dtype=torch.float32)
        plan = engine.linear_tn_plan(dy2d, x2d, dw)
        self._keep.append(plan)
        return dw, plan

    def _wgrad_conv(self, dy, x, ksize, stride):
        """Weight gradient of a 3x3 / strided convolution as an implicit TN GEMM on the NHWC tensors (K chunk = 64
        images at one output pixel, padding = TMA zero fill): no im2col buffer, OIHW fp32 output."""
        dw = self._buf(dy.shape[3], x.shape[3], ksize, ksize, dtype=torch.float32)
        plan = engine.conv_wgrad_plan(dy, x, dw, ksize=ksize, stride=stride)
        self._keep.append(plan)
        return dw, plan

    def _build_backward(self):
        B = self.B
        for bi, r in enumerate(self.l4):
            blk = r["blk"]
            x, out = r["x"], r["out"]
            _, Hin, Win, Cin = x.shape
            _, Ho, Wo, Cout = out.shape
            planes = blk.conv1.out_channels
            s = r["stride"]
            P, Pin = B * Ho * Wo, B * Hin * Win
            Pp, Pinp = _pad64(P), _pad64(Pin)
            need_dx = bi > 0
            bf = torch.bfloat16
            # conv3 / bn3
            r["sums3"] = self._buf(2, Cout, dtype=torch.float32)
            r["draw3"] = self._buf(B, Ho, Wo, Cout)
            r["dw3"], r["wg3"] = self._wgrad_tn(r["draw3"].view(P, Cout), r["a2"].view(P, planes))
            r["da2"] = self._buf(B, Ho, Wo, planes)
            # data gradients read the forward conv's packed weights (MN-major B operand, taps flipped in the kernel)
            r["dg3"] = engine.conv_plan(r["draw3"], r["c3"].w, r["da2"], ksize=1, stride=1, c_in=Cout, fwd_weights=True)
            # conv2 / bn2
            r["sums2"] = self._buf(2, planes, dtype=torch.float32)
            r["draw2"] = self._buf(B, Ho, Wo, planes)
            r["dw2"], r["wg2"] = self._wgrad_conv(r["draw2"], r["a1"], 3, s)
            r["u2"] = self._buf(B, Hin, Win, planes, zero=True) if s == 2 else r["draw2"]
            r["da1"] = self._buf(B, Hin, Win, planes)
            r["dg2"] = engine.conv_plan(r["u2"], r["c2"].w, r["da1"], ksize=3, stride=1, c_in=planes, fwd_weights=True)
            # conv1 / bn1
            r["sums1"] = self._buf(2, planes, dtype=torch.float32)
            r["draw1"] = self._buf(B, Hin, Win, planes)
            r["dw1"], r["wg1"] = self._wgrad_tn(r["draw1"].view(Pin, planes), x.view(Pin, Cin))
            if need_dx:
                r["dx1"] = self._buf(B, Hin, Win, Cin)
                r["dg1"] = engine.conv_plan(r["draw1"], r["c1"].w, r["dx1"], ksize=1, stride=1, c_in=planes,
                                            fwd_weights=True)
                r["g_in"] = self._buf(B, Hin, Win, Cin)
            if r["rawd"] is not None:
                r["sumsd"] = self._buf(2, Cout, dtype=torch.float32)
                r["drawd"] = self._buf(B, Ho, Wo, Cout)
                r["dwd"], r["wgd"] = self._wgrad_conv(r["drawd"], x, 1, blk.downsample[0].stride[0])
            r["dims"] = (Hin, Win, Cin, Ho, Wo, Cout, planes, P, Pin, Pp, Pinp)

    # ------------------------------------------------------------------ execution
    def _graph_key(self):
        """Pointers the captured graphs bake in: if a parameter / buffer was re-allocated, capture again."""
        key = []
        for st in self.bns:
            bn = st.bn
            key += [bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                    bn.num_batches_tracked.data_ptr()]
        key += [c.conv.weight.data_ptr() for c in self.convs]
        return tuple(key)

    def _run(self, which, body):
        """Eager on the first call (kernel attributes are set lazily), captured into a CUDA graph on the second,
        replayed afterwards (one graph launch instead of ~130 ctypes calls: the step is CPU-launch bound
        whenever something synchronises the host, e.g. reading the loss)."""
        st = self._graphs.setdefault(which, {"runs": 0, "graph": None, "key": None, "kernels": 0})
        st["runs"] += 1
        if not self.use_graph or st["runs"] == 1:
            body()
            return
        key = self._graph_key()
        if st["graph"] is None or st["key"] != key:
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                body()
            st["graph"], st["key"], st["kernels"] = g, key, _lib.launch_count() - l0
        st["graph"].replay()
        engine.GRAPH_LAUNCHES += st["kernels"]   # libmmbs kernels inside the replayed graph (bench.py's gpu_launches)

    def _fwd_body(self):
        L = _lib.lib()
        for c in self._hot_convs:    # trainable convolutions: their weights change every step
            c.version = None
            c.refresh()
        self.arena.zero_()   # the epilogue sums accumulate; scale/shift/mean/invstd are rewritten by the finalize steps
        for step in self.fwd_steps:
            step()
        _ck(L.mmbs_avgpool_global(_lib.ptr(self.final), _lib.ptr(self.feats), self.B, 49, 2048, _lib.stream_ptr()),
            "mmbs_avgpool_global")
        torch._foreach_add_([st.bn.num_batches_tracked for st in self.bns], 1)

    def forward(self, x_nchw: torch.Tensor) -> torch.Tensor:
        """x_nchw fp32 [B,3,224,224] -> features fp32 [B,2048]; running statistics are updated."""
        L = _lib.lib()
        hot = [c for c in self.convs if c.conv.weight.requires_grad]
        if [id(c) for c in hot] != [id(c) for c in self._hot_convs]:
            self._hot_convs = hot
            self._graphs.pop("fwd", None)
        for c in self.convs:
            if not c.conv.weight.requires_grad:
                c.refresh()          # frozen weights: repacked only when their version counter moves
        self.step += 1
        # the kernels (and graph replays) write the running statistics behind autograd's back: tell the
        # inference engine's weight cache (engine.ResNetEngine.weights_version) that they moved
        net = self._net_ref()
        if net is not None:
            net._mmbs_train_steps = getattr(net, "_mmbs_train_steps", 0) + 1
        if x_nchw.dtype == torch.uint8:   # raw pixels (e.g. pipeline.augment output): ToTensor + Normalize in the pack kernel
            import ctypes
            mean = (ctypes.c_float * 3)(*(net.input_mean if net is not None else (0.485, 0.456, 0.406)))
            std = (ctypes.c_float * 3)(*(net.input_std if net is not None else (0.229, 0.224, 0.225)))
            _ck(L.mmbs_stem_pack_input_u8(_lib.ptr(x_nchw), _lib.ptr(self.x_s2d), self.B, mean, std, _lib.stream_ptr()),
                "mmbs_stem_pack_input_u8")
        else:
            _ck(L.mmbs_stem_pack_input(_lib.ptr(x_nchw), _lib.ptr(self.x_s2d), self.B, _lib.stream_ptr()),
                "mmbs_stem_pack_input")
        self._run("fwd", self._fwd_body)
        return self.feats.clone()

    def _bn_backward(self, g, mask, raw, st, sums, out):
        L = _lib.lib()
        rows, c = raw.numel() // raw.shape[-1], raw.shape[-1]
        sums.zero_()
        _ck(L.mmbs_bn_bwd_reduce(_lib.ptr(g), _lib.ptr(mask), _lib.ptr(raw), _lib.ptr(st.mean), _lib.ptr(st.invstd),
                                 _lib.ptr(sums), rows, c, _lib.stream_ptr()), "mmbs_bn_bwd_reduce")
        _ck(L.mmbs_bn_bwd_apply(_lib.ptr(g), _lib.ptr(mask), _lib.ptr(raw), _lib.ptr(st.mean), _lib.ptr(st.invstd),
                                _lib.ptr(st.scale), _lib.ptr(sums), _lib.ptr(out), rows, c, _lib.stream_ptr()),
            "mmbs_bn_bwd_apply")

    def backward(self, dfeat: torch.Tensor) -> dict:
        """dfeat fp32 [B,2048] -> {parameter: gradient tensor (engine-owned, valid until the next step)}."""
        self.dfeat.copy_(dfeat.detach())
        self._run("bwd", self._bwd_body)
        return self._grads

    def _bwd_body(self):
        L = _lib.lib()
        B = self.B
        _ck(L.mmbs_avgpool_global_bwd(_lib.ptr(self.dfeat), _lib.ptr(self.g_final), B, 49, 2048, _lib.stream_ptr()),
            "mmbs_avgpool_global_bwd")
        grads = {}
        g = self.g_final
        for bi in range(len(self.l4) - 1, -1, -1):
            r = self.l4[bi]
            blk = r["blk"]
            Hin, Win, Cin, Ho, Wo, Cout, planes, P, Pin, Pp, Pinp = r["dims"]
            # ---- bn3 / conv3 (the block's ReLU mask comes from its output)
            self._bn_backward(g, r["out"], r["raw3"], r["b3"], r["sums3"], r["draw3"])
            r["wg3"].run()
            r["dg3"].run()
            grads[blk.conv3.weight] = r["dw3"].view(Cout, planes, 1, 1)
            grads[blk.bn3.weight], grads[blk.bn3.bias] = r["sums3"][1], r["sums3"][0]
            # ---- bn2 / conv2
            self._bn_backward(r["da2"], r["a2"], r["raw2"], r["b2"], r["sums2"], r["draw2"])
            r["wg2"].run()
            if r["stride"] == 2:
                _ck(L.mmbs_scatter_stride2(_lib.ptr(r["draw2"]), _lib.ptr(r["u2"]), B, Ho, Wo, planes,
                                           _lib.stream_ptr()), "mmbs_scatter_stride2")
            r["dg2"].run()
            grads[blk.conv2.weight] = r["dw2"]
            grads[blk.bn2.weight], grads[blk.bn2.bias] = r["sums2"][1], r["sums2"][0]
            # ---- bn1 / conv1
            self._bn_backward(r["da1"], r["a1"], r["raw1"], r["b1"], r["sums1"], r["draw1"])
            r["wg1"].run()
            grads[blk.conv1.weight] = r["dw1"].view(planes, Cin, 1, 1)
            grads[blk.bn1.weight], grads[blk.bn1.bias] = r["sums1"][1], r["sums1"][0]
            # ---- downsample branch (first block) / identity shortcut
            if r["rawd"] is not None:
                self._bn_backward(g, r["out"], r["rawd"], r["bd"], r["sumsd"], r["drawd"])
                r["wgd"].run()
                grads[blk.downsample[0].weight] = r["dwd"]
                grads[blk.downsample[1].weight], grads[blk.downsample[1].bias] = r["sumsd"][1], r["sumsd"][0]
            if bi > 0:
                r["dg1"].run()
                _ck(L.mmbs_add_relu_mask(_lib.ptr(r["dx1"]), _lib.ptr(g), _lib.ptr(r["out"]), _lib.ptr(r["g_in"]),
                                         r["g_in"].numel(), _lib.stream_ptr()), "mmbs_add_relu_mask")
                g = r["g_in"]
        self._grads = grads


# ---------------------------------------------------------------------------------- autograd glue
class _TrunkTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, *params):
        ctx.eng = eng
        ctx.params = params
        out = eng.forward(x)
        ctx.step = eng.step
        return out

    @staticmethod
    def backward(ctx, dfeat):
        if ctx.eng.step != ctx.step:
            raise RuntimeError("ResNet training engine: backward() of a step whose saved activations were overwritten "
                               "by a later forward of the same batch size (one outstanding forward per engine)")
        grads = ctx.eng.backward(dfeat)
        out = []
        for p, need in zip(ctx.params, ctx.needs_input_grad[2:]):
            out.append(grads[p].clone().view_as(p) if need else None)   # engine buffers are reused next step
        return (None, None) + tuple(out)


def trainable_outside_layer4(net) -> bool:
    """True when a parameter the kernels do not differentiate (stem .. layer3) requires a gradient."""
    mods = [net.conv1, net.bn1, net.layer1, net.layer2, net.layer3]
    return any(p.requires_grad for m in mods for p in m.parameters())


def run_train(net, x: torch.Tensor, engines: dict) -> torch.Tensor:
    """Training-mode ``forward_extract`` of ``net`` on fp32 (normalised) or uint8 (raw pixels) NCHW ``x`` (no gradient
    into x)."""
    B = x.shape[0]
    key = ("train", x.device.index, B)
    eng = engines.get(key)
    if eng is None:
        # keep two batch sizes (the regular one and the tail batch of an epoch): the buffers are large
        old = [k for k in engines if isinstance(k, tuple) and k and k[0] == "train"]
        for k in old[:-1]:
            engines.pop(k)
        eng = ResNetTrainEngine(net, B)
    else:
        engines.pop(key)                   # re-insert: most recently used last
    engines[key] = eng
    params = [p for p in net.layer4.parameters()]
    x = x.detach().contiguous() if x.dtype == torch.uint8 else x.detach().float().contiguous()
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _TrunkTrainFn.apply(x, eng, *params)
    return eng.forward(x)
