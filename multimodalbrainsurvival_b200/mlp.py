"""Fused execution of the reference's inline MLPs on the tcgen05 GEMM kernel.

The reference builds its MLPs from stock ``nn.Linear`` inside the scripts
(2_GeneExpression/1_GeneExpress_train.py:247-257, 3_EarlyFusion/2_EarlyFusion_train.py:242-253,
5_JointFusion/1_JointFusion_train.py:314-325), so this module only ever *receives* an
``nn.Sequential``; it pattern-matches ``Dropout / Linear / ReLU`` stacks and runs them as
GEMMs with the bias and ReLU fused into the epilogue, reading the ``nn.Linear`` parameters in
place (bf16 padded copies are caches keyed on the parameters' version counters).

Served by the kernels: inference AND training on CUDA tensors.  Training goes through
``_MLPTrainFn``: forward GEMMs with bias/ReLU epilogues, Philox dropout regenerated (not
stored) in the backward pass, dgrad/wgrad as tcgen05 GEMMs on transposed bf16 operands with
fp32 accumulation; the fp32 master weights stay in the ``nn.Linear`` parameters (what Adam sees).
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.nn as nn

from . import engine


def _pad(n, m):
    return (n + m - 1) // m * m


def _parse(seq, with_dropout=False):
    """-> list of (linear, relu_after[, dropout_p_before]) or None when the stack is not a plain MLP."""
    if not isinstance(seq, nn.Sequential):
        return None
    layers = []
    pending_p = 0.0
    for m in seq:
        if isinstance(m, nn.Dropout):
            if pending_p > 0.0 or m.p >= 1.0:
                return None
            pending_p = float(m.p)
        elif isinstance(m, nn.Linear):
            layers.append([m, False, pending_p])
            pending_p = 0.0
        elif isinstance(m, nn.ReLU) and layers and not layers[-1][1] and pending_p == 0.0:
            layers[-1][1] = True
        else:
            return None
    if not layers or pending_p > 0.0:
        return None
    return layers if with_dropout else [l[:2] for l in layers]


class _MLPEngine:
    def __init__(self, layers, m_rows, device):
        self.layers = layers
        self.m = m_rows
        self.device = device
        self.k0 = layers[0][0].in_features
        with torch.cuda.device(device):
            self.x_bf = torch.empty((m_rows, _pad(self.k0, 64)), dtype=torch.bfloat16, device=device)
            self.w_bf, self.b_pad, self.outs, self.plans = [], [], [], []
            cur = self.x_bf
            for i, (lin, relu) in enumerate(layers):
                last = i == len(layers) - 1
                kp = cur.shape[1]
                npad = _pad(lin.out_features, 32 if last else 64)
                w = torch.zeros((npad, kp), dtype=torch.bfloat16, device=device)
                b = torch.zeros(npad, dtype=torch.float32, device=device)
                out = torch.empty((m_rows, npad), dtype=torch.float32 if last else torch.bfloat16, device=device)
                self.w_bf.append(w)
                self.b_pad.append(b)
                self.outs.append(out)
                self.plans.append(engine.linear_plan(cur, w, b, out, relu=relu))
                cur = out
        self.version = None

    def _weights_version(self):
        v = []
        for lin, _ in self.layers:
            v.append((lin.weight.data_ptr(), lin.weight._version))
            if lin.bias is not None:
                v.append((lin.bias.data_ptr(), lin.bias._version))
        return tuple(v)

    def refresh(self):
        ver = self._weights_version()
        if ver == self.version:
            return
        for (lin, _), w, b in zip(self.layers, self.w_bf, self.b_pad):
            n = lin.out_features
            engine.cast_pad_bf16(lin.weight, w.shape[1], out=w[:n])
            if lin.bias is not None:
                b[:n].copy_(lin.bias.detach())
        self.version = ver

    def run(self, x):
        self.refresh()
        engine.cast_pad_bf16(x, self.x_bf.shape[1], out=self.x_bf)
        for p in self.plans:
            p.run()
        return self.outs[-1][:, :self.layers[-1][0].out_features].clone()




# ------------------------------------------------------------------------------- training
def _pad64(n):
    return _pad(n, 64)


class _MLPTrainEngine:
    """Buffers + GEMM plans of one MLP for a fixed batch size M (forward, dgrad, wgrad)."""

    def __init__(self, layers, m_rows, device, need_dx, world=1, group=None):
        from . import _lib
        self.layers = layers
        self.world, self.group = int(world), group
        self.m, self.mp = m_rows, _pad64(m_rows)
        self.device = device
        self.need_dx = need_dx
        L = len(layers)
        self.kp = [_pad64(l[0].in_features) for l in layers]
        self.np_ = [_pad64(l[0].out_features) for l in layers]
        for i in range(1, L):
            if self.kp[i] != self.np_[i - 1]:
                raise ValueError("layer widths do not chain")
        bf, f32 = torch.bfloat16, torch.float32
        M, Mp = self.m, self.mp
        with torch.cuda.device(device):
            E = lambda *s, dtype=bf: torch.zeros(s, dtype=dtype, device=device)  # noqa: E731
            self.hin = [E(M, self.kp[0])]                      # dropped inputs of every layer
            self.w, self.b, self.act = [], [], []
            self.dz, self.dW, self.db, self.dh = [], [], [], []
            for i, (lin, relu, p) in enumerate(layers):
                last = i == L - 1
                self.w.append(E(self.np_[i], self.kp[i]))
                self.b.append(E(self.np_[i], dtype=f32))
                self.act.append(E(M, self.np_[i], dtype=f32 if last else bf))
                if not last:
                    # the dropout in front of the NEXT Linear is applied by THIS layer's GEMM epilogue
                    # (mmbs_plan_set_dropout): the layer writes the next layer's input directly
                    self.hin.append(self.act[i])
                self.dz.append(E(M, self.np_[i]))
                self.dW.append(E(self.np_[i], self.kp[i], dtype=f32))
                self.db.append(E(self.np_[i], dtype=f32))
                self.dh.append(E(M, self.kp[i]) if (i > 0 or need_dx) else None)
            self.fwd = [engine.linear_plan(self.hin[i], self.w[i], self.b[i], self.act[i], relu=layers[i][1])
                        for i in range(L)]
            # dW[n, k] = sum_m dz[m, n] hin[m, k]: a TN GEMM on the row-major tensors themselves (MN-major operands).
            # Data parallel (dist.enable_factored_mlp_gradients): where the two factors of all ranks are smaller than
            # the gradient itself, they are all-gathered and the GEMM runs over the GLOBAL batch - the result is the
            # rank-summed gradient, nothing is left to all-reduce for that weight.
            self.dz_all, self.hin_all = [None] * L, [None] * L
            for i in range(L):
                factors = self.world * M * (self.np_[i] + self.kp[i]) * 2          # bf16 bytes gathered per rank
                # vs the fp32 gradient (whose all-reduce moves it twice: reduce-scatter + all-gather)
                if self.world > 1 and factors < self.np_[i] * self.kp[i] * 4:
                    self.dz_all[i] = E(self.world * M, self.np_[i])
                    self.hin_all[i] = E(self.world * M, self.kp[i])
            self.wgrad = [engine.linear_tn_plan(self.dz[i] if self.dz_all[i] is None else self.dz_all[i],
                                                self.hin[i] if self.hin_all[i] is None else self.hin_all[i], self.dW[i])
                          for i in range(L)]
            # dh[m, k] = sum_n dz[m, n] W[n, k]: an NN GEMM on the forward weight matrix itself (MN-major B operand)
            self.dgrad = [engine.linear_nn_plan(self.dz[i], self.w[i], None, self.dh[i]) if self.dh[i] is not None
                          else None for i in range(L)]
        self.version = None
        self._lib = _lib
        # The activations / dropout probabilities / gradient buffers of ONE forward live in this engine.  `step`
        # counts forwards; the autograd node that owns the current contents is tracked by weak reference, so that a
        # second forward while the first one still awaits its backward (two micro-batches summed before backward, a
        # siamese call) gets ANOTHER engine instead of overwriting the saved state (see _run_train).
        self.step = 0
        self._owner = None
        self._backward_done = True

    def busy(self):
        """True while a forward's autograd node is alive and has not run its backward yet."""
        owner = self._owner() if self._owner is not None else None
        return owner is not None and not self._backward_done

    def _weights_version(self):
        v = []
        for lin, _, _ in self.layers:
            v.append((lin.weight.data_ptr(), lin.weight._version))
            if lin.bias is not None:
                v.append((lin.bias.data_ptr(), lin.bias._version))
        return tuple(v)

    def refresh(self):
        ver = self._weights_version()
        if ver == self.version:
            return
        _lib = self._lib
        L = _lib.lib()
        for i, (lin, _, _) in enumerate(self.layers):
            n, k = lin.out_features, lin.in_features
            wsrc = lin.weight.detach().float().contiguous()
            engine.cast_pad_bf16(wsrc, self.kp[i], out=self.w[i][:n])
            if lin.bias is not None:
                self.b[i][:n].copy_(lin.bias.detach())
        self.version = ver

    def forward(self, x, seed, training):
        _lib = self._lib
        L = _lib.lib()
        self.refresh()
        x = x.detach()
        x = x.contiguous() if x.dtype == torch.bfloat16 else x.float().contiguous()   # bf16 cohorts are read as is
        ps = [(l[2] if training else 0.0) for l in self.layers]
        self.ps = ps
        _lib.check(L.mmbs_dropout_cast_bf16(_lib.ptr(x), int(x.dtype == torch.bfloat16), x.shape[1],
                                            _lib.ptr(self.hin[0]), self.m, x.shape[1],
                                            self.kp[0], ps[0], seed, 0, _lib.stream_ptr()), "mmbs_dropout_cast_bf16")
        for i in range(len(self.layers)):
            # bias + ReLU + the dropout of layer i+1's input in one epilogue (north_star (2)); mask (seed, tag i+1) is
            # the one mmbs_mlp_bwd_elementwise regenerates.  act[i] then holds the DROPPED activations: the backward's
            # ReLU test (act > 0) is unaffected - a dropped element carries no gradient either way.
            p_next = ps[i + 1] if i + 1 < len(self.layers) else 0.0
            _lib.check(L.mmbs_plan_set_dropout(self.fwd[i]._h, float(p_next), seed, i + 1), "mmbs_plan_set_dropout")
            self.fwd[i].run()
        return self.act[-1][:, :self.layers[-1][0].out_features].clone()

    def backward(self, dy, seed):
        """dy: fp32 [M, N_last].  Returns (dx or None, [dW_i views], [db_i views])."""
        _lib = self._lib
        L = _lib.lib()
        g, g_bf16, g_stride = dy.detach().float().contiguous(), 0, dy.shape[1]
        n_layers = len(self.layers)
        # 1. the chain dz[L-1] -> dgrad -> dz[L-2] -> ... (the weight gradients hang off it, nothing waits for them)
        for i in range(n_layers - 1, -1, -1):
            lin, relu, _ = self.layers[i]
            p_after = self.ps[i + 1] if i + 1 < n_layers else 0.0   # dropout applied to this layer's output
            self.db[i].zero_()
            _lib.check(L.mmbs_mlp_bwd_elementwise(_lib.ptr(g), g_bf16, g_stride, _lib.ptr(self.act[i]), self.np_[i],
                                                  int(relu), p_after, seed, i + 1, self.m, lin.out_features,
                                                  self.np_[i], self.mp, _lib.ptr(self.dz[i]), None,
                                                  _lib.ptr(self.db[i]), _lib.stream_ptr()), "mmbs_mlp_bwd_elementwise")
            if self.dgrad[i] is not None:
                self.dgrad[i].run()
                g, g_bf16, g_stride = self.dh[i], 1, self.kp[i]
        # 2. data parallel: ONE coalesced all-gather of the (dz, h) factors of every factored layer
        pairs = [(self.dz_all[i], self.dz[i]) for i in range(n_layers) if self.dz_all[i] is not None]
        pairs += [(self.hin_all[i], self.hin[i]) for i in range(n_layers) if self.hin_all[i] is not None]
        if pairs:
            import torch.distributed as tdist
            done = False
            if tdist.get_backend(self.group) == "nccl" and hasattr(tdist, "_coalescing_manager"):
                try:
                    with tdist._coalescing_manager(group=self.group, device=pairs[0][0].device, async_ops=False):
                        for out_t, in_t in pairs:
                            tdist.all_gather_into_tensor(out_t, in_t, group=self.group)
                    done = True
                except TypeError:   # private API signature moved (raised before any collective)
                    pass
            if not done:
                for out_t, in_t in pairs:
                    tdist.all_gather_into_tensor(out_t, in_t, group=self.group)
        # 3. weight gradients (over the global batch where the factors were gathered)
        for i in range(n_layers):
            self.wgrad[i].run()
        dx = None
        if self.need_dx:
            k0 = self.layers[0][0].in_features
            if self.ps[0] > 0:   # gradient through the input dropout: same Philox mask as the forward
                _lib.check(L.mmbs_dropout_cast_bf16(_lib.ptr(self.dh[0]), 1, self.kp[0], _lib.ptr(self.dh[0]), self.m,
                                                    self.kp[0], self.kp[0], self.ps[0], seed, 0, _lib.stream_ptr()),
                           "mmbs_dropout_cast_bf16")
            dx = self.dh[0][:, :k0].float()
        dWs = [self.dW[i][:l[0].out_features, :l[0].in_features] for i, l in enumerate(self.layers)]
        dbs = [self.db[i][:l[0].out_features] for i, l in enumerate(self.layers)]
        return dx, dWs, dbs


_TRAIN_ENGINES = {}      # key -> [engines]; more than one only while several forwards await their backward
_MAX_ENGINES_PER_KEY = 8


class _MLPTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, training, *params):
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # CPU generator: torch.manual_seed reproducible
        ctx.eng, ctx.seed = eng, seed
        ctx.n_params = len(params)
        ctx.has_bias = [l[0].bias is not None for l in eng.layers]
        eng.step += 1
        ctx.step = eng.step
        eng._owner, eng._backward_done = weakref.ref(ctx), False
        return eng.forward(x, seed, training)

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        if eng.step != ctx.step:
            raise RuntimeError("fused MLP backward: the activations saved by this forward were overwritten by a later "
                               "forward of the same module (a second backward through a retained graph after the "
                               "engine was reused); call backward once per forward, or set MMBS_MLP_TRAIN=0")
        eng._backward_done = True
        dx, dWs, dbs = eng.backward(dy, ctx.seed)
        grads = []
        for i, hb in enumerate(ctx.has_bias):
            grads.append(dWs[i].clone())   # the engine's buffers are reused by the next step
            if hb:
                grads.append(dbs[i].clone())
        return (dx, None, None) + tuple(grads)


def _run_train(seq, layers, x):
    from . import dist as _dist
    need_dx = bool(x.requires_grad)
    world, group = _dist.factored_mlp_world()
    key = (id(seq), x.shape[0], x.device.index, need_dx, world)
    pool = _TRAIN_ENGINES.get(key)
    if pool is None or any(a[0] is not b[0] for a, b in zip(pool[0].layers, layers)):
        if len(_TRAIN_ENGINES) > 32:
            _TRAIN_ENGINES.clear()
        pool = _TRAIN_ENGINES[key] = []
    eng = next((e for e in pool if not e.busy()), None)
    if eng is None:   # every engine of this module still holds the saved state of a forward awaiting its backward
        if len(pool) >= _MAX_ENGINES_PER_KEY:
            raise RuntimeError(f"fused MLP: {len(pool)} forwards of the same module are waiting for their backward; "
                               "run backward (or drop the outputs) before calling it again, or set MMBS_MLP_TRAIN=0")
        eng = _MLPTrainEngine(layers, x.shape[0], x.device, need_dx, world, group)
        pool.append(eng)
    params = []
    for i, (lin, _, _) in enumerate(layers):
        # (allreduce_gradients skips weights whose gradient already is the sum over the ranks)
        setattr(lin.weight, _dist.GLOBAL_GRAD_ATTR, eng.dz_all[i] is not None)
        params.append(lin.weight)
        if lin.bias is not None:
            params.append(lin.bias)
    return _MLPTrainFn.apply(x, eng, seq.training, *params)


_ENGINES = {}


def _fusable(seq, x):
    if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1":
        return None
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2):
        return None
    if torch.is_grad_enabled() and (seq.training or x.requires_grad or any(p.requires_grad for p in seq.parameters())):
        return None
    layers = _parse(seq)
    if layers is None or layers[0][0].in_features != x.shape[1]:
        return None
    return layers


def _trainable(seq, x):
    """Plain MLP on a CUDA 2-D input with autograd active -> the fused training path."""
    if os.environ.get("MMBS_DISABLE_KERNELS", "0") == "1" or os.environ.get("MMBS_MLP_TRAIN", "1") != "1":
        return None
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and torch.is_grad_enabled()):
        return None
    layers = _parse(seq, with_dropout=True)
    if layers is None or layers[0][0].in_features != x.shape[1] or layers[-1][1]:
        return None
    for i in range(1, len(layers)):
        if _pad(layers[i][0].in_features, 64) != _pad(layers[i - 1][0].out_features, 64):
            return None
    return layers


def run_mlp(seq, x):
    """Evaluate ``seq(x)`` through the fused kernels when the stack is a plain MLP."""
    layers = _fusable(seq, x)
    if layers is None:
        tl = _trainable(seq, x)
        if tl is not None:
            return _run_train(seq, tl, x)
        return nn.Sequential.forward(seq, x) if isinstance(seq, AcceleratedSequential) else seq(x)
    key = (id(seq), x.shape[0], x.device.index)
    eng = _ENGINES.get(key)
    if eng is None or any(a is not b[0] for a, b in zip([l[0] for l in eng.layers], layers)):
        eng = _MLPEngine(layers, x.shape[0], x.device)
        if len(_ENGINES) > 64:
            _ENGINES.clear()
        _ENGINES[key] = eng
    return eng.run(x)


class AcceleratedSequential(nn.Sequential):
    """``nn.Sequential`` whose forward goes through :func:`run_mlp` (same children/keys)."""

    @classmethod
    def wrap(cls, seq):
        if isinstance(seq, cls):
            return seq
        if not isinstance(seq, nn.Sequential):
            raise TypeError("accelerate() expects an nn.Sequential MLP")
        new = cls()
        for name, child in seq.named_children():
            new.add_module(name, child)
        new.train(seq.training)
        return new

    def forward(self, x):
        return run_mlp(self, x)
