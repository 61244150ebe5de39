"""GPU numerics: the tcgen05 implicit-GEMM kernel (csrc/gemm_tcgen05.cu) vs a plain torch
fp32 reference of the same op on bf16-rounded operands.  Tolerance: bf16 output rounding
(2^-8 relative) + fp32 accumulation-order noise."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(x):
    return x.to(torch.bfloat16).float()


def _ref_conv(x_nhwc, w_oihw, scale, shift, res_nhwc, relu, stride, pad):
    x = _bf(x_nhwc).permute(0, 3, 1, 2).double().cpu()
    w = _bf(w_oihw).double().cpu()
    y = F.conv2d(x, w, stride=stride, padding=pad)
    y = y * scale.double().cpu().view(1, -1, 1, 1) + shift.double().cpu().view(1, -1, 1, 1)
    if res_nhwc is not None:
        y = y + _bf(res_nhwc).permute(0, 3, 1, 2).double().cpu()
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).float()


def _assert_close(got, ref, tol, name):
    got = got.float().cpu()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-6, f"{name}: max err {err:.4g} vs scale {scale:.4g}"


CASES = [
    # B, H, W, Cin, Cout, k, stride, residual, relu, out_f32
    (2, 56, 56, 64, 64, 1, 1, False, True, False),
    (2, 56, 56, 64, 256, 1, 1, True, True, False),
    (2, 56, 56, 64, 64, 3, 1, False, True, False),
    (3, 56, 56, 128, 128, 3, 2, False, True, False),
    (2, 56, 56, 256, 512, 1, 2, False, False, False),
    (4, 28, 28, 128, 128, 3, 1, False, True, False),
    (5, 14, 14, 256, 256, 3, 1, False, True, False),
    (3, 14, 14, 512, 512, 3, 2, False, True, False),
    (130, 7, 7, 512, 512, 3, 1, False, True, False),
    (2, 7, 7, 512, 2048, 1, 1, True, True, True),
    (2, 14, 14, 1024, 256, 1, 1, False, True, False),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B{}_{}x{}_c{}-{}_k{}s{}".format(*c[:7]))
def test_conv_matches_torch(case):
    from multimodalbrainsurvival_b200 import engine
    B, H, W, Cin, Cout, k, s, use_res, relu, out_f32 = case
    torch.manual_seed(hash(case) % 1000)
    x = torch.randn(B, H, W, Cin, device=DEV)
    w = torch.randn(Cout, Cin, k, k, device=DEV) / (k * k * Cin) ** 0.5
    scale = torch.rand(Cout, device=DEV) + 0.5
    shift = torch.randn(Cout, device=DEV) * 0.1
    Ho, Wo = H // s, W // s
    res = torch.randn(B, Ho, Wo, Cout, device=DEV).to(torch.bfloat16) if use_res else None
    xb = x.to(torch.bfloat16).contiguous()
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device=DEV,
                     dtype=torch.float32 if out_f32 else torch.bfloat16)
    plan = engine.conv_plan(xb, engine.pack_conv_weight(w), out, ksize=k, stride=s, c_in=Cin, scale=scale,
                            shift=shift, residual=res, relu=relu)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, scale, shift, res.float() if use_res else None, relu, s, k // 2)
    _assert_close(out, ref, 1e-5 if out_f32 else 6e-3, str(case))


@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 64), (20, 56, 56, 64), (3, 24, 40, 64)])
def test_conv3x3_halo_variant_matches_torch(B, H, W, C):
    """3x3/1 conv with 64 output channels: weights resident in smem, one halo box per kw."""
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(B + H)
    x = torch.randn(B, H, W, C, device=DEV)
    w = torch.randn(64, C, 3, 3, device=DEV) / (9 * C) ** 0.5
    scale = torch.rand(64, device=DEV) + 0.5
    shift = torch.randn(64, device=DEV) * 0.1
    out = torch.full((B, H, W, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    conv = torch.nn.Conv2d(C, 64, 3, padding=1, bias=False)
    assert engine.halo_eligible(conv)
    engine.conv_plan(x.to(torch.bfloat16).contiguous(), engine.pack_conv_weight_halo(w), out, ksize=3, stride=1,
                     c_in=C, scale=scale, shift=shift, relu=True, halo_weights=True).run()
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, scale, shift, None, True, 1, 1)
    _assert_close(out, ref, 6e-3, f"halo {B}x{H}x{W}")


@pytest.mark.parametrize("M,N,K,relu", [(128, 4096, 12800, True), (6, 2048, 4096, False), (300, 224, 2048, True),
                                        (1000, 32, 64, False)])
def test_linear_matches_torch(M, N, K, relu):
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(M + N)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV)
    xb, wb = x.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous()
    out = torch.full((M, N), float("nan"), device=DEV)
    engine.linear_plan(xb, wb, b, out, relu=relu).run()
    torch.cuda.synchronize()
    ref = xb.double() @ wb.double().t() + b.double()
    if relu:
        ref = torch.relu(ref)
    _assert_close(out, ref.float().cpu(), 2e-5, f"linear {M}x{N}x{K}")


def test_stem_matches_torch():
    """7x7/2 stem conv + BN + ReLU through the space-to-depth GEMM, then MaxPool(3,2,1)."""
    from multimodalbrainsurvival_b200 import _lib, engine
    torch.manual_seed(3)
    B = 3
    x = torch.randn(B, 3, 224, 224, device=DEV)
    w = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
    scale = torch.rand(64, device=DEV) + 0.5
    shift = torch.randn(64, device=DEV) * 0.1
    L = _lib.lib()
    s2d = torch.empty(B, 116, 116, 16, device=DEV, dtype=torch.bfloat16)
    _lib.check(L.mmbs_stem_pack_input(_lib.ptr(x), _lib.ptr(s2d), B, _lib.stream_ptr()))
    out = torch.full((B, 112, 112, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    engine.conv_plan(s2d, engine.pack_stem_weight(w), out, ksize=4, stride=1, c_in=16, scale=scale, shift=shift,
                     relu=True, in_hw=(116, 116)).run()
    pooled = torch.empty(B, 56, 56, 64, device=DEV, dtype=torch.bfloat16)
    _lib.check(L.mmbs_maxpool_3x3s2(_lib.ptr(out), _lib.ptr(pooled), B, 112, 112, 64, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv2d(_bf(x).double().cpu(), _bf(w).double().cpu(), stride=2, padding=3)
    ref = torch.relu(ref * scale.double().cpu().view(1, -1, 1, 1) + shift.double().cpu().view(1, -1, 1, 1))
    _assert_close(out, ref.permute(0, 2, 3, 1).float(), 6e-3, "stem")
    ref_pool = F.max_pool2d(out.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).cpu()
    assert torch.equal(pooled.float().cpu(), ref_pool)


@pytest.mark.parametrize("m,n,k", [(256, 512, 6400), (32, 256, 64 * 333), (2048, 512, 6272 + 0), (128, 4608, 64 * 7),
                                    (512, 1024, 64 * 5)])
def test_split_k_weight_gradient_gemm(m, n, k):
    """fp32-output GEMMs without epilogue arithmetic (the weight-gradient shape: few output tiles, long K) take
    the split-K path: partial sums are accumulated with fp32 reductions into the cleared output; repeated runs must
    not accumulate across calls."""
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
    b = torch.randn(n, k, device=DEV).to(torch.bfloat16)
    out = torch.full((m, n), 7.0, dtype=torch.float32, device=DEV)     # stale contents must be discarded
    plan = engine.linear_plan(a, b, None, out)
    plan.run()
    plan.run()
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    err = float((out.double() - ref).abs().max())
    assert err <= 2e-5 * float(ref.abs().max()) + 1e-3, f"max err {err}"


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (512, 256, 6272), (2048, 512, 6272), (512, 1024, 25088), (64, 64, 100),
                                    (4096, 12800, 128), (256, 2048, 333)])
def test_tn_gemm_mn_major_operands(m, n, k):
    """y[M, N] = a[K, M]^T b[K, N] with both operands row-major (MN-major tcgen05 descriptors): the weight-gradient
    GEMM straight from [pixels, channels] tensors.  K need not be a multiple of 64 (TMA zero-fills the tail)."""
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(m + n + k)
    a = torch.randn(k, m, device=DEV).to(torch.bfloat16)
    b = torch.randn(k, n, device=DEV).to(torch.bfloat16)
    out = torch.full((m, n), -3.0, dtype=torch.float32, device=DEV)
    plan = engine.linear_tn_plan(a, b, out)
    plan.run()
    plan.run()
    torch.cuda.synchronize()
    ref = a.double().t() @ b.double()
    err = float((out.double() - ref).abs().max())
    assert err <= 2e-5 * float(ref.abs().max()) + 1e-3, f"max err {err}"


@pytest.mark.parametrize("m,n,k,relu,f32", [(128, 4096, 2048, False, False), (300, 64, 256, True, True),
                                            (1024, 2048, 256, False, False)])
def test_nn_gemm_mn_major_weights(m, n, k, relu, f32):
    """y = act(x[M, K] @ w[K, N] + bias): the B operand is a row-major [K, N] matrix fed MN-major (the data
    gradient of a linear layer on its forward weights)."""
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(m + n + k)
    x = torch.randn(m, k, device=DEV).to(torch.bfloat16)
    w = (torch.randn(k, n, device=DEV) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(n, device=DEV)
    out = torch.empty(m, n, device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
    engine.linear_nn_plan(x, w, bias, out, relu=relu).run()
    torch.cuda.synchronize()
    ref = x.double() @ w.double() + bias.double()
    if relu:
        ref = torch.relu(ref)
    err = float((out.double() - ref).abs().max())
    assert err <= (1e-5 if f32 else 1e-2) * float(ref.abs().max()) + 1e-4, f"max err {err}"


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s", [(5, 7, 7, 512, 512, 3, 1), (70, 14, 14, 512, 512, 3, 2),
                                                (128, 7, 7, 256, 128, 3, 1), (9, 14, 14, 1024, 2048, 1, 2),
                                                (3, 8, 8, 64, 64, 3, 1)])
def test_conv_weight_gradient_implicit_gemm(B, H, W, Cin, Cout, k, s):
    """mmbs_conv_wgrad_plan_create vs torch autograd (fp64 on the bf16-rounded operands): OIHW layout, padding
    through TMA zero fill, stride 2, image counts that are not multiples of the 64-image K chunk."""
    from multimodalbrainsurvival_b200 import engine
    torch.manual_seed(B + H + Cin + Cout + k + s)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
    dy = torch.randn(B, Ho, Wo, Cout, device=DEV).to(torch.bfloat16)
    dw = torch.full((Cout, Cin, k, k), 5.0, device=DEV)
    plan = engine.conv_wgrad_plan(dy, x, dw, ksize=k, stride=s)
    plan.run()
    plan.run()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (Cout, Cin, k, k), dy.permute(0, 3, 1, 2).double(),
                                      stride=s, padding=k // 2)
    err = float((dw.double() - ref).abs().max())
    assert err <= 2e-5 * float(ref.abs().max()) + 1e-3, f"max err {err} (scale {float(ref.abs().max())})"


def test_cta_pair_variant_matches_torch():
    """MMBS_CLUSTER=1 (CTA pairs sharing the weight boxes through TMA multicast) is read once per process:
    run the conv parity cases again in a child process with the variant switched on."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, MMBS_CLUSTER="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_gemm.py"), "-q", "-x", "-k",
                        "conv and not cta_pair", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True,
                       timeout=600, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
