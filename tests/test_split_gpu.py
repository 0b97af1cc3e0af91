"""fp32 mode on the tensor cores (STFB_BF16X3, include/stfb200.h; csrc/split.cu): every fp32 GEMM operand as three bf16 planes,
six tcgen05 products per MAC, the five correction terms accumulated BEFORE the hi*hi chain (the TMEM accumulator rounds toward zero
once per K = 16 instruction).  Checked against torch in fp64 (the exact answer) with the FFMA family -- what the fp32 mode ran on
before -- measured beside it on the same inputs.  Measured on B200: forward 3.6e-7 (K = 576) .. 1.5e-6 (K = 2304) against
3.0e-7 .. 5.5e-7 for FFMA; weight gradients 0.6e-7 .. 2.4e-7 against 0.9e-7 .. 3.0e-7.  Bars: TOL = 4e-6 rel-L2 on a forward
layer, 4 x TOL on dgrad / wgrad (north_star's bar for whole-model logits is 1e-4, held by tests/test_models_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from stf_unet_b200 import ops  # noqa: E402

DEV = "cuda"
TOL = 4e-6


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(y):
    return y.permute(0, 3, 1, 2).contiguous()


def test_split_planes_sum_to_the_fp32_value():
    x = rnd(3, 5, 7, 64, seed=1) * torch.logspace(-12, 12, 64, device=DEV)       # 24 decades of magnitude
    x[0, 0, 0, :4] = torch.tensor([0.0, -0.0, 1.0, -3.5], device=DEV)
    s = ops.split_bf16x3(x)
    assert s.dtype == torch.bfloat16 and s.shape == (3, 5, 7, 192)
    hi, mid, lo = s[..., :64].double(), s[..., 64:128].double(), s[..., 128:].double()
    assert torch.equal(s[..., :64], x.to(torch.bfloat16))                        # hi = round-to-nearest bf16
    err = (hi + mid + lo - x.double()).abs()
    assert (err <= x.double().abs() * 2.0 ** -24).all()
    assert (mid.abs() <= hi.abs() * 2.0 ** -8 + 1e-300).all() and (lo.abs() <= hi.abs() * 2.0 ** -16 + 1e-300).all()


SPLIT_CASES = [
    # N, H, W, C1, C2, Cout, k, stride
    (2, 16, 16, 64, 0, 64, 3, 1),       # halo kernel, CTA pairs need >= 16 pairs: single-CTA halo here
    (40, 32, 32, 64, 0, 64, 3, 1),      # halo on CTA pairs (cta_group::2), 6 channel blocks per tap set
    (4, 24, 20, 128, 0, 256, 3, 1),     # BN = 256
    (2, 16, 16, 64, 64, 64, 3, 1),      # concat conv (UNet decoder): two operands, six segments each
    (3, 20, 24, 96, 0, 160, 3, 1),      # 32-channel k-blocks (BK = 32): segments of 96 channels
    (2, 16, 16, 128, 0, 64, 1, 1),      # 1x1, streaming kernel
    (2, 17, 19, 64, 0, 128, 3, 2),      # stride 2, ragged
    (24, 8, 8, 256, 0, 512, 3, 1),      # streaming 3x3 on small maps
]


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_conv_split_forward_is_fp32_accurate(case):
    N, H, W, C1, C2, Cout, k, s = case
    pad = (k - 1) // 2
    Cin = C1 + C2
    x = rnd(N, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, k, k, seed=2, scale=(1.0 / (k * k * Cin) ** 0.5))
    bias = rnd(Cout, seed=3)
    exact = F.conv2d(x.double(), w.double(), bias.double(), s, pad)
    x1, x2 = nhwc(x[:, :C1]), (nhwc(x[:, C1:]) if C2 else None)
    assert ops.tcgen05_ok(x1, Cout, k, s, pad, x2=x2, as_split=True)
    wp = ops.pack_weight_split(w.contiguous(), True)
    assert wp.shape == (Cout, k * k * 6 * Cin) and wp.dtype == torch.bfloat16
    y = ops.conv2d(ops.split_bf16x3(x1), wp, Cout, k, s, pad, x2=None if x2 is None else ops.split_bf16x3(x2), bias=bias,
                   impl=ops.IMPL_TCGEN05, split=True)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32
    e_split = rel(nchw(y), exact)
    # the FFMA family on the same inputs: the split path may not be (meaningfully) further from the exact answer
    y_simt = ops.conv2d(x1, ops.pack_weight(w.contiguous(), True, torch.float32), Cout, k, s, pad, x2=x2, bias=bias)
    e_simt = rel(nchw(y_simt), exact)
    print(f"conv split {case}: bf16x3 {e_split:.2e}  ffma {e_simt:.2e}")
    assert e_split < TOL, (e_split, e_simt)


def test_conv_split_epilogue_residual_relu():
    N, H, W, C, Cout = 3, 16, 16, 64, 128
    x, w = rnd(N, C, H, W, seed=1), rnd(Cout, C, 3, 3, seed=2, scale=0.05)
    scale, shift, res = rnd(Cout, seed=4).abs() + 0.5, rnd(Cout, seed=5), rnd(N, Cout, H, W, seed=6)
    exact = F.relu(F.conv2d(x.double(), w.double(), None, 1, 1) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
                   + res.double())
    y = ops.conv2d(ops.split_bf16x3(nhwc(x)), ops.pack_weight_split(w.contiguous(), True), Cout, 3, 1, 1, scale=scale, shift=shift,
                   residual=nhwc(res), relu=True, impl=ops.IMPL_TCGEN05, split=True)
    assert rel(nchw(y), exact) < TOL


@pytest.mark.parametrize("case", [(2, 16, 16, 128, 64, 3, 1), (2, 8, 8, 256, 128, 2, 2), (2, 9, 9, 64, 64, 3, 2), (6, 32, 32, 64, 64, 3, 1)])
def test_conv_split_dgrad_and_transposed(case):
    """Conv2d dgrad = transposed-gather conv over dy; ConvTranspose2d forward is the same kernel mode."""
    N, H, W, Cin, Cout, k, s = case
    pad = 1 if k == 3 else 0
    x = rnd(N, Cin, H, W, seed=1).double().requires_grad_(True)
    w = rnd(Cout, Cin, k, k, seed=2, scale=0.05)
    yref = F.conv2d(x, w.double(), None, s, pad)
    dy = rnd(*yref.shape, seed=3)
    yref.backward(dy.double())
    wpd = ops.pack_weight_split(w.contiguous(), False)                  # [n = ci][(tap, 6 x co)]
    assert wpd.shape == (Cin, k * k * 6 * Cout)
    base = rnd(N, H, W, Cin, seed=4)                                     # accumulate into an existing gradient (residual = out)
    g = base.clone()
    ops.conv2d(ops.split_bf16x3(nhwc(dy)), wpd, Cin, k, s, pad, mode=ops.CONV_TRANSPOSED, out_hw=(H, W), residual=g, out=g,
               impl=ops.IMPL_TCGEN05, split=True)
    assert rel(nchw(g - base), x.grad) < 4 * TOL
    # ConvTranspose2d forward with the [in, out, k, k] parameter layout
    wt = rnd(Cout, Cin, k, k, seed=5, scale=0.05)                        # maps Cout-channel dy-like input to Cin channels
    op = H - ((yref.shape[2] - 1) * s - 2 * pad + k)
    exact = F.conv_transpose2d(dy.double(), wt.double(), None, stride=s, padding=pad, output_padding=op)
    y = ops.conv2d(ops.split_bf16x3(nhwc(dy)), ops.pack_weight_split(wt.contiguous(), False), Cin, k, s, pad,
                   mode=ops.CONV_TRANSPOSED, out_hw=(H, W), impl=ops.IMPL_TCGEN05, split=True)
    assert rel(nchw(y), exact) < TOL


WG_CASES = [
    # N, H, W, Cin, Cout, k, stride
    (2, 16, 16, 64, 64, 3, 1),        # halo wgrad kernel
    (8, 32, 32, 128, 128, 3, 1),      # halo, split-K over many patches, partial-tile scratch
    (3, 8, 8, 128, 256, 3, 1),        # BN = 256 (maps of 8 x 8: halo too)
    (2, 4, 4, 512, 512, 3, 1),        # streaming kernel (maps below 8 x 8)
    (4, 16, 16, 256, 64, 1, 1),       # 1x1
    (2, 16, 16, 64, 128, 3, 2),       # stride 2
]


@pytest.mark.parametrize("case", WG_CASES)
def test_wgrad_split_is_fp32_accurate(case):
    N, H, W, Cin, Cout, k, s = case
    pad = (k - 1) // 2
    x = rnd(N, Cin, H, W, seed=1)
    w = torch.zeros(Cout, Cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), w, None, s, pad)
    dy = rnd(*y.shape, seed=2)
    y.backward(dy.double())
    xs, dys = ops.split_bf16x3(nhwc(x)), ops.split_bf16x3(nhwc(dy))
    assert ops.wgrad_tcgen05_ok(dys, xs, k, s, pad, split=True)
    dW = torch.full((Cout, Cin, k, k), 0.25, device=DEV)          # accumulates into what is there
    ops.conv2d_wgrad(dys, xs, dW, k, s, pad, 0, Cin, impl=ops.IMPL_TCGEN05, split=True)
    torch.cuda.synchronize()
    e_split = rel(dW - 0.25, w.grad)
    dW2 = torch.zeros(Cout, Cin, k, k, device=DEV)
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), dW2, k, s, pad, 0, Cin, impl=ops.IMPL_SIMT)
    e_simt = rel(dW2, w.grad)
    print(f"wgrad split {case}: bf16x3 {e_split:.2e}  ffma {e_simt:.2e}")
    assert e_split < 4 * TOL, (e_split, e_simt)
    # channel-window form (concat convs: one launch per source) + deferred accumulation buffer
    if Cin >= 128:
        h = Cin // 2
        acc = torch.zeros(k * k * Cin, Cout, device=DEV)
        for off, sl in ((0, slice(0, h)), (h, slice(h, Cin))):
            assert ops.conv2d_wgrad(dys, ops.split_bf16x3(nhwc(x[:, sl])), None, k, s, pad, off, Cin, acc=acc, split=True)
        got = acc.view(k, k, Cin, Cout).permute(3, 2, 0, 1)
        assert rel(got, w.grad) < 4 * TOL


def test_wgrad_split_conv_transpose():
    N, H, W, Cin, Cout, k = 2, 8, 8, 128, 64, 2
    x = rnd(N, Cin, H, W, seed=1)
    w = torch.zeros(Cin, Cout, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x.double(), w, None, stride=2)
    dy = rnd(*y.shape, seed=2)
    y.backward(dy.double())
    dW = torch.zeros(Cin, Cout, k, k, device=DEV)
    ops.conv2d_wgrad(ops.split_bf16x3(nhwc(x)), ops.split_bf16x3(nhwc(dy)), dW, k, 2, 0, 0, Cout, impl=ops.IMPL_TCGEN05, split=True)
    assert rel(dW, w.grad) < 4 * TOL


def test_split_operands_are_refused_by_the_simt_family():
    x = ops.split_bf16x3(rnd(1, 8, 8, 64, seed=1))
    wp = ops.pack_weight_split(rnd(64, 64, 3, 3, seed=2), True)
    from stf_unet_b200 import _lib
    p = _lib.ConvParams(x=x.data_ptr(), w=wp.data_ptr(), y=torch.empty(1, 8, 8, 64, device=DEV).data_ptr(), N=1, H=8, W=8, C1=64,
                        C2=0, Ho=8, Wo=8, Cout=64, kh=3, kw=3, stride=1, pad=1, ldw=64, mode=0, x_dtype=_lib.BF16X3, y_dtype=_lib.F32,
                        impl=_lib.IMPL_SIMT)
    import ctypes as C
    assert _lib.load().stfb_conv2d(C.byref(p), None) != 0
    assert b"tcgen05" in _lib.load().stfb_last_error()
