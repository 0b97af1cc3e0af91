"""GPU parity of every libstfb200 operator against a plain fp32 torch reference of the same op
(TF32 disabled).  Tolerances: fp32 path rel-L2 <= 2e-5 (north_star bar for logits is 1e-4); bf16 path
rel-L2 <= 1e-2 against the fp32 reference evaluated on bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from stf_unet_b200 import ops  # noqa: E402

DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(y):
    return y.float().permute(0, 3, 1, 2).contiguous()


def tol(dtype):
    return 2e-5 if dtype == torch.float32 else 1e-2


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def q(x, dtype):
    """round through the storage dtype so the fp32 reference sees what the kernel sees"""
    return x.to(dtype).float()


CONV_CASES = [
    # N, H, W, C1, C2, Cout, k, stride, pad
    (2, 16, 16, 64, 0, 64, 3, 1, 1),
    (2, 9, 11, 64, 0, 128, 3, 2, 1),      # ragged, stride 2
    (1, 8, 8, 128, 0, 256, 1, 2, 0),      # 1x1 downsample
    (3, 20, 20, 1, 0, 64, 7, 2, 3),       # stem (slow A path)
    (2, 12, 12, 32, 32, 32, 1, 1, 0),     # dual source fusion
    (2, 8, 8, 64, 64, 64, 3, 1, 1),       # UNet concat conv
    (2, 10, 10, 32, 0, 2, 1, 1, 0),       # head, Cout=2 (scalar epilogue)
    (1, 6, 6, 8, 0, 16, 3, 1, 1),         # Cin % 16 != 0
    (2, 4, 4, 512, 0, 512, 3, 1, 1),      # deep K, wide tile candidate
    (4, 32, 32, 128, 0, 128, 3, 1, 1),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_epilogue(case, dtype):
    N, H, W, C1, C2, Cout, k, s, p = case
    x = q(rnd(N, C1 + C2, H, W, seed=1), dtype)
    w = q(rnd(Cout, C1 + C2, k, k, seed=2, scale=(1.0 / (k * k * (C1 + C2)) ** 0.5)), dtype)
    bias, scale, shift = rnd(Cout, seed=3), rnd(Cout, seed=4).abs() + 0.5, rnd(Cout, seed=5)
    ref = F.conv2d(x, w, bias, s, p)
    res = q(rnd(*ref.shape, seed=6), dtype)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    wp = ops.pack_weight(w.contiguous(), True, dtype)
    x1 = nhwc(x[:, :C1], dtype)
    x2 = nhwc(x[:, C1:], dtype) if C2 else None
    y = ops.conv2d(x1, wp, Cout, k, s, p, x2=x2, bias=bias, scale=scale, shift=shift, residual=nhwc(res, dtype), relu=True)
    assert rel(nchw(y), ref) < tol(dtype)
    # plain (no epilogue), fp32 output from bf16 inputs
    y2 = ops.conv2d(x1, wp, Cout, k, s, p, x2=x2, y_dtype=torch.float32)
    assert y2.dtype == torch.float32
    assert rel(nchw(y2), F.conv2d(x, w, None, s, p)) < (2e-5 if dtype == torch.float32 else 2e-3)


CONVT_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, out_pad
    (2, 8, 8, 64, 32, 3, 2, 1, 1),
    (1, 5, 7, 128, 64, 3, 2, 1, 1),
    (2, 6, 6, 32, 16, 2, 2, 0, 0),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transposed_fwd(case, dtype):
    N, H, W, Cin, Cout, k, s, p, op = case
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = q(rnd(Cin, Cout, k, k, seed=2, scale=0.1), dtype)
    b = rnd(Cout, seed=3)
    ref = F.conv_transpose2d(x, w, b, stride=s, padding=p, output_padding=op)
    wp = ops.pack_weight(w.contiguous(), False, dtype)
    y = ops.conv2d(nhwc(x, dtype), wp, Cout, k, s, p, mode=ops.CONV_TRANSPOSED, out_hw=ref.shape[2:], bias=b)
    assert rel(nchw(y), ref) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad_wgrad(case, dtype):
    N, H, W, C1, C2, Cout, k, s, p = case
    Cin = C1 + C2
    x = q(rnd(N, Cin, H, W, seed=1), dtype).requires_grad_(True)
    w = q(rnd(Cout, Cin, k, k, seed=2, scale=0.1), dtype).requires_grad_(True)
    y = F.conv2d(x, w, None, s, p)
    dy = q(rnd(*y.shape, seed=7), dtype)
    y.backward(dy)
    # dgrad = transposed-gather conv over dy with the [ (tap,co) ][ ci ] packing
    wpd = ops.pack_weight(w.detach().contiguous(), False, dtype)
    dyn = nhwc(dy, dtype)
    dx1 = ops.conv2d(dyn, wpd, C1, k, s, p, mode=ops.CONV_TRANSPOSED, out_hw=(H, W), ldw=Cin, w_offset=0)
    assert rel(nchw(dx1), x.grad[:, :C1]) < tol(dtype)
    if C2:
        dx2 = ops.conv2d(dyn, wpd, C2, k, s, p, mode=ops.CONV_TRANSPOSED, out_hw=(H, W), ldw=Cin, w_offset=C1)
        assert rel(nchw(dx2), x.grad[:, C1:]) < tol(dtype)
    dW = torch.zeros_like(w)
    xd = x.detach()
    ops.conv2d_wgrad(dyn, nhwc(xd[:, :C1], dtype), dW, k, s, p, 0, Cin)
    if C2:
        ops.conv2d_wgrad(dyn, nhwc(xd[:, C1:], dtype), dW, k, s, p, C1, Cin)
    assert rel(dW, w.grad) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transposed_bwd(case, dtype):
    N, H, W, Cin, Cout, k, s, p, op = case
    x = q(rnd(N, Cin, H, W, seed=1), dtype).requires_grad_(True)
    w = q(rnd(Cin, Cout, k, k, seed=2, scale=0.1), dtype).requires_grad_(True)
    y = F.conv_transpose2d(x, w, None, stride=s, padding=p, output_padding=op)
    dy = q(rnd(*y.shape, seed=7), dtype)
    y.backward(dy)
    dyn = nhwc(dy, dtype)
    wpd = ops.pack_weight(w.detach().contiguous(), True, dtype)
    dx = ops.conv2d(dyn, wpd, Cin, k, s, p, mode=ops.CONV_FWD, out_hw=(H, W))
    assert rel(nchw(dx), x.grad) < tol(dtype)
    dW = torch.zeros_like(w)
    ops.conv2d_wgrad(nhwc(x.detach(), dtype), dyn, dW, k, s, p, 0, Cout)
    assert rel(dW, w.grad) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("G,B,H,W,C,relu,use_res", [(3, 2, 8, 8, 64, True, True), (1, 4, 5, 7, 32, True, False),
                                                    (2, 1, 6, 6, 128, False, False), (1, 2, 4, 4, 2, False, False),
                                                    (4, 2, 16, 16, 512, True, True), (8, 4, 40, 36, 64, True, True),
                                                    (8, 2, 24, 24, 256, True, False), (2, 3, 7, 5, 24, True, True)])
def test_batchnorm_train_fwd_bwd(G, B, H, W, C, relu, use_res, dtype):
    N = G * B
    x = q(rnd(N, C, H, W, seed=1) * 2 + 0.5, dtype).requires_grad_(True)
    res = q(rnd(N, C, H, W, seed=2), dtype).requires_grad_(True)
    gamma = (rnd(C, seed=3).abs() + 0.5).requires_grad_(True)
    beta = rnd(C, seed=4).requires_grad_(True)
    rm0, rv0 = rnd(C, seed=5), rnd(C, seed=6).abs() + 0.5
    rm, rv = rm0.clone(), rv0.clone()
    outs = []
    for g in range(G):   # reference semantics: one BN call per group, sequential running-stat updates
        o = F.batch_norm(x[g * B:(g + 1) * B], rm, rv, gamma, beta, True, 0.1, 1e-5)
        outs.append(o)
    ref = torch.cat(outs, 0)
    if use_res:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    dy = q(rnd(*ref.shape, seed=7), dtype)
    ref.backward(dy)

    R = B * H * W
    xn = nhwc(x.detach(), dtype)
    rm_k, rv_k = rm0.clone(), rv0.clone()
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    sums = ops.bn_stats(xn, G, R, C)
    st = ops.bn_finalize_train(sums, gamma.detach(), beta.detach(), rm_k, rv_k, nbt, G, R, C)
    resn = nhwc(res.detach(), dtype) if use_res else None
    y = ops.bn_apply(xn, st[0], st[1], G, R, C, relu, resn)
    assert rel(nchw(y), ref) < tol(dtype)
    assert int(nbt) == G
    assert rel(rm_k, rm) < 1e-5 and rel(rv_k, rv) < (1e-5 if dtype == torch.float32 else 1e-5)
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx, dres = ops.bn_bwd(nhwc(dy, dtype), y, xn, st[2], st[3], gamma.detach(), dgamma, dbeta, G, R, C, relu, use_res)
    t = tol(dtype) * (1 if dtype == torch.float32 else 3)
    assert rel(nchw(dx), x.grad) < t
    assert rel(dgamma, gamma.grad) < t and rel(dbeta, beta.grad) < t
    if relu and not use_res:
        # mask recomputed from x (fma(x, scale, shift) > 0) instead of read from y: bit-identical to the y route
        dg2, db2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dx2, _ = ops.bn_bwd(nhwc(dy, dtype), None, xn, st[2], st[3], gamma.detach(), dg2, db2, G, R, C, relu, False,
                            scale=st[0], shift=st[1])
        assert torch.equal(dx2, dx) and torch.equal(dg2, dgamma) and torch.equal(db2, dbeta)
    if use_res:
        assert rel(nchw(dres), res.grad) < t
        # accumulate form
        acc = nhwc(torch.ones_like(res), dtype)
        _, dres2 = ops.bn_bwd(nhwc(dy, dtype), y, xn, st[2], st[3], gamma.detach(), None, None, G, R, C, relu, True, dres_acc=acc)
        assert rel(nchw(dres2), res.grad + 1) < t

    # ---- the one-launch backward (reduce -> group barrier -> finalize -> apply) against the same reference ----
    def scratch():
        return torch.zeros(ops.bn_bwd_scratch_floats(G, C), device=DEV)
    dg3, db3 = torch.ones(C, device=DEV), torch.full((C,), 2.0, device=DEV)          # accumulated into, not overwritten
    dx3, dres3 = ops.bn_bwd(nhwc(dy, dtype), y, xn, st[2], st[3], gamma.detach(), dg3, db3, G, R, C, relu, use_res, scratch=scratch())
    assert rel(nchw(dx3), x.grad) < t and rel(dx3, dx) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert rel(dg3 - 1, gamma.grad) < t and rel(db3 - 2, beta.grad) < t
    if relu and not use_res:
        dx4, _ = ops.bn_bwd(nhwc(dy, dtype), None, xn, st[2], st[3], gamma.detach(), None, None, G, R, C, relu, False,
                            scale=st[0], shift=st[1], scratch=scratch())
        assert rel(dx4, dx3) < (1e-6 if dtype == torch.float32 else 1e-2)
    if use_res:
        assert rel(nchw(dres3), res.grad) < t
        acc = nhwc(torch.ones_like(res), dtype)
        _, dres4 = ops.bn_bwd(nhwc(dy, dtype), y, xn, st[2], st[3], gamma.detach(), None, None, G, R, C, relu, True, dres_acc=acc,
                              scratch=scratch())
        assert rel(nchw(dres4), res.grad + 1) < t


def test_bn_fold_eval_matches_batch_norm():
    C = 48
    gamma, beta, rm, rv = rnd(C, seed=1), rnd(C, seed=2), rnd(C, seed=3), rnd(C, seed=4).abs() + 0.1
    x = rnd(2, C, 5, 5, seed=5)
    f = ops.bn_fold_eval(gamma, beta, rm, rv)
    ref = F.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert rel(x * f[0].view(1, -1, 1, 1) + f[1].view(1, -1, 1, 1), ref) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,s,p,H,W,C", [(3, 2, 1, 16, 16, 64), (3, 2, 1, 9, 13, 64), (2, 2, 0, 8, 8, 32), (2, 2, 0, 7, 9, 6),
                                             (3, 2, 1, 42, 38, 64), (3, 2, 1, 20, 12, 128), (2, 2, 0, 12, 20, 2056)])
def test_maxpool(k, s, p, H, W, C, dtype):
    x = q(rnd(2, C, H, W, seed=1), dtype).requires_grad_(True)
    ref = F.max_pool2d(x, k, s, p)
    dy = q(rnd(*ref.shape, seed=2), dtype)
    ref.backward(dy)
    xn = nhwc(x.detach(), dtype)
    y = ops.maxpool_fwd(xn, k, s, p)
    assert torch.equal(nchw(y), ref.detach())
    dx = ops.maxpool_bwd(xn, nhwc(dy, dtype), k, s, p)
    assert rel(nchw(dx), x.grad) < (1e-6 if dtype == torch.float32 else 1e-2)
    # indexed (training) variants
    y2, idx = ops.maxpool_fwd_idx(xn, k, s, p)
    assert torch.equal(y2, y) and idx.dtype == torch.uint8 and int(idx.max()) < k * k
    dx2 = ops.maxpool_bwd_idx(idx, nhwc(dy, dtype), tuple(xn.shape), k, s, p)
    if dtype == torch.float32:
        assert torch.equal(dx2, dx)
    else:   # the bf16 fast path adds the (<= 4) window contributions of a pixel in bf16x2 instead of fp32
        assert rel(nchw(dx2), x.grad) < 1e-2 and rel(dx2, dx) < 1e-2
        # same routing: the zero patterns agree (up to the rare a + b == -c cancellation that only one rounding order hits)
        assert ((dx2 == 0) != (dx == 0)).float().mean().item() < 1e-4


@pytest.mark.parametrize("C", [4, 16])
def test_maxpool_ties_route_to_first_max(C):
    x = torch.zeros(1, C, 6, 6, device=DEV).requires_grad_(True)   # all equal: every window ties
    ref = F.max_pool2d(x, 3, 2, 1)
    dy = rnd(*ref.shape, seed=3)
    ref.backward(dy)
    xn = nhwc(x.detach(), torch.float32)
    dx = ops.maxpool_bwd(xn, nhwc(dy, torch.float32), 3, 2, 1)
    assert rel(nchw(dx), x.grad) < 1e-6
    _, idx = ops.maxpool_fwd_idx(xn, 3, 2, 1)
    dx2 = ops.maxpool_bwd_idx(idx, nhwc(dy, torch.float32), tuple(xn.shape), 3, 2, 1)
    assert rel(nchw(dx2), x.grad) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bilinear_align_corners(dtype):
    x = q(rnd(2, 16, 6, 6, seed=1), dtype).requires_grad_(True)
    ref = F.interpolate(x, size=(5, 9), mode="bilinear", align_corners=True)
    dy = q(rnd(*ref.shape, seed=2), dtype)
    ref.backward(dy)
    y = ops.bilinear_fwd(nhwc(x.detach(), dtype), 5, 9)
    assert rel(nchw(y), ref) < (1e-5 if dtype == torch.float32 else 1e-2)
    dx = ops.bilinear_bwd(nhwc(dy, dtype), 6, 6)
    assert rel(nchw(dx), x.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lstm_cell(dtype):
    R, C = 300, 64
    gates = rnd(R, 4 * C, seed=1).requires_grad_(True)
    c_prev = rnd(R, C, seed=2).requires_grad_(True)
    i, f, g, o = gates.split(C, 1)
    c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    dh, dc = rnd(R, C, seed=3), rnd(R, C, seed=4)
    (h * dh).sum().backward(retain_graph=True)
    g_h, cp_h = gates.grad.clone(), c_prev.grad.clone()
    gates.grad = None
    c_prev.grad = None
    (h * dh + c * dc).sum().backward()
    acts = torch.empty(R, 4 * C, dtype=dtype, device=DEV)
    c_out = torch.empty(R, C, device=DEV)
    h_out = torch.empty(R, C, dtype=dtype, device=DEV)
    ops.lstm_cell_fwd(gates.detach(), c_prev.detach(), acts, c_out, h_out, R, C)
    assert rel(c_out, c) < 1e-5 and rel(h_out, h) < (1e-5 if dtype == torch.float32 else 5e-3)
    dgates = torch.empty(R, 4 * C, dtype=dtype, device=DEV)
    dc_io = dc.clone()
    ops.lstm_cell_bwd(dh, dc_io, acts, c_prev.detach(), c_out, dgates, R, C)
    t = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel(dgates, gates.grad) < t and rel(dc_io, c_prev.grad) < t
    # t = 0 form (no previous state)
    ops.lstm_cell_fwd(gates.detach(), None, None, c_out, h_out, R, C)
    c0 = torch.sigmoid(i) * torch.tanh(g)
    assert rel(c_out, c0) < 1e-5
    assert g_h is not None and cp_h is not None


def test_layout_adapters():
    x = rnd(3, 5, 2, 6, 7, seed=1)    # B,T,C,H,W
    y = ops.pack_series(x, torch.float32)
    ref = x.permute(1, 0, 3, 4, 2).reshape(15, 6, 7, 2)
    assert torch.equal(y, ref)
    yb = ops.pack_series(x, torch.bfloat16)
    assert torch.equal(yb, ref.to(torch.bfloat16))
    z = rnd(2, 9, 4, 3, seed=2)       # NHWC
    assert torch.equal(ops.nhwc_to_nchw(z), z.permute(0, 3, 1, 2).contiguous())
    g = rnd(2, 3, 9, 4, seed=3)       # NCHW
    assert torch.equal(ops.nchw_to_nhwc(g, torch.float32), g.permute(0, 2, 3, 1).contiguous())
    a, b = rnd(1000, seed=4), rnd(1000, seed=5)
    assert torch.equal(ops.add_(a.clone(), b), a + b)
    assert torch.equal(ops.cast(a, torch.bfloat16), a.to(torch.bfloat16))
    cs = torch.zeros(12, device=DEV)
    m = rnd(1000, 12, seed=6)
    ops.colsum(m, cs, 1000, 12)
    assert rel(cs, m.sum(0)) < 1e-5


@pytest.mark.parametrize("B,C,H,W", [(3, 2, 24, 40), (2, 4, 17, 5), (16, 2, 128, 128)])
def test_ce_dice_loss(B, C, H, W):
    from oracle import stf_oracle as O
    logits = (rnd(B, C, H, W, seed=1) * 3).requires_grad_(True)
    target = (torch.rand(B, H, W, generator=torch.Generator().manual_seed(2)) * C).long().clamp_(0, C - 1).to(DEV)
    ref = O.criterion(logits, target)
    ref.backward()
    out, stats = ops.ce_dice_fwd(logits.detach(), target)
    assert abs(out[0].item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item()))
    up = torch.tensor([0.5], device=DEV)
    dl = ops.ce_dice_bwd(logits.detach(), target, stats, up)
    assert rel(dl, 0.5 * logits.grad) < 2e-5


def test_ce_dice_golden(golden_dir):
    import numpy as np
    import os
    g = np.load(os.path.join(golden_dir, "criterion_3x2x24x40.npz"))
    logits = torch.from_numpy(g["logits"]).to(DEV)
    target = torch.from_numpy(g["target"]).to(DEV)
    out, stats = ops.ce_dice_fwd(logits, target)
    assert abs(out[0].item() - float(g["loss"])) < 2e-6
    dl = ops.ce_dice_bwd(logits, target, stats, None)
    assert rel(dl.cpu(), torch.from_numpy(g["grad"])) < 2e-5


@pytest.mark.parametrize("tag,weighted,dice", [("w_ign_dice", True, True), ("ign_dice", False, True), ("w_ign_nodice", True, False)])
def test_ce_dice_options_golden(golden_dir, tag, weighted, dice):
    """criterion(loss_weight, ignore_index=255, dice) through stf_unet_b200.criterion == the reference's own criterion
    (train_utils/train_and_eval.py:299-313) on a fixture with ignored pixels and one fully ignored image."""
    import numpy as np
    import os
    import stf_unet_b200 as S
    g = np.load(os.path.join(golden_dir, "criterion_options_4x3x20x28.npz"))
    logits = torch.from_numpy(g["logits"]).to(DEV).requires_grad_(True)
    target = torch.from_numpy(g["target"]).to(DEV)
    w = torch.from_numpy(g["weight"]).to(DEV) if weighted else None
    loss = S.criterion({"out": logits}, target, loss_weight=w, num_classes=3, dice=dice, ignore_index=255)
    loss.backward()
    assert abs(loss.item() - float(g["loss_" + tag])) < 2e-6 * max(1.0, abs(float(g["loss_" + tag])))
    assert rel(logits.grad.cpu(), torch.from_numpy(g["grad_" + tag])) < 2e-5
    # ignored pixels carry exactly zero gradient
    assert logits.grad[3].abs().max().item() == 0.0


def test_ce_dice_flags_labels_out_of_range(monkeypatch):
    """A label outside [0, C) that is not ignore_index: the reference raises; here the loss is NaN (no host sync), or an
    IndexError with STFB_CHECK_TARGETS=1."""
    import stf_unet_b200 as S
    logits = rnd(2, 2, 8, 8, seed=1)
    target = torch.zeros(2, 8, 8, dtype=torch.int64, device=DEV)
    target[0, 0, 0] = 255
    assert torch.isnan(S.criterion({"out": logits}, target))
    assert torch.isfinite(S.criterion({"out": logits}, target, ignore_index=255))
    monkeypatch.setenv("STFB_CHECK_TARGETS", "1")
    with pytest.raises(IndexError):
        S.criterion({"out": logits}, target)


def test_errors_are_raised_not_swallowed():
    x = torch.zeros(1, 4, 4, 8, device=DEV)
    wp = torch.zeros(8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="output size"):
        ops.conv2d(x, wp, 8, 1, 1, 0, out_hw=(9, 9))
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.conv2d(x.cpu(), wp, 8, 1, 1, 0)
    with pytest.raises(RuntimeError, match="tcgen05"):
        ops.conv2d(x, wp, 8, 1, 1, 0, impl=ops.IMPL_TCGEN05)   # fp32 is never a tensor-core shape


# ------------------------------------------------------------------------------------------------
# tcgen05 / TMEM family
# ------------------------------------------------------------------------------------------------
TC_CASES = [
    # N, H, W, C1, C2, Cout, k
    (2, 16, 16, 64, 0, 64, 3),       # one 16x8 patch per tile, BN=64
    (3, 8, 8, 128, 0, 256, 3),       # two images per tile, ragged N (3 = 2 + 1), BN=256
    (1, 25, 25, 64, 0, 128, 3),      # ragged H/W (100/4), BN=128
    (2, 4, 4, 512, 0, 512, 3),       # TN=8 patch > N, deep K (72 k-blocks), two N tiles
    (2, 32, 32, 64, 64, 64, 1),      # dual-source 1x1 fusion
    (2, 16, 16, 128, 128, 128, 3),   # dual-source 3x3 (UNet decoder)
    (4, 16, 16, 64, 0, 256, 1),      # LSTM gate GEMM shape (1x1, N = 4C)
    (1, 64, 64, 64, 0, 32, 3),       # BN=32
    (16, 32, 32, 128, 0, 128, 3),    # many tiles: every persistent CTA walks several
    (2, 32, 32, 32, 0, 32, 3),       # 32-channel layer (final_res): 64-byte swizzle path
    (2, 16, 16, 32, 0, 64, 1),       # 32 -> 64, 1x1
    (2, 16, 16, 64, 32, 64, 3),      # mixed 64 + 32 channel concat
    (3, 20, 24, 96, 0, 160, 3),      # channel counts that are multiples of 32 but not 64; Cout = 5 x 32
    (2, 16, 16, 256, 0, 256, 3),     # halo kernel, BN=256, 4 channel blocks
    (36, 32, 32, 128, 0, 128, 3),    # halo, two pixel tiles per weight pass (MT=2), BN=128, double-buffered TMEM
    (112, 16, 16, 256, 0, 256, 3),   # halo MT=2 at BN=256: 512 TMEM columns, single buffered
    (225, 16, 8, 64, 0, 64, 3),      # halo MT=2 with an odd number of pixel tiles (dead second accumulator in the tail)
    (230, 16, 8, 128, 64, 64, 3),    # halo MT=2, dual source, streamed weights at BN=64
    (40, 32, 32, 64, 0, 128, 1),     # streaming kernel, 320 tiles: every persistent CTA walks 2-3 (staging vs barriers)
    (24, 8, 8, 256, 0, 512, 3),      # streaming 3x3 on 8x8 maps, 2 N tiles x 12 pixel tiles... and
    (160, 8, 8, 128, 0, 128, 3),     # ...80 pixel tiles x 1: with the 1x1 above, several tiles per CTA on every path
    (300, 8, 8, 64, 0, 64, 3),       # 150 tiles of two images at BN=64
    (40, 32, 32, 64, 0, 64, 3),      # halo kernel with resident weights: 320 tiles, every CTA walks 2-3
    (5, 13, 9, 128, 0, 64, 3),       # halo kernel, ragged patch (13 x 9 inside 16 x 16), streamed weights at BN=64
]


@pytest.mark.parametrize("halo_mode", ["single", "mt2", "pair"])
@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_matches_reference(case, halo_mode, monkeypatch):
    """single: one CTA per tile; mt2: two pixel tiles per weight pass (STFB_HALO_MT=2); pair: CTA pairs driving M = 256
    MMAs (tcgen05 cta_group::2, the default whenever a 3x3 layer's pixel tiles pair up)."""
    N, H, W, C1, C2, Cout, k = case
    if halo_mode == "mt2" and not (k == 3 and N * H * W >= 148 * 96):
        pytest.skip("tile pairing (STFB_HALO_MT=2) only engages on 3x3 layers with >= 111 super tiles")
    if halo_mode == "pair" and not (k == 3 and H >= 12 and W >= 8):
        pytest.skip("CTA pairs only engage on the halo (3x3 / stride 1, maps >= 12 x 8) path")
    monkeypatch.setenv("STFB_HALO_MT", "2" if halo_mode == "mt2" else "0")
    monkeypatch.setenv("STFB_HALO_PAIR", "1" if halo_mode == "pair" else "0")
    dtype = torch.bfloat16
    pad = (k - 1) // 2
    x = q(rnd(N, C1 + C2, H, W, seed=1), dtype)
    w = q(rnd(Cout, C1 + C2, k, k, seed=2, scale=(1.0 / (k * k * (C1 + C2)) ** 0.5)), dtype)
    x1 = nhwc(x[:, :C1], dtype)
    x2 = nhwc(x[:, C1:], dtype) if C2 else None
    assert ops.tcgen05_ok(x1, Cout, k, 1, pad, x2=x2)
    wp = ops.pack_weight(w.contiguous(), True, dtype, n_major=True)
    assert wp.shape == (Cout, k * k * (C1 + C2))
    ref = F.conv2d(x, w, None, 1, pad)
    y = ops.conv2d(x1, wp, Cout, k, 1, pad, x2=x2, y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert rel(nchw(y), ref) < 2e-3
    # fused epilogue, bf16 out
    bias, scale, shift = rnd(Cout, seed=3), rnd(Cout, seed=4).abs() + 0.5, rnd(Cout, seed=5)
    res = q(rnd(*ref.shape, seed=6), dtype)
    ref2 = F.relu((ref + bias.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    y2 = ops.conv2d(x1, wp, Cout, k, 1, pad, x2=x2, bias=bias, scale=scale, shift=shift, residual=nhwc(res, dtype),
                    relu=True, impl=ops.IMPL_TCGEN05)
    assert y2.dtype == dtype
    assert rel(nchw(y2), ref2) < 1e-2
    # agrees with the SIMT family on identical inputs
    wps = ops.pack_weight(w.contiguous(), True, dtype)
    y3 = ops.conv2d(x1, wps, Cout, k, 1, pad, x2=x2, y_dtype=torch.float32, impl=ops.IMPL_SIMT)
    assert rel(y, y3) < 1e-5


def test_conv_tcgen05_dgrad_via_flipped_weights():
    N, H, W, Cin, Cout, k = 2, 16, 16, 128, 64, 3
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype).requires_grad_(True)
    w = q(rnd(Cout, Cin, k, k, seed=2, scale=0.05), dtype)
    y = F.conv2d(x, w, None, 1, 1)
    dy = q(rnd(*y.shape, seed=3), dtype)
    y.backward(dy)
    # dgrad as a forward conv over dy: B[n = ci][(ky', kx', co)] = W[co][ci][k-1-ky'][k-1-kx']
    wpd = ops.pack_weight(w.contiguous(), False, dtype, n_major=True, flip=True)
    assert wpd.shape == (Cin, k * k * Cout)
    dx = ops.conv2d(nhwc(dy, dtype), wpd, Cin, k, 1, 1, y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
    assert rel(nchw(dx), x.grad) < 2e-3


WG_TC_CASES = [
    # N, H, W, Cin, Cout, k
    (2, 16, 16, 64, 64, 3),        # 9 column blocks -> last M tile half empty; taps share tiles
    (3, 8, 8, 128, 256, 3),        # BN=256
    (1, 25, 25, 64, 128, 3),       # ragged
    (2, 4, 4, 512, 512, 3),        # TN=4 patches, two N tiles
    (4, 16, 16, 256, 64, 1),       # 1x1 (LSTM dW shape: P = dgates, G = x)
    (8, 32, 32, 128, 128, 3),      # split-K over many patches
]


@pytest.mark.parametrize("case", WG_TC_CASES)
def test_wgrad_tcgen05_matches_reference(case):
    N, H, W, Cin, Cout, k = case
    dtype = torch.bfloat16
    pad = (k - 1) // 2
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
    y = F.conv2d(x, w, None, 1, pad)
    dy = q(rnd(*y.shape, seed=2), dtype)
    y.backward(dy)
    dW = torch.zeros(Cout, Cin, k, k, device=DEV)
    ops.conv2d_wgrad(nhwc(dy, dtype), nhwc(x, dtype), dW, k, 1, pad, 0, Cin, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert rel(dW, w.grad) < 1e-4
    # channel window form (dual-source convs call wgrad once per source)
    if Cin >= 128:
        h = Cin // 2
        dW2 = torch.zeros_like(dW)
        ops.conv2d_wgrad(nhwc(dy, dtype), nhwc(x[:, :h], dtype), dW2, k, 1, pad, 0, Cin, impl=ops.IMPL_TCGEN05)
        ops.conv2d_wgrad(nhwc(dy, dtype), nhwc(x[:, h:], dtype), dW2, k, 1, pad, h, Cin, impl=ops.IMPL_TCGEN05)
        assert rel(dW2, w.grad) < 1e-4


TC_S2_FWD = [(2, 16, 16, 64, 128, 3, 1), (3, 9, 11, 64, 64, 3, 1), (2, 32, 32, 128, 256, 1, 0), (16, 64, 64, 64, 128, 3, 1)]


@pytest.mark.parametrize("case", TC_S2_FWD)
def test_conv_tcgen05_stride2_forward(case):
    N, H, W, Cin, Cout, k, pad = case
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = q(rnd(Cout, Cin, k, k, seed=2, scale=(1.0 / (k * k * Cin) ** 0.5)), dtype)
    ref = F.conv2d(x, w, None, 2, pad)
    xn = nhwc(x, dtype)
    assert ops.tcgen05_ok(xn, Cout, k, 2, pad)
    wp = ops.pack_weight(w.contiguous(), True, dtype, n_major=True)
    y = ops.conv2d(xn, wp, Cout, k, 2, pad, y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert y.shape[1:3] == ref.shape[2:]
    assert rel(nchw(y), ref) < 2e-3


TC_TRANSPOSED = [
    # N, Hin, Win, Cin(gemm K), Cout(gemm N), k, stride, pad, out_pad
    (2, 8, 8, 128, 64, 3, 2, 1, 1),      # ConvTranspose2d 3x3/2 (decoder up) == 3x3/2 conv dgrad to an even size
    (1, 13, 13, 64, 64, 3, 2, 1, 0),     # dgrad of a 3x3/2 conv over a 25x25 input (odd size: ragged phases)
    (2, 8, 8, 256, 128, 2, 2, 0, 0),     # UNet ConvTranspose2d 2x2/2
    (2, 8, 8, 128, 64, 1, 2, 0, 1),      # dgrad of the 1x1/2 downsample conv: three of four phases are pure zeros
    (2, 16, 16, 64, 128, 3, 1, 1, 0),    # stride-1 dgrad without flipping the weights
    (16, 16, 16, 256, 128, 3, 2, 1, 1),
    (2, 16, 16, 64, 32, 3, 2, 1, 1),     # upconv1: 64 -> 32 channels
    (2, 16, 16, 32, 64, 3, 2, 1, 0),     # its dgrad direction: K = 32 channels
]


@pytest.mark.parametrize("case", TC_TRANSPOSED)
def test_conv_tcgen05_transposed_phases(case):
    N, H, W, Cin, Cout, k, s, pad, op = case
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = q(rnd(Cin, Cout, k, k, seed=2, scale=0.05), dtype)       # ConvTranspose2d layout [in, out, k, k]
    bias = rnd(Cout, seed=3)
    ref = F.conv_transpose2d(x, w, bias, stride=s, padding=pad, output_padding=op)
    res = q(rnd(*ref.shape, seed=4), dtype)
    xn = nhwc(x, dtype)
    out_hw = tuple(ref.shape[2:])
    assert ops.tcgen05_ok(xn, Cout, k, s, pad, mode=ops.CONV_TRANSPOSED, out_hw=out_hw)
    wp = ops.pack_weight(w.contiguous(), False, dtype, n_major=True)    # [co][(tap, ci)]
    y = ops.conv2d(xn, wp, Cout, k, s, pad, mode=ops.CONV_TRANSPOSED, out_hw=out_hw, bias=bias, residual=nhwc(res, dtype).float(),
                   y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert rel(nchw(y), ref + res) < 2e-3


WG_S2 = [(2, 16, 16, 64, 128, 3, 1), (3, 9, 11, 64, 64, 3, 1), (2, 32, 32, 128, 256, 1, 0), (8, 64, 64, 64, 128, 3, 1)]


@pytest.mark.parametrize("case", WG_S2)
def test_wgrad_tcgen05_stride2(case):
    N, H, W, Cin, Cout, k, pad = case
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
    y = F.conv2d(x, w, None, 2, pad)
    dy = q(rnd(*y.shape, seed=2), dtype)
    y.backward(dy)
    dW = torch.zeros_like(w)
    ops.conv2d_wgrad(nhwc(dy, dtype), nhwc(x, dtype), dW, k, 2, pad, 0, Cin, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert rel(dW, w.grad) < 1e-4


@pytest.mark.parametrize("case", [(2, 16, 16, 32, 32), (1, 12, 20, 32, 32), (3, 7, 10, 32, 96)])
def test_wgrad_pixel_pairs_32_channels(case):
    """32-channel 3x3 layers (final_res, src/stf_lstm_unet.py:136) through the 64-channel tcgen05 wgrad kernel via pixel
    pairing ([N,H,W,32] viewed as [N,H,W/2,64]) + the fold kernel."""
    N, H, W, Cin, Cout = case
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = torch.zeros(Cout, Cin, 3, 3, device=DEV, requires_grad=True)
    y = F.conv2d(x, w, None, 1, 1)
    dy = q(rnd(*y.shape, seed=2), dtype)
    y.backward(dy)
    dW = torch.full((Cout, Cin, 3, 3), 0.5, device=DEV)       # accumulates into what is there
    from stf_unet_b200 import _lib
    n0 = _lib.launch_count()
    ops.conv2d_wgrad(nhwc(dy, dtype), nhwc(x, dtype), dW, 3, 1, 1, 0, Cin)
    torch.cuda.synchronize()
    assert rel(dW - 0.5, w.grad) < 1e-4


def test_wgrad_tcgen05_conv_transpose():
    N, H, W, Cin, Cout, k = 2, 8, 8, 128, 64, 3
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = torch.zeros(Cin, Cout, k, k, device=DEV, requires_grad=True)
    y = F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=1)
    dy = q(rnd(*y.shape, seed=2), dtype)
    y.backward(dy)
    dW = torch.zeros_like(w)
    ops.conv2d_wgrad(nhwc(x, dtype), nhwc(dy, dtype), dW, k, 2, 1, 0, Cout, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert rel(dW, w.grad) < 1e-4


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad", [(3, 20, 20, 1, 64, 7, 2, 3), (2, 16, 16, 8, 64, 3, 1, 1), (2, 33, 31, 1, 32, 3, 1, 1),
                                                            (2, 70, 138, 1, 64, 7, 2, 3), (2, 26, 74, 4, 64, 7, 2, 3), (1, 9, 70, 8, 64, 3, 1, 1),
                                                            (1, 12, 12, 2, 64, 5, 1, 2)])
def test_small_cin_conv_via_im2col(N, H, W, Cin, Cout, k, stride, pad):
    """7x7 stem / UNet first conv: im2col (K padded to 64) + 1x1 tcgen05 GEMM, and the wgrad through the same buffer."""
    dtype = torch.bfloat16
    x = q(rnd(N, Cin, H, W, seed=1), dtype)
    w = q(rnd(Cout, Cin, k, k, seed=2, scale=0.2), dtype).requires_grad_(True)
    y = F.conv2d(x, w, None, stride, pad)
    dy = q(rnd(*y.shape, seed=3), dtype)
    y.backward(dy)
    kpad = (Cin * k * k + 63) // 64 * 64
    col = ops.im2col_small(nhwc(x, dtype), k, stride, pad, kpad)
    wp = ops.pack_weight(w.detach().contiguous(), True, dtype, n_major=True, kpad=kpad)
    out = ops.conv2d(col, wp, Cout, 1, 1, 0, y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
    assert rel(nchw(out), y.detach()) < 2e-3
    scratch = torch.zeros(Cout, kpad, device=DEV)
    ops.conv2d_wgrad(nhwc(dy, dtype), col, scratch, 1, 1, 0, 0, kpad)
    dW = torch.zeros(Cout, Cin, k, k, device=DEV)
    ops.unpad_wgrad(dW, scratch)
    assert rel(dW, w.grad) < 1e-4


def _acts_il_to_gate_major(acts_il, R, C):
    """[R, 4C] in accumulator column order ((u/16)*64 + gate*16 + u%16) -> gate-major [R, gate*C + u]."""
    return acts_il.view(R, C // 16, 4, 16).permute(0, 2, 1, 3).reshape(R, 4 * C)


@pytest.mark.parametrize("B,h,w,C", [(2, 16, 16, 64), (3, 8, 8, 128), (1, 10, 12, 256), (2, 4, 4, 512), (12, 64, 64, 64)])
def test_lstm_step_fused_matches_reference(B, h, w, C):
    """[x_t, h_{t-1}] @ [W_ih | W_hh]^T + biases + cell update in one tcgen05 kernel vs the gate-by-gate reference
    (reference: nn.LSTM at src/stf_lstm_unet.py:124-127, :216-242)."""
    bf = torch.bfloat16
    R = B * h * w
    xt = q(rnd(R, C, seed=7) * 0.7, bf)
    hp = q(rnd(R, C, seed=1) * 0.5, bf)
    wih = q(rnd(4 * C, C, seed=5, scale=1.0 / C ** 0.5), bf)
    whh = q(rnd(4 * C, C, seed=2, scale=1.0 / C ** 0.5), bf)
    bih = rnd(4 * C, seed=3) * 0.1
    bhh = rnd(4 * C, seed=6) * 0.1
    cp = rnd(R, C, seed=4)
    wp = ops.pack_lstm_xh(wih.contiguous(), whh.contiguous(), bf)

    def ref(with_h):
        gates = xt @ wih.t() + bih + bhh
        cprev = torch.zeros_like(cp)
        if with_h:
            gates = gates + hp @ whh.t()
            cprev = cp
        i, f, g, o = gates.split(C, 1)
        c_ref = torch.sigmoid(f) * cprev + torch.sigmoid(i) * torch.tanh(g)
        h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
        a_ref = torch.cat([torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)], 1)
        return c_ref, h_ref, a_ref

    for with_h in (True, False):
        c_ref, h_ref, a_ref = ref(with_h)
        c_out = torch.full((R, C), float("nan"), device=DEV)
        h_out = torch.empty(B, h, w, C, device=DEV, dtype=bf)
        acts = torch.empty(B, h, w, 4 * C, device=DEV, dtype=bf)
        ops.lstm_step_fused(xt.to(bf).view(B, h, w, C), hp.to(bf).view(B, h, w, C) if with_h else None, wp, bih, bhh,
                            cp if with_h else None, c_out, h_out, acts)
        torch.cuda.synchronize()
        assert rel(c_out, c_ref) < 3e-3
        assert rel(h_out.view(R, C), h_ref) < 6e-3
        assert rel(_acts_il_to_gate_major(acts.view(R, 4 * C).float(), R, C), a_ref) < 6e-3
        # eval form: no saved activations
        c2 = torch.empty_like(c_out)
        ops.lstm_step_fused(xt.to(bf).view(B, h, w, C), hp.to(bf).view(B, h, w, C) if with_h else None, wp, bih, bhh,
                            cp if with_h else None, c2, h_out, None)
        assert rel(c2, c_ref) < 3e-3


@pytest.mark.parametrize("B,T,h,w", [(4, 5, 16, 16), (2, 8, 64, 64), (6, 3, 8, 8), (3, 4, 20, 20), (1, 1, 16, 16)])
def test_lstm_seq_fused_equals_step_by_step_and_oracle(B, T, h, w):
    """The all-T kernel (hidden size 64: weights, c and h resident in shared memory for the whole time loop) against T launches
    of the per-step kernel -- same arithmetic, same accumulation order: bit-identical c, h and saved activations -- and against
    the oracle's gate-by-gate LSTM (reference: nn.LSTM over [B*h*w, T, C], src/stf_lstm_unet.py:124-127, :216-221)."""
    from oracle import stf_oracle as O
    C = 64
    bf = torch.bfloat16
    R = B * h * w
    assert ops.lstm_seq_supported(T, B, h, w, C)
    x = q(rnd(T * B, h, w, C, seed=3) * 0.7, bf).to(bf).contiguous()
    wih = q(rnd(4 * C, C, seed=5, scale=1.0 / C ** 0.5), bf)
    whh = q(rnd(4 * C, C, seed=2, scale=1.0 / C ** 0.5), bf)
    bih, bhh = rnd(4 * C, seed=3) * 0.1, rnd(4 * C, seed=6) * 0.1
    wp = ops.pack_lstm_xh(wih.contiguous(), whh.contiguous(), bf)
    # step by step
    cs = torch.empty(T, R, C, device=DEV)
    hs = torch.empty(T, B, h, w, C, device=DEV, dtype=bf)
    acts = torch.empty(T, R, 4 * C, device=DEV, dtype=bf)
    xs = x.view(T, B, h, w, C)
    for t in range(T):
        ops.lstm_step_fused(xs[t], hs[t - 1] if t else None, wp, bih, bhh, cs[t - 1] if t else None, cs[t], hs[t],
                            acts[t].view(B, h, w, 4 * C))
    # one launch, training form
    cs2, hs2, acts2 = torch.full_like(cs, float("nan")), torch.zeros_like(hs), torch.zeros_like(acts)
    ops.lstm_seq_fused(x, wp, bih, bhh, T, cs2, hs2, acts2)
    torch.cuda.synchronize()
    assert torch.equal(cs2, cs) and torch.equal(hs2, hs) and torch.equal(acts2, acts)
    # inference form: h_T only
    hT = torch.zeros(B, h, w, C, device=DEV, dtype=bf)
    ops.lstm_seq_fused(x, wp, bih, bhh, T, None, hT, None)
    assert torch.equal(hT, hs[T - 1])
    # oracle: seq [B, T, C, h, w] -> h_T [B, C, h, w]
    seq = x.float().view(T, B, h, w, C).permute(1, 0, 4, 2, 3).contiguous()
    h_ref = O.pixel_lstm_last(seq, wih, whh, bih, bhh)
    assert rel(hT.float().permute(0, 3, 1, 2), h_ref) < 1e-2


@pytest.mark.parametrize("B,h,w,C", [(4, 16, 16, 64), (2, 32, 32, 128), (3, 8, 8, 256), (2, 8, 8, 512), (1, 20, 20, 64)])
def test_lstm_bwd_step_fused_equals_two_kernel_route(B, h, w, C):
    """One backward time step as ONE kernel (dG_t W_hh on the tensor cores, cell backward of step t-1 in the epilogue) against
    the two-launch route (recurrent GEMM to an fp32 dh, then lstm_cell_bwd): same arithmetic -> identical dG_{t-1} and dc;
    with and without a previous cell state (autograd of nn.LSTM, src/stf_lstm_unet.py:216-242)."""
    bf = torch.bfloat16
    R = B * h * w
    dg_next = q(rnd(B, h, w, 4 * C, seed=11) * 0.3, bf).to(bf).contiguous()
    whh = q(rnd(4 * C, C, seed=2, scale=1.0 / C ** 0.5), bf)
    whh_d = ops.pack_weight(whh.contiguous(), False, bf, n_major=True)          # [C][4C]: dgrad operand of W_hh
    acts = torch.sigmoid(rnd(R, 4 * C, seed=5)).to(bf).contiguous()            # post-activation gates, interleaved layout
    cc, cp = rnd(R, C, seed=6), rnd(R, C, seed=7)
    for with_prev in (True, False):
        dc_ref = rnd(R, C, seed=8) * 0.2
        dc_new = dc_ref.clone()
        dh = ops.conv2d(dg_next, whh_d, C, 1, 1, 0, y_dtype=torch.float32, impl=ops.IMPL_TCGEN05)
        dg_ref = torch.empty(R, 4 * C, device=DEV, dtype=bf)
        ops.lstm_cell_bwd(dh.view(R, C), dc_ref, acts, cp if with_prev else None, cc, dg_ref, R, C, acts_il=True)
        dg_new = torch.zeros(B, h, w, 4 * C, device=DEV, dtype=bf)
        ops.lstm_bwd_step_fused(dg_next, whh_d, acts.view(B, h, w, 4 * C), cp if with_prev else None, cc, dc_new, dg_new)
        torch.cuda.synchronize()
        # same operations in the same order; the compiler may contract a multiply-add differently in the two kernels, which
        # can flip the last bf16 bit of an element here and there
        assert rel(dg_new.view(R, 4 * C), dg_ref) < 1e-4
        assert rel(dc_new, dc_ref) < 1e-6
        assert (dg_new.view(R, 4 * C) != dg_ref).float().mean().item() < 1e-3


def test_lstm_seq_fused_rejects_what_it_cannot_tile():
    assert not ops.lstm_seq_supported(8, 16, 32, 32, 128)      # hidden size 128: weights do not fit shared memory
    assert not ops.lstm_seq_supported(4, 3, 8, 8, 64)          # two 8x8 images per tile, odd batch: a tile would straddle steps


def test_lstm_cell_bwd_interleaved_acts_matches_gate_major():
    """lstm_cell_bwd reads the fused step's activation layout (acts_il) and must agree with the gate-major form."""
    bf = torch.bfloat16
    R, C = 96, 128
    a_gm = torch.rand(R, 4 * C, device=DEV).to(bf)
    a_il = a_gm.view(R, 4, C // 16, 16).permute(0, 2, 1, 3).reshape(R, 4 * C).contiguous()
    dh, cp, cc = rnd(R, C, seed=1), rnd(R, C, seed=2), rnd(R, C, seed=3)
    outs = []
    for acts, il in ((a_gm, False), (a_il, True)):
        dc = rnd(R, C, seed=4).clone()
        dg = torch.empty(R, 4 * C, device=DEV, dtype=bf)
        ops.lstm_cell_bwd(dh, dc, acts, cp, cc, dg, R, C, acts_il=il)
        outs.append((dg.float(), dc))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_eval_metrics_kernel_matches_golden_and_oracle(golden_dir):
    """stfb_eval_metrics (argmax + confusion matrix + Dice counts in one pass) vs the fixture generated from the
    reference's ConfusionMatrix / DiceCoefficient (train_utils/train_and_eval.py:25-132): integer work, bit-exact."""
    import os
    import numpy as np
    from oracle import stf_oracle as O
    from stf_unet_b200.metrics import EvalMetrics, argmax_mask
    g = np.load(os.path.join(golden_dir, "eval_metrics_2x3x2x24x40.npz"))
    m = EvalMetrics(2, ignore_index=255, device=DEV)
    for i in range(2):
        logits, tgt = torch.from_numpy(g[f"logits{i}"]).to(DEV), torch.from_numpy(g[f"target{i}"]).to(DEV)
        mask = m.update({"out": logits}, tgt, want_mask=True)
        assert torch.equal(mask.cpu().long(), torch.from_numpy(g[f"logits{i}"]).argmax(1))
        assert np.array_equal(m.mat.cpu().numpy(), g[f"mat{i}"])
        assert np.allclose(m.cumulative_dice.cpu().numpy(), g[f"dice_cum{i}"], atol=1e-6)
    assert np.allclose(m.compute_dice().cpu().numpy(), g["dice"], atol=1e-6)
    acc_global, acc, iu = m.compute_confusion()
    assert np.allclose(acc_global.cpu().numpy(), g["acc_global"], atol=1e-6) and np.allclose(iu.cpu().numpy(), g["iu"], atol=1e-6)
    # larger random case with 4 classes against the oracle, ragged size, no ignore index
    gen = torch.Generator().manual_seed(3)
    logits = torch.randn(5, 4, 37, 53, generator=gen)
    tgt = torch.randint(0, 4, (5, 37, 53), generator=gen)
    tgt[0, :5] = 7                                       # out-of-range labels are skipped by the confusion matrix
    m4 = EvalMetrics(4, ignore_index=None, device=DEV)
    m4.update(logits.to(DEV), tgt.to(DEV))
    mat_ref, dice_ref = O.eval_metrics_batch(logits, tgt, 4, None)
    assert torch.equal(m4.mat.cpu(), mat_ref)
    assert torch.allclose(m4.compute_dice().cpu().double(), dice_ref, atol=1e-6)
    assert torch.equal(argmax_mask(logits.to(DEV)).cpu().long(), logits.argmax(1))
    # empty batch
    m4.update(torch.empty(0, 4, 8, 8, device=DEV), torch.empty(0, 8, 8, dtype=torch.int64, device=DEV))


@pytest.mark.parametrize("case", [(16, 32, 32, 64, 64, 3, 1, 8), (8, 16, 16, 128, 256, 3, 1, 4), (32, 8, 8, 128, 128, 3, 1, 8),
                                  (6, 25, 19, 64, 128, 3, 1, 3), (8, 32, 32, 64, 128, 3, 2, 2), (12, 16, 16, 64, 64, 1, 1, 4),
                                  (40, 32, 32, 64, 64, 3, 1, 5)])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_conv_fused_bn_statistics(case, pair, monkeypatch):
    """Train-mode BatchNorm statistics (per image group sum / sum of squares of the bf16 outputs) reduced in the conv
    epilogue == the separate bn_stats pass over the stored output (reference: nn.BatchNorm2d after every conv,
    src/stf_lstm_unet.py:14,17 ; one call per time step -> per-group statistics, :168-186)."""
    N, H, W, Cin, Cout, k, stride, G = case
    monkeypatch.setenv("STFB_HALO_PAIR", pair)
    dtype = torch.bfloat16
    pad = (k - 1) // 2
    x = nhwc(q(rnd(N, Cin, H, W, seed=1), dtype), dtype)
    w = q(rnd(Cout, Cin, k, k, seed=2, scale=(1.0 / (k * k * Cin)) ** 0.5), dtype)
    wp = ops.pack_weight(w.contiguous(), True, dtype, n_major=True)
    assert ops.conv_stats_fusable(x, Cout, k, stride, pad, G)
    partial = torch.zeros(ops.STAT_SLOTS, 2, G, Cout, device=DEV)
    y = ops.conv2d(x, wp, Cout, k, stride, pad, impl=ops.IMPL_TCGEN05, stat_partial=partial, stat_groups=G)
    y_plain = ops.conv2d(x, wp, Cout, k, stride, pad, impl=ops.IMPL_TCGEN05)
    assert torch.equal(y, y_plain)
    yf = y.float().view(G, -1, Cout)
    s_ref, q_ref = yf.sum(1), (yf * yf).sum(1)
    got = partial.double().sum(0)
    assert rel(got[0], s_ref) < 1e-5 and rel(got[1], q_ref) < 1e-5
    # and through the finalize kernel: same scale / shift as the two-pass route
    R = yf.shape[1]
    gamma, beta = rnd(Cout, seed=3).abs() + 0.5, rnd(Cout, seed=4)
    st_f = ops.bn_finalize_train(partial, gamma, beta, None, None, None, G, R, Cout)
    st_2 = ops.bn_finalize_train(ops.bn_stats(y, G, R, Cout), gamma, beta, None, None, None, G, R, Cout)
    assert rel(st_f[0], st_2[0]) < 1e-5 and (st_f[1] - st_2[1]).abs().max().item() < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,H,W", [(3, 4, 32, 48), (2, 3, 9, 7), (1, 1, 16, 16)])
def test_pack_series_u8_is_totensor_plus_normalize(B, T, H, W, dtype):
    """8-bit series -> normalised time-major NHWC in one pass; fp32 is bit-identical to the loader's arithmetic
    (ToTensor: x.float().div(255); Normalize: sub(mean).div(std); /root/reference/train.py:147-148)."""
    g = torch.Generator().manual_seed(B * 100 + H)
    u8 = torch.randint(0, 256, (B, T, H, W), generator=g, dtype=torch.uint8)
    u8[0, 0, 0, :4] = torch.tensor([0, 1, 254, 255], dtype=torch.uint8)
    ref = u8.float().div(255).sub(torch.tensor(0.709)).div(torch.tensor(0.127))          # [B,T,H,W] on the CPU
    ref = ref.permute(1, 0, 2, 3).reshape(T * B, H, W, 1)
    y = ops.pack_series_u8(u8.to(DEV), dtype, 0.709, 0.127)
    assert y.shape == (T * B, H, W, 1) and y.dtype == dtype
    if dtype == torch.float32:
        assert torch.equal(y.cpu(), ref)
    else:
        assert torch.equal(y.cpu(), ref.to(torch.bfloat16))
    y5 = ops.pack_series_u8(u8.to(DEV).unsqueeze(2), dtype, 0.709, 0.127)                  # [B,T,1,H,W] form
    assert torch.equal(y5, y)
    with pytest.raises(TypeError):
        ops.pack_series_u8(u8.to(DEV).float(), dtype, 0.709, 0.127)


@pytest.mark.parametrize("n", [1000003, 64, 5])
def test_adamw_flat_matches_torch_adamw(n):
    """stfb_adamw_flat against torch.optim.AdamW on the same flat tensor over several steps with a changing lr."""
    torch.manual_seed(n)
    p0 = torch.randn(n, device=DEV)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 8):
        lr = 1e-3 * (1.0 - step / 10.0)
        g = torch.randn(n, device=DEV) * (0.1 if step % 2 else 3.0)
        for grp in opt.param_groups:
            grp["lr"] = lr
        ref_p.grad = g.clone()
        opt.step()
        ops.adamw_flat_(p, g, m, v, lr, 0.9, 0.999, 1e-8, 1e-2, step)
        assert rel(p, ref_p.detach()) < 1e-6, step
    st = opt.state[ref_p]
    assert rel(m, st["exp_avg"]) < 1e-6 and rel(v, st["exp_avg_sq"]) < 1e-6
    # gradient pre-scaling (a summing all-reduce over 4 ranks): same as feeding g / 4
    p1, m1, v1 = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p2, m2, v2 = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    g = torch.randn(n, device=DEV)
    ops.adamw_flat_(p1, g, m1, v1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, grad_scale=0.25)
    ops.adamw_flat_(p2, g * 0.25, m2, v2, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)


@pytest.mark.parametrize("G,B,H,W,C,k,s,p", [(2, 2, 16, 16, 64, 3, 2, 1), (3, 1, 22, 18, 32, 3, 2, 1), (1, 2, 8, 12, 16, 2, 2, 0)])
def test_bn_relu_maxpool_one_pass_equals_two_launches(G, B, H, W, C, k, s, p):
    """The fused stem pass (BatchNorm apply + ReLU + max-pool + argmax index from the statistics slots) gives exactly the
    pooled values and indices of bn_apply_from_stats followed by maxpool_fwd_idx, and of torch on the same bf16 map."""
    N = G * B
    x = q(rnd(N, C, H, W, seed=1) * 1.5 + 0.3, torch.bfloat16)
    xn = nhwc(x, torch.bfloat16)
    gamma, beta = rnd(C, seed=2).abs() + 0.5, rnd(C, seed=3)
    gamma[::3] *= -1                                    # negative scales: max-pool does not commute with the affine map
    R = B * H * W
    xg = xn.float().view(G, R, C)
    slots = torch.zeros(ops.STAT_SLOTS, 2, G, C, device=DEV)
    for sl in range(ops.STAT_SLOTS):                    # spread the rows over the eight slots like the conv epilogue does
        part = xg[:, sl::ops.STAT_SLOTS]
        slots[sl, 0], slots[sl, 1] = part.sum(1), (part * part).sum(1)
    y_full = ops.bn_apply_from_stats(xn, slots, gamma, beta, G, R, C, True)
    y2, idx2 = ops.maxpool_fwd_idx(y_full, k, s, p)
    y1, idx1 = ops.bn_relu_maxpool_from_stats(xn, slots, gamma, beta, G, k, s, p)
    assert torch.equal(y1, y2) and torch.equal(idx1, idx2)
    ref = F.max_pool2d(nchw(y_full), k, s, p)
    assert torch.equal(nchw(y1), ref)
    # and the apply kernel's in-kernel coefficients are those of the finalize kernel
    st = ops.bn_finalize_train(slots, gamma, beta, None, None, None, G, R, C)
    assert torch.equal(ops.bn_apply(xn, st[0], st[1], G, R, C, True), y_full)


@pytest.mark.parametrize("mask", ["1", "2", "3"])
def test_programmatic_dependent_launch_switch_is_bit_identical(mask, monkeypatch):
    """STFB_PDL marks the convolution (1) / BatchNorm (2) launches for programmatic dependent launch: the kernels wait
    (griddepcontrol.wait) before they touch what their predecessor wrote, so a conv -> BN-apply -> conv chain must give the same
    bits with and without it."""
    dtype = torch.bfloat16
    x = nhwc(q(rnd(8, 64, 32, 32, seed=1), dtype), dtype)
    w1 = ops.pack_weight(q(rnd(64, 64, 3, 3, seed=2, scale=0.05), dtype).contiguous(), True, dtype, n_major=True)
    w2 = ops.pack_weight(q(rnd(128, 64, 3, 3, seed=3, scale=0.05), dtype).contiguous(), True, dtype, n_major=True)
    scale, shift = rnd(1, 64, seed=4).abs() + 0.5, rnd(1, 64, seed=5)

    def chain():
        y = x
        for _ in range(3):
            y = ops.conv2d(y, w1, 64, 3, 1, 1, impl=ops.IMPL_TCGEN05)
            y = ops.bn_apply(y, scale, shift, 1, 8 * 32 * 32, 64, relu=True)
        return ops.conv2d(y, w2, 128, 3, 1, 1, impl=ops.IMPL_TCGEN05)

    monkeypatch.setenv("STFB_PDL", "0")
    ref = chain()
    torch.cuda.synchronize()
    monkeypatch.setenv("STFB_PDL", mask)
    got = chain()
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
    assert torch.isfinite(got.float()).all() and got.float().abs().max() > 0
