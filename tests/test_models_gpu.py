"""GPU parity of the B200 models against the oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): fp32 logits rel-err <= 1e-4; bf16 <= 2e-2; identical argmax on >= 99.9 % of
pixels.  Gradients: fp32 rel-L2 <= 1e-3 per parameter (fp32 atomics reorder sums), bf16 checked on the loss
and on aggregate gradient direction."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import stf_unet_b200 as S  # noqa: E402
from oracle import stf_oracle as O  # noqa: E402
from oracle import weights as W  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def argmax_agree(a, b):
    return (a.argmax(1).cpu() == b.argmax(1).cpu()).float().mean().item()


def load_model(model, sd):
    model.load_state_dict(sd)
    return model.to(DEV)


_WARM = {}


def warm_stf_state(T=4, hw=64, steps=40):
    """Warm weights (SURVEY.md section 7 "hard parts" #1): a short fp32 AdamW run of the B200 model on structured
    synthetic DCE series from the reference's default init distributions.  bf16 parity bars are only meaningful on
    weights whose BatchNorm statistics match the data -- the reference's own bf16 autocast is 23 % away from its
    fp32 on cold weights.  (The fp32 training path used here is itself pinned to the golden fixtures above.)"""
    key = (T, hw, steps)
    if key not in _WARM:
        torch.manual_seed(0)
        m = S.STFLSTMUNet(1, 2, T).to(DEV)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
        x, t = W.synthetic_dce_batch(8, T, hw, hw, seed=77)
        x, t = x.to(DEV), t.to(DEV)
        m.train()
        for _ in range(steps):
            loss = S.criterion(m(x), t)
            opt.zero_grad()
            loss.backward()
            opt.step()
        _WARM[key] = ({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}, loss.item())
    return _WARM[key][0]


def oracle_bf16_floor(sd_dev, x, train):
    """The reference arithmetic's own bf16-autocast-vs-fp32 gap on the same weights (reported beside ours)."""
    with torch.no_grad():
        ref = O.stf_forward(sd_dev, x, train=train)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lo = O.stf_forward(sd_dev, x, train=train).float()
    return rel(lo, ref), argmax_agree(lo, ref)


def test_stf_eval_fp32_vs_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "stf_eval_b2_t3_64.npz"))
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=11)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    m = load_model(S.STFLSTMUNet(1, 2, 3), sd).eval()
    with torch.no_grad():
        y = m(x.to(DEV))["out"]
    assert y.shape == (2, 2, 32, 32) and y.dtype == torch.float32
    ref = torch.from_numpy(g["logits"])
    assert rel(y, ref) < 1e-4
    assert argmax_agree(y, ref) >= 0.999
    loss = S.criterion({"out": y}, t.to(DEV))
    assert abs(loss.item() - float(g["loss"])) < 1e-4


def test_stf_eval_ragged_size_uses_bilinear(golden_dir):
    g = np.load(os.path.join(golden_dir, "stf_eval_b1_t2_80.npz"))
    x, _ = W.synthetic_dce_batch(1, 2, 80, 80, seed=12)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    m = load_model(S.STFLSTMUNet(1, 2, 2), sd).eval()
    with torch.no_grad():
        y = m(x.to(DEV))["out"]
    assert y.shape == (1, 2, 40, 40)
    assert rel(y, torch.from_numpy(g["logits"])) < 1e-4


def test_stf_eval_bf16_autocast():
    x, _ = W.synthetic_dce_batch(4, 4, 128, 128, seed=31)
    sd = warm_stf_state()
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = O.stf_forward(sd_dev, x.to(DEV), train=False)
    floor = oracle_bf16_floor(sd_dev, x.to(DEV), False)
    m = load_model(S.STFLSTMUNet(1, 2, 4), sd).eval()
    with torch.no_grad():
        y32 = m(x.to(DEV))["out"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x.to(DEV))["out"]
    assert y.dtype == torch.float32
    r32, r, a = rel(y32, ref), rel(y, ref), argmax_agree(y, ref)
    print(f"stf eval warm: fp32 rel={r32:.3e} | bf16 rel={r:.3e} argmax={a:.5f} | torch-autocast floor rel={floor[0]:.3e} argmax={floor[1]:.5f}")
    assert r32 < 1e-4
    assert r < 2e-2 and a >= 0.999


def check_grads_vs_golden(g, grads, sd, x, t, kind):
    """fp32 gradients against the CPU-reference fixtures.  Train-mode BatchNorm on tiny batches amplifies fp32
    summation-order noise, so the bar is max(2e-3, 5 x floor) where floor is how far the SAME oracle arithmetic run
    through cuDNN on this GPU lands from the CPU fixture."""
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    _, _, og, _ = O.loss_and_grads(sd_dev, x.to(DEV), t.to(DEV), model=kind, train=True)
    scale = float(np.max(g["grad_norms"]))
    worst = (0.0, 0.0, "")
    for k in g.files:
        if not k.startswith("grad::"):
            continue
        ref = torch.from_numpy(g[k]).double()
        n = ref.norm().item()
        floor = (og[k[6:]].cpu().double() - ref).norm().item()
        err = (grads[k[6:]].cpu().double() - ref).norm().item()
        bar = max(2e-3 * n, 5 * floor) + 1e-6 * scale
        worst = max(worst, (err / max(n, 1e-12), floor / max(n, 1e-12), k))
        assert err < bar, (k, err / max(n, 1e-12), floor / max(n, 1e-12))
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        mine = grads[str(name)].double().norm().item()
        floor = abs(og[str(name)].double().norm().item() - norm)
        assert abs(mine - norm) < max(5e-3 * norm, 5 * floor) + 1e-6 * scale, (str(name), mine, norm, floor)
    print("worst full-grad rel err %.3e (cuDNN-oracle floor %.3e) at %s" % worst)


def _train_step(model, x, t):
    model.train()
    out = model(x)["out"]
    loss = S.criterion({"out": out}, t)
    loss.backward()
    return out.detach(), loss.detach()


def test_stf_train_fp32_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stf_train_b2_t3_64.npz"))
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=11)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    m = load_model(S.STFLSTMUNet(1, 2, 3), sd)
    out, loss = _train_step(m, x.to(DEV), t.to(DEV))
    assert rel(out, torch.from_numpy(g["logits"])) < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    grads = {n: p.grad for n, p in m.named_parameters()}
    check_grads_vs_golden(g, grads, sd, x, t, "stf")
    for k in g.files:
        if k.startswith("buf::"):
            mine = m.state_dict()[k[5:]].cpu()
            if g[k].dtype.kind == "i":
                assert int(mine) == int(g[k]), k
            else:
                assert rel(mine, torch.from_numpy(g[k])) < 1e-4, k


def test_stf_train_bf16_vs_oracle():
    B, T, HW = 4, 4, 128
    x, t = W.synthetic_dce_batch(B, T, HW, HW, seed=33)
    sd = warm_stf_state()
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd_dev, x.to(DEV), t.to(DEV), model="stf", train=True)
    floor = oracle_bf16_floor(sd_dev, x.to(DEV), True)
    m = load_model(S.STFLSTMUNet(1, 2, T), sd)
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x.to(DEV))["out"]
        loss = S.criterion({"out": out}, t.to(DEV))
    loss.backward()
    r, a = rel(out, ref_logits), argmax_agree(out, ref_logits)
    print(f"stf train bf16 warm: rel={r:.3e} argmax={a:.5f} loss {loss.item():.5f} vs {ref_loss.item():.5f} | "
          f"torch-autocast floor rel={floor[0]:.3e} argmax={floor[1]:.5f}")
    assert r < 2e-2 and a >= 0.999
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    num = den1 = den2 = 0.0
    for n, p in m.named_parameters():
        gr = ref_grads[n].double().flatten()
        gm = p.grad.double().flatten()
        num += (gr * gm).sum().item()
        den1 += (gr * gr).sum().item()
        den2 += (gm * gm).sum().item()
    cos = num / (den1 ** 0.5 * den2 ** 0.5)
    print("bf16 grad cosine", cos, "norm ratio", (den2 / den1) ** 0.5)
    assert cos > 0.98 and 0.9 < (den2 / den1) ** 0.5 < 1.1


@pytest.mark.parametrize("name,cin,c,hw,seed", [("unet_train_in1_c16_32", 1, 16, 32, 21), ("unet_train_in8_c8_48", 8, 8, 48, 22)])
def test_unet_train_fp32_vs_golden(golden_dir, name, cin, c, hw, seed):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x, t = W.synthetic_dce_batch(2, cin, hw, hw, seed=seed, half_res_target=False)
    sd = W.make_state_dict(W.unet_param_spec(cin, 2, c), seed=0)
    m = load_model(S.UNet(cin, 2, c), sd)
    out, loss = _train_step(m, x.view(2, cin, hw, hw).to(DEV), t.to(DEV))
    assert rel(out, torch.from_numpy(g["logits"])) < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    grads = {n: p.grad for n, p in m.named_parameters()}
    check_grads_vs_golden(g, grads, sd, x.view(2, cin, hw, hw), t, "unet")
    for k in g.files:
        if k.startswith("buf::"):
            mine = m.state_dict()[k[5:]].cpu()
            if g[k].dtype.kind == "i":
                assert int(mine) == int(g[k]), k
            else:
                assert rel(mine, torch.from_numpy(g[k])) < 1e-4, k


def test_unet_eval_fp32_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_eval_in1_c16_32.npz"))
    x, t = W.synthetic_dce_batch(2, 1, 32, 32, seed=21, half_res_target=False)
    sd = W.make_state_dict(W.unet_param_spec(1, 2, 16), seed=0)
    m = load_model(S.UNet(1, 2, 16), sd).eval()
    with torch.no_grad():
        y = m(x[:, :, 0].to(DEV))["out"]
    assert rel(y, torch.from_numpy(g["logits"])) < 1e-4


def test_unet_config1_fp32_vs_oracle():
    """BASELINE.json configs[0]: UNet(in=1, classes=2, base_c=64), B=4, 256x256, fp32 fwd + CE/Dice + bwd.
    The oracle runs on the host CPU here (as the reference's config does): cuDNN's fp32 wgrad is itself ~1e-2 away
    from the CPU result on the long pixel reductions (see the floors printed by the golden tests)."""
    x, t = W.synthetic_dce_batch(4, 1, 256, 256, seed=41, half_res_target=False)
    sd = W.make_state_dict(W.unet_param_spec(1, 2, 64), seed=0)
    xin = x[:, :, 0].contiguous()
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd, xin, t, model="unet", train=True)
    m = load_model(S.UNet(1, 2, 64), sd)
    out, loss = _train_step(m, xin.to(DEV), t.to(DEV))
    r, a = rel(out, ref_logits), argmax_agree(out, ref_logits)
    print(f"unet config1 fp32: rel={r:.3e} argmax={a:.5f}")
    assert r < 1e-4 and a >= 0.999
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item()) + 1e-5
    worst = (0.0, "")
    scale = max(g.double().norm().item() for g in ref_grads.values())
    for n, p in m.named_parameters():
        gr = ref_grads[n].double()
        err = (p.grad.cpu().double() - gr).norm().item()
        worst = max(worst, (err / max(gr.norm().item(), 1e-12), n))
        assert err < 2e-2 * gr.norm().item() + 1e-6 * scale, (n, err, gr.norm().item())
    print("unet config1 worst grad rel err %.3e at %s" % worst)


def test_training_loop_reduces_loss_bf16():
    """A few AdamW steps through the public API in bf16: loss must fall (the reference's own smoke criterion)."""
    torch.manual_seed(0)
    x, t = W.synthetic_dce_batch(4, 4, 64, 64, seed=51)
    m = S.STFLSTMUNet(1, 2, 4).to(DEV)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
    x, t = x.to(DEV), t.to(DEV)
    losses = []
    for _ in range(12):
        m.train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = S.criterion(m(x), t)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < 0.8 * losses[0]
    assert int(m.bn1.num_batches_tracked) == 12 * 4


def test_cpu_input_raises():
    m = S.UNet(1, 2, 8).to(DEV)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32))


def test_stf_pk_maps_eval_fp32_vs_golden(golden_dir):
    """use_pk_maps=True (reference :146-156, :172-174, :189-200): 4-channel stem + bilinear PK maps fused at 4 scales."""
    g = np.load(os.path.join(golden_dir, "stf_pk_eval_b1_t2_64.npz"))
    x, _ = W.synthetic_dce_batch(1, 5, 64, 64, seed=13)
    sd = W.make_state_dict(W.stf_param_spec(1, 2, use_pk_maps=True), seed=0)
    m = load_model(S.STFLSTMUNet(1, 2, 2, use_pk_maps=True), sd).eval()
    with torch.no_grad():
        y = m(x.to(DEV))["out"]
    assert y.shape == (1, 2, 32, 32)
    assert rel(y, torch.from_numpy(g["logits"])) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stf_pk_maps_train_vs_oracle(dtype):
    B, T = 2, 3
    x, t = W.synthetic_dce_batch(B, T + 3, 64, 64, seed=61)
    sd = W.make_state_dict(W.stf_param_spec(1, 2, use_pk_maps=True), seed=0)
    if dtype == torch.bfloat16:      # bf16 bars need BatchNorm statistics that match the data: a few fp32 steps first
        m0 = load_model(S.STFLSTMUNet(1, 2, T, use_pk_maps=True), sd)
        opt = torch.optim.AdamW(m0.parameters(), lr=1e-3)
        for _ in range(8):
            loss = S.criterion(m0(x.to(DEV)), t.to(DEV))
            opt.zero_grad()
            loss.backward()
            opt.step()
        sd = {k: v.detach().cpu().clone() for k, v in m0.state_dict().items()}
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd_dev, x.to(DEV), t.to(DEV), model="stf", train=True,
                                                          use_pk_maps=True)
    m = load_model(S.STFLSTMUNet(1, 2, T, use_pk_maps=True), sd)
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        out = m(x.to(DEV))["out"]
        loss = S.criterion({"out": out}, t.to(DEV))
    loss.backward()
    r = rel(out, ref_logits)
    print(f"stf pk-maps train {dtype}: rel={r:.3e} loss {loss.item():.5f} vs {ref_loss.item():.5f}")
    assert r < (1e-4 if dtype == torch.float32 else 2e-2)
    if dtype == torch.float32:
        for name in ("pk_fusion1.weight", "pk_fusion4.bias", "conv1.weight"):
            assert rel(dict(m.named_parameters())[name].grad, ref_grads[name]) < 2e-2, name
    else:   # bf16: individual small tensors are noisy at B=2; check the direction of the PK-branch gradients together
        num = d1 = d2 = 0.0
        for name, p in m.named_parameters():
            if name.startswith("pk_fusion") or name == "conv1.weight":
                a_, b_ = p.grad.double().flatten(), ref_grads[name].double().flatten()
                num += (a_ * b_).sum().item(); d1 += (a_ * a_).sum().item(); d2 += (b_ * b_).sum().item()
        cos = num / (d1 ** 0.5 * d2 ** 0.5)
        print("pk-branch bf16 grad cosine", cos)
        assert cos > 0.95


@pytest.mark.parametrize("fused_stats", [False, True])
def test_graphed_step_matches_eager(fused_stats, monkeypatch):
    """CUDA-graph replay of fwd + loss + bwd gives the eager step's loss and gradients, step after step.  With the
    BatchNorm statistics fused into the conv epilogue the per-channel sums are fp32 red.adds whose order differs from
    run to run, so that variant is compared at bf16 noise level instead of bit level."""
    from stf_unet_b200.graph import GraphedStep
    from stf_unet_b200 import engine, ops
    monkeypatch.setattr(engine, "USE_FUSED_BN_STATS", fused_stats)
    # the deterministic variant also takes the three-launch BatchNorm backward (the one-launch kernel adds its group sums
    # with fp32 atomics: same noise class as the fused statistics)
    monkeypatch.setattr(ops, "USE_FUSED_BN_BWD", fused_stats)
    ltol, gtol = (1e-2, 3e-2) if fused_stats else (1e-5, 1e-3)   # cold weights: the atomics-order noise is amplified
    torch.manual_seed(0)
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=71)
    x2, t2 = W.synthetic_dce_batch(2, 3, 64, 64, seed=72)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    a = load_model(S.STFLSTMUNet(1, 2, 3), sd)
    b = load_model(S.STFLSTMUNet(1, 2, 3), sd)
    gs = GraphedStep(a, S.criterion, x.to(DEV), t.to(DEV), warmup=2)
    b.load_state_dict(a.state_dict())          # warm-up + capture advanced a's BatchNorm buffers
    for xb, tb in ((x, t), (x2, t2)):
        la = gs(xb.to(DEV), tb.to(DEV))
        b.train()
        for p in b.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lb = S.criterion(b(xb.to(DEV)), tb.to(DEV))
        lb.backward()
        assert abs(la.item() - lb.item()) < ltol * max(1.0, abs(lb.item()))
        if fused_stats:
            continue      # B=2 puts 8..32 samples behind each deep BatchNorm: atomics-order noise is amplified chaotically there;
                          # gradients of the fused path are checked at a well-conditioned size in the next test
        for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert rel(pa.grad, pb.grad) < gtol, n
    assert int(a.bn1.num_batches_tracked) == int(b.bn1.num_batches_tracked)
    assert gs.launches_per_replay > 100


def test_fused_bn_statistics_match_two_pass_training(monkeypatch):
    """One bf16 training step (warm weights) with the BatchNorm statistics reduced in the conv epilogues vs the same step
    with the separate statistics pass: same loss, same running statistics, gradients aligned at bf16 noise level (each
    route is checked against the oracle in test_stf_train_bf16_vs_oracle)."""
    from stf_unet_b200 import engine
    B, T, HW = 4, 4, 128
    x, t = W.synthetic_dce_batch(B, T, HW, HW, seed=91)
    sd = warm_stf_state()
    outs = []
    for fused in (False, True):
        monkeypatch.setattr(engine, "USE_FUSED_BN_STATS", fused)
        m = load_model(S.STFLSTMUNet(1, 2, T), sd)
        m.train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = S.criterion(m(x.to(DEV)), t.to(DEV))
        loss.backward()
        outs.append((loss.item(), {n: p.grad.double().flatten() for n, p in m.named_parameters()},
                     m.layer2[1].bn1.running_var.clone(), int(m.bn1.num_batches_tracked)))
    (l0, g0, rv0, nb0), (l1, g1, rv1, nb1) = outs
    assert abs(l0 - l1) < 2e-3 * max(1.0, abs(l0))
    assert nb0 == nb1 and rel(rv1, rv0) < 2e-3
    num = sum((g0[n] * g1[n]).sum().item() for n in g0)
    d0 = sum((g0[n] * g0[n]).sum().item() for n in g0)
    d1 = sum((g1[n] * g1[n]).sum().item() for n in g0)
    cos = num / (d0 ** 0.5 * d1 ** 0.5)
    print("fused-vs-two-pass grad cosine", cos, "norm ratio", (d1 / d0) ** 0.5)
    assert cos > 0.99 and 0.95 < (d1 / d0) ** 0.5 < 1.05


def test_unet_bf16_train_vs_oracle():
    x, t = W.synthetic_dce_batch(4, 8, 64, 64, seed=81, half_res_target=False)
    xin = x.view(4, 8, 64, 64).to(DEV)
    torch.manual_seed(1)
    m0 = S.UNet(8, 2, 32).to(DEV)
    opt = torch.optim.AdamW(m0.parameters(), lr=1e-3)
    for _ in range(30):
        loss = S.criterion(m0(xin), t.to(DEV))
        opt.zero_grad()
        loss.backward()
        opt.step()
    sd_dev = {k: v.detach().clone() for k, v in m0.state_dict().items()}
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd_dev, xin, t.to(DEV), model="unet", train=True)
    m = S.UNet(8, 2, 32).to(DEV)
    m.load_state_dict(sd_dev)
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(xin)["out"]
        loss = S.criterion({"out": out}, t.to(DEV))
    loss.backward()
    r, a = rel(out, ref_logits), argmax_agree(out, ref_logits)
    print(f"unet bf16 train: rel={r:.3e} argmax={a:.5f}")
    assert r < 2e-2 and a >= 0.995


def test_stf_long_sequence_train_fp32_vs_oracle():
    """BASELINE.json configs[3] geometry in small (T=16 phases, non-square 128x96 slices): fp32 path against the oracle run
    on this GPU (forward logits, loss, every gradient).  The gradient bar is relative to a float64 run of the oracle:
    train-mode BatchNorm over two-sample batches amplifies fp32 summation-order noise (see check_grads_vs_golden), so
    the bar is max(2e-3, 5 x the distance of the oracle's own fp32 run from its float64 run)."""
    B, T, H, Wd = 2, 16, 128, 96
    x, t = W.synthetic_dce_batch(B, T, H, Wd, seed=41)
    sd = warm_stf_state()
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd_dev, x.to(DEV), t.to(DEV), model="stf", train=True)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_dev.items()}
    _, _, g64, _ = O.loss_and_grads(sd64, x.to(DEV).double(), t.to(DEV), model="stf", train=True)
    m = load_model(S.STFLSTMUNet(1, 2, T), sd)
    out, loss = _train_step(m, x.to(DEV), t.to(DEV))
    assert rel(out, ref_logits) < 1e-4                       # north_star: fp32 logits rel-err <= 1e-4
    assert abs(loss.item() - ref_loss.item()) < 1e-4
    assert int(m.state_dict()["bn1.num_batches_tracked"]) == int(sd["bn1.num_batches_tracked"]) + T

    def whole(gr):
        num = sum((gr[n].double() - g64[n]).pow(2).sum().item() for n in g64)
        return (num / sum(g64[n].pow(2).sum().item() for n in g64)) ** 0.5

    mine, floor = whole({n: p.grad for n, p in m.named_parameters()}), whole(ref_grads)
    print("long sequence: whole-gradient rel err vs float64 oracle: ours %.3e, fp32 oracle %.3e" % (mine, floor))
    assert mine < max(2e-3, 5 * floor)


def test_whole_volume_masks_match_per_batch_argmax():
    """configs[4]: a case of 11 slices through VolumePredictor (batches of 4, ragged tail, double-buffered host copies)
    gives the argmax masks of the model's own logits, slice by slice, and sharding by slice does not change them."""
    from stf_unet_b200.volume import VolumePredictor, predict_volume, slice_range
    Sn, T, HWs = 11, 3, 64
    series, _ = W.synthetic_dce_batch(Sn, T, HWs, HWs, seed=51)
    sd = warm_stf_state()
    m = load_model(S.STFLSTMUNet(1, 2, T), sd).eval()
    with torch.no_grad():
        logits = torch.cat([m(series[i:i + 4].to(DEV))["out"] for i in range(0, Sn, 4)])      # fp32 path
    ref = logits.argmax(1).to(torch.uint8).cpu()
    pred = VolumePredictor(m, tuple(series.shape[1:]), batch=4, autocast_dtype=None)
    masks = pred(series.pin_memory())
    assert masks.dtype == torch.uint8 and masks.shape == (Sn, HWs // 2, HWs // 2)
    assert torch.equal(masks, ref)
    assert torch.equal(pred(series), ref)                     # second case through the same graph; pageable host memory
    parts = []
    for r in range(3):                                        # three "ranks" on one device: same masks, no exchange
        (lo, hi), mk = predict_volume(m, series, batch=4, rank=r, world=3, autocast_dtype=None, predictor=pred)
        assert (lo, hi) == slice_range(Sn, r, 3) and mk.shape[0] == hi - lo
        parts.append(mk)
    assert torch.equal(torch.cat(parts), ref)
    # bf16 tensor-core path: the argmax agrees on >= 99.9 % of the pixels (north_star)
    pred16 = VolumePredictor(m, tuple(series.shape[1:]), batch=4)
    agree = (pred16(series) == ref).float().mean().item()
    print("volume bf16 argmax agreement", agree)
    assert agree >= 0.999


def test_uint8_series_input_matches_normalised_float_input():
    """raw 8-bit grey levels go in, the device normalises; logits equal those of the float path fed
    with the loader's ToTensor + Normalize output (fp32: bit-identical input, identical logits)."""
    from stf_unet_b200.synthetic import synthetic_dce_batch_u8
    u8, _ = synthetic_dce_batch_u8(2, 3, 64, 64, seed=61)
    xf = u8.float().div(255).sub(torch.tensor(0.709)).div(torch.tensor(0.127)).unsqueeze(2)      # [B,T,1,H,W]
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    m = load_model(S.STFLSTMUNet(1, 2, 3), sd).eval()
    with torch.no_grad():
        a = m(xf.to(DEV))["out"]
        b = m(u8.unsqueeze(2).to(DEV))["out"]
        ref = O.stf_forward({k: v.to(DEV) for k, v in sd.items()}, xf.to(DEV), train=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a16 = m(xf.to(DEV))["out"]
            c16 = m(u8.unsqueeze(2).to(DEV))["out"]
    assert torch.equal(a, b)
    assert rel(b, ref) < 1e-4
    assert torch.equal(c16, a16)          # bf16 path: the same fp32 value is rounded to bf16 once on either route


def test_flat_adamw_training_matches_torch_adamw():
    """FlatAdamW (parameters re-homed in one flat buffer, one launch per step) follows
    torch.optim.AdamW on the same model, same batches, through an LR schedule; state_dict keys are unchanged."""
    x, t = W.synthetic_dce_batch(2, 2, 64, 64, seed=71)
    x, t = x.to(DEV), t.to(DEV)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    ma, mb = load_model(S.STFLSTMUNet(1, 2, 2), sd), load_model(S.STFLSTMUNet(1, 2, 2), sd)
    keys = list(mb.state_dict().keys())
    oa = torch.optim.AdamW(ma.parameters(), lr=1e-3, weight_decay=1e-4)
    ob = S.FlatAdamW(mb, lr=1e-3, weight_decay=1e-4)
    assert list(mb.state_dict().keys()) == keys
    assert all(p.data_ptr() >= ob.flat_param.data_ptr() for p in mb.parameters())
    sa = torch.optim.lr_scheduler.LambdaLR(oa, lambda s: 1.0 / (1 + s))
    sb = torch.optim.lr_scheduler.LambdaLR(ob, lambda s: 1.0 / (1 + s))
    for _ in range(3):
        for mdl, opt, sch in ((ma, oa, sa), (mb, ob, sb)):
            mdl.train()
            opt.zero_grad(set_to_none=True)
            S.criterion(mdl(x), t).backward()
            opt.step()
            sch.step()
    worst = max(rel(pb.data, pa.data) for pa, pb in zip(ma.parameters(), mb.parameters()))
    print("FlatAdamW vs torch AdamW after 3 steps: worst parameter rel diff", worst)
    # the kernel itself matches torch to 1e-6 on identical gradients (test_adamw_flat_matches_torch_adamw); here the two
    # models' gradients differ by fp32 atomic-order noise, which Adam's sign-like first steps (update ~ lr * g / |g|)
    # turn into O(lr) differences on near-zero gradients (zero-initialised biases)
    assert worst < 5e-3
    # checkpoint round trip of the optimizer state
    st = ob.state_dict()
    oc = S.FlatAdamW(load_model(S.STFLSTMUNet(1, 2, 2), mb.state_dict()), lr=1.0)
    oc.load_state_dict(st)
    assert oc.steps == 3 and torch.equal(oc.exp_avg, ob.exp_avg) and oc.param_groups[0]["lr"] == ob.param_groups[0]["lr"]


def test_eval_weight_packs_are_cached_until_a_parameter_changes():
    """Eval-mode forwards reuse the packed bf16 operands (no pack launch) until a weight moves: torch-side updates are seen
    through Parameter._version, the raw-pointer FlatAdamW step through the module's pack epoch."""
    from stf_unet_b200 import _lib
    x, t = W.synthetic_dce_batch(2, 2, 64, 64, seed=81)
    x, t = x.to(DEV), t.to(DEV)
    m = load_model(S.STFLSTMUNet(1, 2, 2), warm_stf_state()).eval()

    def fwd():
        n0 = _lib.launch_count()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)["out"]
        return y, _lib.launch_count() - n0

    fwd()                                   # learns the pack plan
    y1, n1 = fwd()                          # packs through the plan
    y2, n2 = fwd()                          # reuses
    assert n2 == n1 - 1 and torch.equal(y1, y2)
    with torch.no_grad():
        m.final.bias.add_(1.0)              # torch-side update: version bump
    y3, n3 = fwd()
    assert n3 == n1 and rel(y3, y1 + 1.0) < 1e-3
    opt = S.FlatAdamW(m, lr=1e-2, weight_decay=0.0)      # re-homes the parameters (new addresses: plan rebuilt)
    fwd(); fwd()
    y4, n4 = fwd()
    assert n4 == n1 - 1
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        S.criterion(m(x), t).backward()
    opt.step()                              # raw-pointer update
    m.eval()
    y5, n5 = fwd()
    assert n5 == n1 and not torch.equal(y5, y4)
    y6, n6 = fwd()
    assert n6 == n1 - 1 and torch.equal(y6, y5)


# ---------------------------------------------------------------------------------------------------------------------
# bf16 parity AT THE BENCHMARKED SIZES (BASELINE.json configs[1..3]) against the oracle run in fp32 (TF32 off) through cuDNN on
# the same GPU.  These sizes engage code paths the small cases never reach: the two-kernel BatchNorm backward for maps
# > 48 MB, 57-patch split-K weight gradients, several tiles per persistent CTA, CTA pairs on every 3x3 layer.
# ---------------------------------------------------------------------------------------------------------------------
def _grad_report(named_params, ref_grads, floor_grads=None):
    """-> (cosine, norm ratio, worst per-parameter rel-L2 among parameters carrying >= 1e-3 of the gradient norm, its name,
    the same worst figure for `floor_grads` = torch's own bf16-autocast gradients of the oracle)."""
    num = den1 = den2 = 0.0
    per, per_floor = {}, {}
    for n, p in named_params:
        gr = ref_grads[n].double().flatten()
        gm = p.grad.double().flatten()
        num += (gr * gm).sum().item()
        den1 += (gr * gr).sum().item()
        den2 += (gm * gm).sum().item()
        per[n] = ((gm - gr).norm().item(), gr.norm().item())
        if floor_grads is not None:
            per_floor[n] = ((floor_grads[n].double().flatten() - gr).norm().item(), gr.norm().item())
    total = den1 ** 0.5
    big = [n for n, (_, nr) in per.items() if nr >= 1e-3 * total]
    worst = max(big, key=lambda n: per[n][0] / per[n][1])
    wf = max((per_floor[n][0] / per_floor[n][1] for n in big), default=float("nan")) if per_floor else float("nan")
    return num / (den1 ** 0.5 * den2 ** 0.5), (den2 / den1) ** 0.5, per[worst][0] / per[worst][1], worst, wf


def _train_parity_case(B, T, HW, seed, check_buffers=True):
    x, t = W.synthetic_dce_batch(B, T, HW, HW, seed=seed)
    x, t = x.to(DEV), t.to(DEV)
    sd = warm_stf_state()
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    ref_logits, ref_loss, ref_grads, ref_bufs = O.loss_and_grads(sd_dev, x, t, model="stf", train=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):          # the reference arithmetic's own bf16 gap: the noise floor
        fl_logits, fl_loss, fl_grads, _ = O.loss_and_grads(sd_dev, x, t, model="stf", train=True)
    fl_grads = {k: (None if v is None else v.float()) for k, v in fl_grads.items()}
    floor = (rel(fl_logits.float(), ref_logits), argmax_agree(fl_logits.float(), ref_logits))
    torch.cuda.empty_cache()
    m = load_model(S.STFLSTMUNet(1, 2, T), sd)
    m.train()
    nbt0 = int(m.bn1.num_batches_tracked)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)["out"]
        loss = S.criterion({"out": out}, t)
    loss.backward()
    r, a = rel(out, ref_logits), argmax_agree(out, ref_logits)
    cos, ratio, worst, wname, wfloor = _grad_report(list(m.named_parameters()), ref_grads, fl_grads)
    print(f"train bf16 B={B} T={T} {HW}^2: logits rel={r:.3e} argmax={a:.5f} loss {loss.item():.6f} vs {ref_loss.item():.6f} | "
          f"grad cos={cos:.5f} norm ratio={ratio:.4f} worst param rel={worst:.3e} ({wname}) | torch-autocast floor: logits "
          f"rel={floor[0]:.3e} argmax={floor[1]:.5f} worst param rel={wfloor:.3e}")
    assert r < 2e-2 and a >= 0.999                               # north_star: bf16 <= 2e-2, argmax >= 99.9 %
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    assert cos > 0.99 and 0.95 < ratio < 1.05
    assert worst < max(5e-2, 2.0 * wfloor), (wname, worst, wfloor)   # no parameter may be worse than twice torch's own bf16
    if check_buffers:
        assert int(m.bn1.num_batches_tracked) == nbt0 + T            # T BatchNorm calls per forward (stf_lstm_unet.py:168-186)
        assert int(m.decoder2.res_conv.conv_block["1"].num_batches_tracked) == int(sd["decoder2.res_conv.conv_block.1.num_batches_tracked"]) + 1
        mine = m.state_dict()
        for k, v in ref_bufs.items():
            if k.endswith("num_batches_tracked"):
                assert int(mine[k]) == int(v), k
            else:
                assert rel(mine[k], v) < 2e-2, (k, rel(mine[k], v))
    return r, a


def test_bench_size_eval_bf16_vs_oracle():
    """BASELINE.json configs[1]: eval forward, batch 16 x T=8 x 256x256, bf16 vs the fp32 oracle on warm weights."""
    x, _ = W.synthetic_dce_batch(16, 8, 256, 256, seed=1234)
    x = x.to(DEV)
    sd = warm_stf_state()
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = O.stf_forward(sd_dev, x, train=False)
    floor = oracle_bf16_floor(sd_dev, x, False)
    m = load_model(S.STFLSTMUNet(1, 2, 8), sd).eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)["out"]
    r, a = rel(y, ref), argmax_agree(y, ref)
    print(f"configs[1] eval bf16: rel={r:.3e} argmax={a:.5f} | torch-autocast floor rel={floor[0]:.3e} argmax={floor[1]:.5f}")
    assert y.shape == (16, 2, 128, 128)
    assert r < 2e-2 and a >= 0.999


def test_bench_size_train_step_bf16_vs_oracle():
    """BASELINE.json configs[2], one GPU's shard: 16 slices x T=8 x 256x256, train-mode forward + CE/Dice + backward in
    bf16 vs the fp32 oracle: logits, argmax, loss, the whole gradient (cosine / norm / worst parameter against torch's own
    bf16-autocast floor), every BatchNorm buffer, num_batches_tracked += 8."""
    _train_parity_case(16, 8, 256, 1234)


def test_config3_train_step_bf16_vs_oracle():
    """BASELINE.json configs[3]: T=16 phases x 512x512 (B=2 here: the oracle keeps every fp32 activation alive)."""
    _train_parity_case(2, 16, 512, 4321)


# ---------------------------------------------------------------------------------------------------------------------
# round-1 advisor findings, pinned
# ---------------------------------------------------------------------------------------------------------------------
def _bf16_grads(model, x, t):
    for p in model.parameters():
        p.grad = None
    model.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = S.criterion(model(x), t)
    loss.backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()}


def test_wgrad_side_streams_do_not_share_partial_tiles(monkeypatch):
    """Halo weight gradients write per-split partial tiles that a second kernel sums; launches rotate over three side streams
    that are ordered only against the main stream.  Every stream owns its own scratch slot (csrc/wgrad_tc.cu): the gradients with
    three side streams equal the single-stream gradients (deterministic BatchNorm routes; the partial sums meet through fp32
    red.adds, so equality is to rounding, 1e-5 -- an overwritten partial tile shows up as O(1))."""
    from stf_unet_b200 import engine, ops
    monkeypatch.setattr(engine, "USE_FUSED_BN_STATS", False)
    monkeypatch.setattr(ops, "USE_FUSED_BN_BWD", False)
    x, t = W.synthetic_dce_batch(8, 4, 128, 128, seed=81)
    x, t = x.to(DEV), t.to(DEV)
    sd = warm_stf_state()
    m = load_model(S.STFLSTMUNet(1, 2, 4), sd)
    monkeypatch.setattr(engine, "USE_WGRAD_STREAM", False)
    ref = _bf16_grads(m, x, t)
    monkeypatch.setattr(engine, "USE_WGRAD_STREAM", True)
    for _ in range(3):                                      # several runs: the race was timing dependent
        m.load_state_dict(sd)
        got = _bf16_grads(m, x, t)
        for n in ref:
            assert rel(got[n], ref[n]) < 1e-5, n


def test_two_forwards_before_backward_and_gradient_accumulation():
    """engine.ModelFunction scatters into THE flat buffer of its own forward (not the latest one) and accumulates into the buffer
    the parameters' .grad view when gradients already exist (zero_grad(set_to_none=False), micro-batches)."""
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    xa, ta = (v.to(DEV) for v in W.synthetic_dce_batch(2, 2, 64, 64, seed=91))
    xb, tb = (v.to(DEV) for v in W.synthetic_dce_batch(2, 2, 64, 64, seed=92))
    m = load_model(S.STFLSTMUNet(1, 2, 2), sd).train()

    def grads_of(x, t):
        m.load_state_dict(sd)
        for p in m.parameters():
            p.grad = None
        S.criterion(m(x), t).backward()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    ga, gb = grads_of(xa, ta), grads_of(xb, tb)
    # two forwards, then both backwards: the sum of the two gradients (running stats are irrelevant to train-mode gradients)
    m.load_state_dict(sd)
    for p in m.parameters():
        p.grad = None
    la = S.criterion(m(xa), ta)
    lb = S.criterion(m(xb), tb)
    la.backward()
    lb.backward()
    flat = m._last_flat_grad
    for n, p in m.named_parameters():
        assert rel(p.grad, ga[n] + gb[n]) < 2e-4, n
    assert m.conv1.weight.grad.data_ptr() == flat.data_ptr()          # .grad still views the flat buffer the optimizer steps on
    # zero_grad(set_to_none=False) keeps the views; the next backward adds into the same buffer
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    opt.zero_grad(set_to_none=False)
    S.criterion(m(xa), ta).backward()
    assert m._last_flat_grad is flat
    for n, p in m.named_parameters():
        assert rel(p.grad, ga[n]) < 2e-4, n


def test_graphed_step_rebinds_grads_for_any_optimizer():
    """Graph replay runs no Python; GraphedStep re-binds the .grad views after every replay, so torch.optim.AdamW with its
    default zero_grad(set_to_none=True) between steps keeps training (it silently skipped every parameter before)."""
    from stf_unet_b200.graph import GraphedStep
    torch.manual_seed(0)
    x, t = (v.to(DEV) for v in W.synthetic_dce_batch(2, 2, 64, 64, seed=95))
    m = S.STFLSTMUNet(1, 2, 2).to(DEV)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True)
    g = GraphedStep(m, S.criterion, x, t, warmup=2)
    w0 = m.final.weight.detach().clone()
    losses = []
    for _ in range(4):
        opt.zero_grad()                                   # set_to_none=True
        losses.append(g(x, t).item())
        assert m.final.weight.grad is not None
        opt.step()
    assert not torch.equal(w0, m.final.weight) and losses[-1] < losses[0]


def test_parameter_hooks_are_refused():
    """torch DDP / register_hook rely on per-parameter autograd hooks that never fire here (one autograd node): refuse loudly."""
    m = S.STFLSTMUNet(1, 2, 2).to(DEV).train()
    m.final.weight.register_hook(lambda g: g)
    x, _ = W.synthetic_dce_batch(1, 2, 64, 64, seed=3)
    with pytest.raises(RuntimeError, match="hooks"):
        m(x.to(DEV))


def test_eval_forward_with_grad_enabled_warns_once():
    """eval mode is the folded-BatchNorm inference path with no tape: the result carries no graph, and the module says so."""
    from stf_unet_b200 import modules as M
    M.B200Module._warned_eval_grad = False
    m = S.STFLSTMUNet(1, 2, 2).to(DEV).eval()
    x, _ = W.synthetic_dce_batch(1, 2, 64, 64, seed=3)
    with pytest.warns(UserWarning, match="NON-differentiable"):
        out = m(x.to(DEV))
    assert not out["out"].requires_grad
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("error")
        m(x.to(DEV))                      # second call: silent
        with torch.no_grad():
            m(x.to(DEV))
