"""CPU: the oracle restatement reproduces the fixtures generated from the live reference
(tests/golden/make_golden.py).  Tolerances allow for oneDNN picking different ISA paths
on different hosts; they are ~100x tighter than the product's fp32 parity bar (1e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import weights as W
from oracle import stf_oracle as O


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_stf_eval(golden_dir):
    g = load(golden_dir, "stf_eval_b2_t3_64")
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=11)
    assert abs(x.double().sum().item() - float(g["x_checksum"])) < 1e-6 * abs(float(g["x_checksum"])) + 1e-6
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    with torch.no_grad():
        y = O.stf_forward(sd, x, train=False)
    assert rel(y.numpy(), g["logits"]) < 1e-5
    assert abs(O.criterion(y, t).item() - float(g["loss"])) < 1e-5


def test_stf_eval_ragged_bilinear(golden_dir):
    g = load(golden_dir, "stf_eval_b1_t2_80")
    x, _ = W.synthetic_dce_batch(1, 2, 80, 80, seed=12)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    with torch.no_grad():
        y = O.stf_forward(sd, x, train=False)
    assert y.shape == (1, 2, 40, 40)
    assert rel(y.numpy(), g["logits"]) < 1e-5


def test_stf_pk_maps(golden_dir):
    g = load(golden_dir, "stf_pk_eval_b1_t2_64")
    x, _ = W.synthetic_dce_batch(1, 5, 64, 64, seed=13)
    sd = W.make_state_dict(W.stf_param_spec(1, 2, use_pk_maps=True), seed=0)
    with torch.no_grad():
        y = O.stf_forward(sd, x, train=False, use_pk_maps=True)
    assert rel(y.numpy(), g["logits"]) < 1e-5


def _check_train(g, logits, loss, grads, bufs):
    assert rel(logits.numpy(), g["logits"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        mine = grads[str(name)].double().norm().item()
        assert abs(mine - norm) <= 2e-4 * norm + 1e-6, (name, mine, norm)
    for k in g.files:
        if k.startswith("grad::"):
            # conv biases in front of a train-mode BN have a mathematically zero gradient
            err = np.linalg.norm(grads[k[6:]].numpy().astype(np.float64) - g[k])
            assert err < 2e-4 * np.linalg.norm(g[k]) + 1e-6, k
        if k.startswith("buf::"):
            if g[k].dtype.kind == "i":
                assert int(bufs[k[5:]]) == int(g[k]), k
            else:
                assert rel(bufs[k[5:]].numpy(), g[k]) < 1e-5, k


def test_stf_train(golden_dir):
    g = load(golden_dir, "stf_train_b2_t3_64")
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=11)
    sd = W.make_state_dict(W.stf_param_spec(1, 2), seed=0)
    logits, loss, grads, bufs = O.loss_and_grads(sd, x, t, model="stf", train=True)
    _check_train(g, logits, loss, grads, bufs)
    # the encoder BNs are called T=3 times per forward
    assert int(bufs["bn1.num_batches_tracked"]) == 3
    assert int(bufs["decoder3.res_conv.conv_block.4.num_batches_tracked"]) == 1


@pytest.mark.parametrize("name,cin,c,hw,seed", [("unet_train_in1_c16_32", 1, 16, 32, 21), ("unet_train_in8_c8_48", 8, 8, 48, 22)])
def test_unet_train(golden_dir, name, cin, c, hw, seed):
    g = load(golden_dir, name)
    x, t = W.synthetic_dce_batch(2, cin, hw, hw, seed=seed, half_res_target=False)
    sd = W.make_state_dict(W.unet_param_spec(cin, 2, c), seed=0)
    logits, loss, grads, bufs = O.loss_and_grads(sd, x.view(2, cin, hw, hw), t, model="unet", train=True)
    _check_train(g, logits, loss, grads, bufs)


def test_unet_eval(golden_dir):
    g = load(golden_dir, "unet_eval_in1_c16_32")
    x, t = W.synthetic_dce_batch(2, 1, 32, 32, seed=21, half_res_target=False)
    sd = W.make_state_dict(W.unet_param_spec(1, 2, 16), seed=0)
    with torch.no_grad():
        y = O.unet_forward(sd, x[:, :, 0], train=False)
    assert rel(y.numpy(), g["logits"]) < 1e-5
    assert abs(O.criterion(y, t).item() - float(g["loss"])) < 1e-5


def test_criterion(golden_dir):
    g = load(golden_dir, "criterion_3x2x24x40")
    logits = torch.from_numpy(g["logits"]).requires_grad_(True)
    loss = O.criterion(logits, torch.from_numpy(g["target"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    assert rel(logits.grad.numpy(), g["grad"]) < 1e-5


@pytest.mark.parametrize("tag,kw", [("w_ign_dice", dict(weighted=True, dice=True)), ("ign_dice", dict(weighted=False, dice=True)),
                                    ("w_ign_nodice", dict(weighted=True, dice=False))])
def test_criterion_options(golden_dir, tag, kw):
    """class weights / ignore_index=255 / dice off, one image ignored entirely: fixture from the reference's criterion."""
    g = load(golden_dir, "criterion_options_4x3x20x28")
    logits = torch.from_numpy(g["logits"]).requires_grad_(True)
    w = torch.from_numpy(g["weight"]) if kw["weighted"] else None
    loss = O.criterion(logits, torch.from_numpy(g["target"]), loss_weight=w, dice=kw["dice"], ignore_index=255)
    loss.backward()
    assert abs(loss.item() - float(g["loss_" + tag])) < 1e-6
    assert rel(logits.grad.numpy(), g["grad_" + tag]) < 1e-5


def test_eval_metrics_oracle_matches_reference_classes(golden_dir):
    """oracle.eval_metrics_batch == ConfusionMatrix / DiceCoefficient of the live reference (fixture from make_golden.py)."""
    g = load(golden_dir, "eval_metrics_2x3x2x24x40")
    mat = torch.zeros(2, 2, dtype=torch.int64)
    cum = torch.zeros(2, dtype=torch.float64)
    for i in range(2):
        m, d = O.eval_metrics_batch(torch.from_numpy(g[f"logits{i}"]), torch.from_numpy(g[f"target{i}"]), 2, 255)
        mat += m
        cum += d
        assert np.array_equal(mat.numpy(), g[f"mat{i}"])
        assert np.allclose(cum.numpy(), g[f"dice_cum{i}"], rtol=0, atol=1e-6)
    assert np.allclose((cum / 2).numpy(), g["dice"], atol=1e-6)


def test_tofts_oracle_matches_reference_class(golden_dir):
    """oracle.tofts_oracle == ToftsModelFitter (pk_fitting.py:193-231, :288-368): fixture from the live reference class."""
    from oracle import tofts_oracle as TO
    g = load(golden_dir, "tofts_fit_80x80")
    t = torch.arange(8, dtype=torch.float32)
    k, e, v = (torch.from_numpy(g[n]) for n in ("fwd_k", "fwd_ve", "fwd_vp"))
    assert np.array_equal(TO.extended_tofts_model_batch(t, k, e, v).numpy(), g["fwd_out"])
    short = torch.from_numpy(g["fwd_t_short"])
    assert np.array_equal(TO.extended_tofts_model_batch(short, k, e, v).numpy(), g["fwd_out_short"])
    mask = torch.from_numpy(g["tissue_mask"]).reshape(-1)
    assert int(mask.sum()) > 1024                       # more than one Adam batch: the zero-gradient momentum steps are exercised
    valid = (torch.from_numpy(g["series"]).float() / 255.0).permute(1, 2, 0).reshape(-1, 8)[mask]
    kk, ee, vv, losses = TO.fit_pixels(t, valid, epochs=100)
    maps = g["maps_100"].reshape(3, -1)[:, mask.numpy()]
    assert np.array_equal(np.stack([kk.numpy(), ee.numpy(), vv.numpy()]), maps)
    assert np.allclose(losses.numpy(), g["losses_100"], rtol=0, atol=0)


def test_augment_oracle_matches_reference_pipeline(golden_dir):
    """oracle.augment_oracle (numpy restatement of Pillow / torchvision arithmetic + the reference's draw order) against the
    fixture generated from /root/reference/transforms.py composed as train.py:58-67: every sample bit for bit."""
    import random
    from oracle import augment_oracle as AO
    g = load(golden_dir, "augment_6x3x256")
    u8, masks = AO.fixture_inputs()
    assert AO.digest(u8) == str(g["series_digest"]) and AO.digest(masks) == str(g["masks_digest"])
    for b, seed in enumerate(g["seeds"]):
        p = AO.draw_params(random.Random(int(seed)))
        x, t = AO.apply(u8[b], masks[b], p)
        assert AO.digest(x) == str(g[f"xdigest{b}"]) and AO.digest(t) == str(g[f"tdigest{b}"]), b
        if b == 1:
            assert np.array_equal(x, g["x1"]) and np.array_equal(t, g["t1"].astype(np.int64))
