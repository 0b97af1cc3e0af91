"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every case it (1) builds the reference nn.Module, (2) checks the oracle's
state_dict spec against the module's own keys/shapes, (3) loads the deterministic
weights of oracle.weights, (4) runs the reference forward / criterion / backward on
a seeded synthetic DCE series and (5) stores logits, loss, per-parameter gradient
norms, a few full gradients and the updated BN buffers in one .npz per case.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from src import STFLSTMUNet, UNet  # noqa: E402  (reference)
from train_utils.train_and_eval import criterion as ref_criterion  # noqa: E402

from oracle import weights as W  # noqa: E402

FULL_GRADS_STF = ["conv1.weight", "bn1.weight", "bn1.bias", "layer2.0.downsample.0.weight",
                  "lstm1.weight_hh_l0", "lstm1.bias_ih_l0", "lstm4.bias_hh_l0", "decoder2.up.bias",
                  "decoder2.fusion.weight", "upconv1.weight", "final.weight", "final.bias",
                  "final_res.conv_block.4.weight"]
BUFFERS_STF = ["bn1.running_mean", "bn1.running_var", "bn1.num_batches_tracked",
               "layer3.0.downsample.1.running_var", "layer4.2.bn2.running_mean",
               "decoder3.res_conv.conv_block.4.running_var", "decoder3.res_conv.conv_block.4.num_batches_tracked"]
FULL_GRADS_UNET = ["enc1.0.weight", "enc1.0.bias", "enc1.1.weight", "bottleneck.3.bias", "up2.weight", "up2.bias",
                   "dec1.0.weight", "out_conv.weight", "out_conv.bias"]
BUFFERS_UNET = ["enc1.1.running_mean", "enc1.1.running_var", "enc1.1.num_batches_tracked", "dec4.4.running_var"]


def check_spec(module, spec):
    ref = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    mine = {n: tuple(s) for n, s, _ in spec}
    assert list(ref.keys()) == [n for n, _, _ in spec], "state_dict key order/name mismatch"
    assert ref == mine, "state_dict shape mismatch"


def run_case(name, module, spec, x, target, train, full_grads, buffers, seed=0):
    check_spec(module, spec)
    sd = W.make_state_dict(spec, seed=seed)
    module.load_state_dict(sd)
    out = {"x_checksum": np.float64(x.double().sum().item())}
    if train:
        module.train()
        logits = module(x)["out"]
        loss = ref_criterion({"out": logits}, target)
        loss.backward()
        out["loss"] = np.float64(loss.item())
        names, norms = [], []
        for k, p in module.named_parameters():
            names.append(k)
            norms.append(0.0 if p.grad is None else p.grad.double().norm().item())
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms, dtype=np.float64)
        for k in full_grads:
            out["grad::" + k] = dict(module.named_parameters())[k].grad.numpy().copy()
        new_sd = module.state_dict()
        for k in buffers:
            out["buf::" + k] = new_sd[k].numpy().copy()
    else:
        module.eval()
        with torch.no_grad():
            logits = module(x)["out"]
        if target is not None:
            out["loss"] = np.float64(ref_criterion({"out": logits}, target).item())
    out["logits"] = logits.detach().numpy().copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: logits {tuple(logits.shape)} std={logits.std().item():.4f} "
          f"argmax1={(logits.argmax(1) == 1).float().mean().item():.3f} "
          f"loss={out.get('loss', float('nan')):.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    # --- STF, 64x64, B=2, T=3 -------------------------------------------------------
    spec = W.stf_param_spec(1, 2)
    x, t = W.synthetic_dce_batch(2, 3, 64, 64, seed=11)
    run_case("stf_eval_b2_t3_64", STFLSTMUNet(1, 2, 3), spec, x, t, False, [], [])
    run_case("stf_train_b2_t3_64", STFLSTMUNet(1, 2, 3), spec, x, t, True, FULL_GRADS_STF, BUFFERS_STF)
    # --- STF, ragged size 80x80 (bilinear fallback in the decoder, :56-57) --------------
    x, t = W.synthetic_dce_batch(1, 2, 80, 80, seed=12)
    run_case("stf_eval_b1_t2_80", STFLSTMUNet(1, 2, 2), spec, x, None, False, [], [])
    # --- STF with PK maps (use_pk_maps=True; 3 extra "time steps" carry the maps) ----------
    spec_pk = W.stf_param_spec(1, 2, use_pk_maps=True)
    x, t = W.synthetic_dce_batch(1, 2 + 3, 64, 64, seed=13)
    run_case("stf_pk_eval_b1_t2_64", STFLSTMUNet(1, 2, 2, use_pk_maps=True), spec_pk, x, t, False, [], [])
    # --- UNet in=1 base_c=16, 32x32 and in=8 (flat_channels) ---------------------------------
    spec_u = W.unet_param_spec(1, 2, 16)
    x, t = W.synthetic_dce_batch(2, 1, 32, 32, seed=21, half_res_target=False)
    run_case("unet_eval_in1_c16_32", UNet(1, 2, 16), spec_u, x[:, :, 0], t, False, [], [])
    run_case("unet_train_in1_c16_32", UNet(1, 2, 16), spec_u, x[:, :, 0], t, True, FULL_GRADS_UNET, BUFFERS_UNET)
    spec_u8 = W.unet_param_spec(8, 2, 8)
    x, t = W.synthetic_dce_batch(2, 8, 48, 48, seed=22, half_res_target=False)
    run_case("unet_train_in8_c8_48", UNet(8, 2, 8), spec_u8, x.view(2, 8, 48, 48), t, True, FULL_GRADS_UNET, BUFFERS_UNET)
    # --- loss alone on random logits (criterion, train_and_eval.py:299-313) ----------------------
    g = np.random.Generator(np.random.PCG64(5))
    logits = torch.from_numpy(g.standard_normal((3, 2, 24, 40)).astype(np.float32) * 2).requires_grad_(True)
    tgt = torch.from_numpy((g.uniform(size=(3, 24, 40)) < 0.3).astype(np.int64))
    loss = ref_criterion({"out": logits}, tgt)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "criterion_3x2x24x40.npz"), logits=logits.detach().numpy(),
                        target=tgt.numpy(), loss=np.float64(loss.item()), grad=logits.grad.numpy())
    print("criterion:", loss.item())


def metrics_case():
    """evaluate()'s metric classes (train_utils/train_and_eval.py:25-132) on two random batches, one holding 255s."""
    from train_utils.train_and_eval import ConfusionMatrix, DiceCoefficient
    g = np.random.Generator(np.random.PCG64(9))
    conf, dice = ConfusionMatrix(2), DiceCoefficient(num_classes=2, ignore_index=255)
    out = {}
    for i in range(2):
        logits = torch.from_numpy(g.standard_normal((3, 2, 24, 40)).astype(np.float32))
        logits[0, :, 3, 5:9] = 0.25                       # exact ties: argmax must take the first class
        tgt = torch.from_numpy((g.uniform(size=(3, 24, 40)) < 0.35).astype(np.int64))
        if i == 1:
            tgt[g.uniform(size=(3, 24, 40)) < 0.1] = 255   # ignored pixels
        conf.update(tgt.flatten(), logits.argmax(1).flatten())
        dice.update(logits, tgt)
        out[f"logits{i}"], out[f"target{i}"] = logits.numpy(), tgt.numpy()
        out[f"mat{i}"] = conf.mat.numpy().copy()
        out[f"dice_cum{i}"] = dice.cumulative_dice.numpy().copy()
    out["dice"] = dice.compute().numpy()
    acc_global, acc, iu = conf.compute()
    out["acc_global"], out["acc"], out["iu"] = acc_global.numpy(), acc.numpy(), iu.numpy()
    np.savez_compressed(os.path.join(HERE, "eval_metrics_2x3x2x24x40.npz"), **out)
    print("metrics:", out["mat1"].tolist(), out["dice"].tolist())


def criterion_options_case():
    """criterion with the arguments the reference's signature accepts beyond its training defaults
    (train_utils/train_and_eval.py:299-313): class weights, ignore_index = 255 (the value collate_fn pads targets with,
    my_dataset.py:243), dice on / off; one image is ignored entirely (the `sets_sum == 0` branch,
    dice_coefficient_loss.py:34-35)."""
    g = np.random.Generator(np.random.PCG64(15))
    out = {}
    base = torch.from_numpy(g.standard_normal((4, 3, 20, 28)).astype(np.float32) * 2)
    tgt = torch.from_numpy(g.integers(0, 3, size=(4, 20, 28)).astype(np.int64))
    tgt[torch.from_numpy(g.uniform(size=(4, 20, 28)) < 0.15)] = 255
    tgt[3] = 255                                       # a fully ignored image
    w = torch.tensor([0.2, 1.0, 2.5])
    out["logits"], out["target"], out["weight"] = base.numpy(), tgt.numpy(), w.numpy()
    for tag, kw in (("w_ign_dice", dict(loss_weight=w, num_classes=3, dice=True, ignore_index=255)),
                    ("ign_dice", dict(loss_weight=None, num_classes=3, dice=True, ignore_index=255)),
                    ("w_ign_nodice", dict(loss_weight=w, num_classes=3, dice=False, ignore_index=255))):
        logits = base.clone().requires_grad_(True)
        loss = ref_criterion({"out": logits}, tgt, **kw)
        loss.backward()
        out["loss_" + tag] = np.float64(loss.item())
        out["grad_" + tag] = logits.grad.numpy()
        print("criterion", tag, loss.item())
    np.savez_compressed(os.path.join(HERE, "criterion_options_4x3x20x28.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        metrics_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "criterion":
        criterion_options_case()
    else:
        main()
        metrics_case()
        criterion_options_case()
