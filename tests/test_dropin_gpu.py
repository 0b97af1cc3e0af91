"""The drop-in boundary, executed: the reference's own UNMODIFIED driver functions (train_one_epoch / evaluate,
/root/reference/train_utils/train_and_eval.py:316-411) drive the B200 modules exactly as they drive the reference's, and the
two runs are held against each other.  north_star: "the model modules keep their constructor and forward() signatures so
train.py, val.py, test.py and train_and_eval.py drive the new path unchanged"; precision is signalled through the driver's own
`torch.amp.autocast(enabled=scaler is not None)`: scaler=None -> fp32-accurate kernels,
set_autocast_dtype(bf16) + GradScaler(enabled=False) -> the tcgen05 family."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import stf_unet_b200 as S  # noqa: E402
from oracle import weights as W  # noqa: E402

DEV = torch.device("cuda")


def _loader(n, B, T, hw, unet=False):
    """A list is all the driver needs of a DataLoader (__len__ + iteration, train_and_eval.py:252,384): CPU batches
    (image [B,T,1,H,W] float, target int64) like collate_fn's; STF targets are half resolution."""
    out = []
    for i in range(n):
        x, t = W.synthetic_dce_batch(B, T, hw, hw, seed=300 + i, half_res_target=not unet)
        out.append((x, t))
    return out


def _drive(ref, model, loader, scaler, epochs=1):
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)       # train.py:230-237
    sched = ref.te.create_lr_scheduler(opt, len(loader), epochs, warmup=True)                  # train.py:245-247
    losses = []
    for ep in range(epochs):
        mean_loss, lr = ref.te.train_one_epoch(model, opt, loader, DEV, ep, 2, lr_scheduler=sched, print_freq=100, scaler=scaler)
        losses.append(mean_loss)
    ev = ref.te.evaluate(model, loader, DEV, 2)
    return losses, lr, ev


def _same_start(ref_model, ours):
    ours.load_state_dict(ref_model.state_dict())          # the checkpoint contract: identical keys / shapes (Appendix B)
    return ours


@pytest.mark.parametrize("family", ["stf", "unet"])
def test_reference_driver_runs_the_b200_modules_fp32(reference, family, capsys):
    """scaler=None: the driver's autocast context is disabled -> fp32-accurate kernels.  Loss trajectory, learning rate,
    evaluate()'s confusion matrix / Dice and the trained weights agree with the reference module driven the same way."""
    torch.manual_seed(0)
    if family == "stf":
        ref_model = reference.STFLSTMUNet(1, 2, 3).to(DEV)
        ours = _same_start(ref_model, S.STFLSTMUNet(1, 2, 3).to(DEV))
        loader = _loader(5, 4, 3, 64)
    else:
        ref_model = reference.UNet(8, 2, 16).to(DEV)
        ours = _same_start(ref_model, S.UNet(8, 2, 16).to(DEV))
        loader = _loader(5, 2, 8, 48, unet=True)          # [B, 8, 1, H, W]: preprocess_input flattens it for UNet (:12-14)
        assert ours.input_format == ref_model.input_format == "flat_channels"
    l_ref, lr_ref, ev_ref = _drive(reference, ref_model, loader, None, epochs=2)
    l_our, lr_our, ev_our = _drive(reference, ours, loader, None, epochs=2)
    capsys.readouterr()
    assert lr_our == lr_ref
    for a, b in zip(l_our, l_ref):
        assert abs(a - b) <= 2e-2 * abs(b), (l_our, l_ref)
    assert abs(ev_our["dice"] - ev_ref["dice"]) < 2e-2
    assert abs(ev_our["global_accuracy"] - ev_ref["global_accuracy"]) < 1e-2
    mat_o, mat_r = ev_our["confusion_matrix"].mat.cpu(), ev_ref["confusion_matrix"].mat.cpu()
    assert mat_o.sum() == mat_r.sum() and (mat_o - mat_r).abs().sum().item() <= 0.01 * mat_r.sum().item()
    # the trained state is a valid checkpoint for the reference module and vice versa (train.py:304-311, test.py:142-146)
    ref_model.load_state_dict(ours.state_dict())
    for k, v in ours.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref_model.state_dict()[k])


def test_reference_driver_bf16_through_its_own_autocast(reference, capsys):
    """bf16: the driver opens torch.amp.autocast(device_type='cuda', enabled=scaler is not None) itself; with the autocast
    dtype set to bfloat16 and a disabled GradScaler the B200 module picks the tcgen05 family, the reference module runs
    torch's own bf16 autocast -- same driver code for both."""
    old = torch.get_autocast_dtype("cuda")
    torch.set_autocast_dtype("cuda", torch.bfloat16)
    try:
        torch.manual_seed(0)
        ref_model = reference.STFLSTMUNet(1, 2, 4).to(DEV)
        ours = _same_start(ref_model, S.STFLSTMUNet(1, 2, 4).to(DEV))
        loader = _loader(6, 8, 4, 128)
        from stf_unet_b200 import _lib
        scaler = torch.amp.GradScaler("cuda", enabled=False)
        l_ref, _, ev_ref = _drive(reference, ref_model, loader, scaler, epochs=2)
        n0 = _lib.launch_count()
        l_our, _, ev_our = _drive(reference, ours, loader, scaler, epochs=2)
        assert _lib.launch_count() - n0 > 1000, "the CUDA path did not run"
        capsys.readouterr()
        # two bf16 implementations on cold weights: trajectories agree to a few per cent and both learn
        for a, b in zip(l_our, l_ref):
            assert abs(a - b) <= 0.1 * abs(b), (l_our, l_ref)
        assert l_our[1] < l_our[0]
        assert abs(ev_our["dice"] - ev_ref["dice"]) < 0.1
    finally:
        torch.set_autocast_dtype("cuda", old)


def test_reference_criterion_accepts_b200_logits_and_matches_fused_one(reference):
    """The reference's own criterion (PyTorch ops, B*C host syncs) on the B200 module's logits == stf_unet_b200.criterion,
    forward and gradient into the model (one autograd node)."""
    torch.manual_seed(1)
    m = S.STFLSTMUNet(1, 2, 2).to(DEV).train()
    x, t = W.synthetic_dce_batch(2, 2, 64, 64, seed=5)
    x, t = x.to(DEV), t.to(DEV)
    la = reference.te.criterion(m(x), t)
    la.backward()
    ga = m.final.weight.grad.clone()
    for p in m.parameters():
        p.grad = None
    m.bn1.num_batches_tracked.zero_()
    lb = S.criterion(m(x), t)
    lb.backward()
    assert abs(la.item() - lb.item()) < 1e-5
    assert ((ga - m.final.weight.grad).norm() / ga.norm()).item() < 1e-4
