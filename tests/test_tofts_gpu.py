"""Tofts PK fitting kernels (csrc/tofts.cu) against the fixture generated from the reference's ToftsModelFitter
(tests/golden/make_golden_tofts.py) and against the oracle restatement on fresh inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from stf_unet_b200.pk_fitting import ToftsModelFitter  # noqa: E402

DEV = torch.device("cuda")


def rel(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "tofts_fit_80x80.npz"))


def test_forward_model_matches_reference(g):
    """extended_tofts_model_batch: fp32, tolerance 2e-5 relative (the reference sums up to 700 exp terms per point with
    torch.sum's tree order, the kernel sequentially) -- default grid, the bi-exponential AIF, and a short irregular grid
    whose first point precedes every convolution sample (stays exactly 0)."""
    k, e, v = (torch.from_numpy(g[n]).to(DEV) for n in ("fwd_k", "fwd_ve", "fwd_vp"))
    f = ToftsModelFitter(device=DEV)
    out = f.extended_tofts_model_batch(f.time_points, k, e, v)
    assert out.shape == (257, 8) and out.dtype == torch.float32
    assert rel(out.cpu(), g["fwd_out"]) < 2e-5
    assert float(out[:, 0].abs().max()) == 0.0
    assert (out.cpu() - torch.from_numpy(g["fwd_out"])).abs().max().item() < 2e-5 * float(np.abs(g["fwd_out"]).max())
    m = ToftsModelFitter(device=DEV, aif_method="modified")
    assert rel(m.extended_tofts_model_batch(m.time_points, k, e, v).cpu(), g["fwd_out_modified"]) < 2e-5
    short = torch.from_numpy(g["fwd_t_short"]).to(DEV)
    assert rel(f.extended_tofts_model_batch(short, k, e, v).cpu(), g["fwd_out_short"]) < 2e-5
    # empty batch
    z = torch.empty(0, device=DEV)
    assert f.extended_tofts_model_batch(f.time_points, z, z, z).shape == (0, 8)


def test_fit_matches_reference_fit_volume_gpu(g):
    """The whole fit_volume_gpu pipeline on the fixture's 8-phase series: same tissue mask, and the fitted maps within
    fp32-Adam tolerance of the reference's after 100 epochs (2 batches of 1024 + a ragged one: pixels keep moving on the
    steps of the other batches).  Adam divides by sqrt(v): a few pixels sitting on a clamp or with a vanishing gradient
    amplify rounding, so the bar is on the bulk (99 % of pixels within 2e-3 absolute) and on the mean."""
    f = ToftsModelFitter(device=DEV)
    _, mask = f.preprocess_images(g["series"])
    assert np.array_equal(mask.cpu().numpy(), g["tissue_mask"])
    maps = f.fit_volume_gpu(g["series"])
    assert maps.shape == (3, 80, 80) and maps.dtype == np.float32
    ref = g["maps_100"]
    assert np.array_equal(maps[:, ~g["tissue_mask"]], np.zeros_like(maps[:, ~g["tissue_mask"]]))
    d = np.abs(maps - ref)[:, g["tissue_mask"]]
    frac_close = float((d < 2e-3).mean())
    print("tofts fit: max abs diff", d.max(), "frac within 2e-3:", frac_close, "mean |diff|", d.mean())
    assert frac_close >= 0.99
    assert d.mean() < 2e-4
    for i in range(3):
        assert abs(maps[i][g["tissue_mask"]].mean() - ref[i][g["tissue_mask"]].mean()) < 1e-4


def test_early_epochs_and_loss_curve(g):
    """Five epochs (tight: rounding has had no time to diverge) and the per-epoch mean batch loss over 100 epochs."""
    f = ToftsModelFitter(device=DEV)
    mask = torch.from_numpy(g["tissue_mask"]).reshape(-1)
    valid = (torch.from_numpy(g["series"]).float() / 255.0).permute(1, 2, 0).reshape(-1, 8)[mask].to(DEV)
    k, e, v = f.fit_pixels(valid, epochs=5)
    got = torch.stack([k, e, v]).cpu().numpy()
    assert np.abs(got - g["fit5"]).max() < 2e-5
    _, _, _, losses = f.fit_pixels(valid, epochs=100, return_losses=True)
    assert np.allclose(losses.cpu().numpy(), g["losses_100"], rtol=2e-3, atol=1e-7)
    # ragged single batch and a batch size that does not divide N
    from oracle import tofts_oracle as TO
    sub = valid[:333].cpu()
    ko, eo, vo, _ = TO.fit_pixels(f.time_points.cpu(), sub, epochs=8, batch_size=100)
    kg, eg, vg = f.fit_pixels(sub.to(DEV), epochs=8, batch_size=100)
    assert (kg.cpu() - ko).abs().max() < 5e-5 and (eg.cpu() - eo).abs().max() < 5e-5 and (vg.cpu() - vo).abs().max() < 5e-5


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        ToftsModelFitter(device="cpu")
    f = ToftsModelFitter(device=DEV)
    with pytest.raises(RuntimeError, match="CUDA"):
        f.extended_tofts_model_batch(f.time_points, torch.zeros(3), torch.ones(3), torch.zeros(3))
