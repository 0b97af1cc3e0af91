"""Golden fixture for the Tofts PK fitting kernels, generated from the UNMODIFIED reference class
(/root/reference/pk_fitting.py ToftsModelFitter) in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tofts.py

pk_fitting.py imports matplotlib at module level (for its debug plots); matplotlib is not installed here and none of the
functions exercised below touch it, so an empty stand-in module is registered before the import.  Nothing else is patched.

Stored: (1) extended_tofts_model_batch on 257 random parameter triples; (2) fit_volume_gpu on a synthetic 8-phase 80x80
subtraction series with three tissue blobs of different enhancement kinetics -> the tissue mask, the fitted maps, and the
same fit stopped after 5 epochs (to pin the early Adam dynamics as well as the end point)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
import pk_fitting as ref  # noqa: E402


def synthetic_subtraction_series(H=80, W=80, T=8, seed=3):
    """uint8 [T, H, W]: dark background, three elliptical tissue regions whose curves follow the Tofts model with different
    (Ktrans, ve, vp) plus noise; values scaled into 8 bits the way the reference's PNG inputs are."""
    g = np.random.Generator(np.random.PCG64(seed))
    fitter = ref.ToftsModelFitter(device=torch.device("cpu"))
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((T, H, W), dtype=np.float32)
    regions = [((25, 23), (17, 14), (0.25, 0.30, 0.05)), ((54, 52), (20, 19), (0.08, 0.15, 0.02)), ((20, 60), (11, 12), (0.6, 0.45, 0.12))]
    for (cy, cx), (ry, rx), (k, e, v) in regions:
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        n = int(m.sum())
        kk = torch.tensor(np.clip(k * (1 + 0.2 * g.standard_normal(n)), 0.01, 0.9), dtype=torch.float32)
        ee = torch.tensor(np.clip(e * (1 + 0.2 * g.standard_normal(n)), 0.02, 0.49), dtype=torch.float32)
        vv = torch.tensor(np.clip(v * (1 + 0.2 * g.standard_normal(n)), 0.0, 0.19), dtype=torch.float32)
        curves = fitter.extended_tofts_model_batch(fitter.time_points, kk, ee, vv).numpy()       # [n, T]
        img[:, m] = curves.T + 0.12                      # baseline so that phase 0 passes the tissue threshold
    img += 0.004 * g.standard_normal(img.shape).astype(np.float32)
    return np.clip(np.round(img * 255.0 / 0.9), 0, 255).astype(np.uint8)


def main():
    torch.manual_seed(0)
    dev = torch.device("cpu")
    fitter = ref.ToftsModelFitter(device=dev)
    g = np.random.Generator(np.random.PCG64(7))
    out = {}
    k = torch.tensor(g.uniform(0.0, 1.0, 257), dtype=torch.float32)
    e = torch.tensor(g.uniform(0.001, 0.5, 257), dtype=torch.float32)
    v = torch.tensor(g.uniform(0.0, 0.2, 257), dtype=torch.float32)
    out["fwd_k"], out["fwd_ve"], out["fwd_vp"] = k.numpy(), e.numpy(), v.numpy()
    out["fwd_out"] = fitter.extended_tofts_model_batch(fitter.time_points, k, e, v).numpy()
    mod = ref.ToftsModelFitter(device=dev, aif_method="modified")
    out["fwd_out_modified"] = mod.extended_tofts_model_batch(mod.time_points, k, e, v).numpy()
    short = torch.tensor([0.0, 0.5, 1.5, 3.0], dtype=torch.float32)
    out["fwd_t_short"] = short.numpy()
    out["fwd_out_short"] = fitter.extended_tofts_model_batch(short, k, e, v).numpy()

    series = synthetic_subtraction_series()
    out["series"] = series
    _, mask = fitter.preprocess_images(series)
    out["tissue_mask"] = mask.numpy()
    print("valid pixels:", int(mask.sum()))
    ref.tqdm = lambda it, **kw: it                          # silence the progress bars (the iterable is returned unchanged)
    maps = fitter.fit_volume_gpu(series)
    out["maps_100"] = maps.astype(np.float32)
    # the same optimisation stopped after 5 epochs: the reference hard-codes num_epochs = 100, so replay its loop verbatim
    # through the oracle restatement and ALSO check that restatement against the 100-epoch reference result here
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import tofts_oracle as TO
    images_tensor, _ = fitter.preprocess_images(series)
    valid = images_tensor.permute(1, 2, 0).reshape(-1, 8)[mask.reshape(-1)]
    k100, e100, v100, losses = TO.fit_pixels(fitter.time_points, valid, epochs=100)
    chk = np.zeros((3, mask.numel()), dtype=np.float32)
    chk[0, mask.reshape(-1).numpy()], chk[1, mask.reshape(-1).numpy()], chk[2, mask.reshape(-1).numpy()] = k100, e100, v100
    assert np.array_equal(chk.reshape(maps.shape), maps), "oracle restatement != reference fit_volume_gpu"
    out["losses_100"] = losses.numpy()
    k5, e5, v5, _ = TO.fit_pixels(fitter.time_points, valid, epochs=5)
    out["fit5"] = np.stack([k5.numpy(), e5.numpy(), v5.numpy()])
    np.savez_compressed(os.path.join(HERE, "tofts_fit_80x80.npz"), **out)
    print("saved; final loss", float(losses[-1]), "mean Ktrans", float(k100.mean()))


if __name__ == "__main__":
    main()
