"""Synthetic DCE-MRI series for benches and smoke runs (there are no BreaDM files here).

Product-side generator: ``bench.py``'s own arm and the tools draw their batches from here, never from ``oracle/``.
It follows the recipe of SURVEY 8(d) draw for draw (numpy PCG64, seed 1234 + rank), so the oracle's generator and this
one give identical batches (``tests/test_host_cpu.py`` holds them against each other).

Per sample: smooth anatomy ``0.5 + 0.2 sin(x/17 + phase) cos(y/23)``, a tumour disk (centre in the central half,
radius in [H/16, H/6]) with wash-in ``0.4 (1 - e^{-t/2})`` over a global ``0.05 (1 - e^{-t/4})``, Gaussian noise
sigma 0.02, clamp to [0, 1], then the reference's normalisation ``(img - 0.709) / 0.127``
(/root/reference/train.py:147-148).  The target is the disk; STF-LSTM-UNet's logits are half resolution
(/root/reference/src/stf_lstm_unet.py:245-256), so its targets are the nearest ``[::2, ::2]`` subsample.
"""
from __future__ import annotations

import numpy as np
import torch

MEAN, STD = 0.709, 0.127


def _series(batch, T, H, W, seed, channels):
    """Images in [0, 1] as float32 [B,T,C,H,W] and the full-resolution disk masks [B,H,W]."""
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    rows, cols = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    times = np.arange(T, dtype=np.float32)
    imgs = np.empty((batch, T, channels, H, W), dtype=np.float32)
    disks = np.empty((batch, H, W), dtype=np.int64)
    for b in range(batch):
        # draw order matters (phase, centre y, centre x, radius, then one noise field per (t, c))
        phase = rng.uniform(0, 2 * np.pi)
        anatomy = 0.5 + 0.2 * np.sin(cols / 17.0 + phase) * np.cos(rows / 23.0)
        cy, cx = rng.uniform(H * 0.25, H * 0.75), rng.uniform(W * 0.25, W * 0.75)
        rad = rng.uniform(H / 16.0, H / 6.0)
        disk = ((rows - cy) ** 2 + (cols - cx) ** 2 <= rad * rad).astype(np.float32)
        for t in range(T):
            enhanced = disk * 0.4 * (1 - np.exp(-times[t] / 2.0)) + 0.05 * (1 - np.exp(-times[t] / 4.0))
            for c in range(channels):
                noisy = anatomy + enhanced + rng.standard_normal((H, W)).astype(np.float32) * 0.02
                imgs[b, t, c] = np.clip(noisy, 0.0, 1.0)
        disks[b] = disk.astype(np.int64)
    return imgs, disks


def synthetic_dce_batch(batch, T, H, W, seed=1234, half_res_target=True, channels=1):
    """-> (x [B,T,C,H,W] float32 normalised, target int64 {0,1} [B,H/2,W/2] or [B,H,W])."""
    imgs, disks = _series(batch, T, H, W, seed, channels)
    x = (imgs - np.float32(MEAN)) / np.float32(STD)
    tgt = disks[:, ::2, ::2] if half_res_target else disks
    return torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(tgt))


def synthetic_dce_batch_u8(batch, T, H, W, seed=1234, half_res_target=True):
    """The same series as 8-bit grey levels [B,T,H,W] uint8 (what the reference's loader reads from PNG files,
    /root/reference/my_dataset.py:143-232) + targets: the input of the device-side normalisation kernel."""
    imgs, disks = _series(batch, T, H, W, seed, 1)
    u8 = np.clip(np.rint(imgs[:, :, 0] * 255.0), 0, 255).astype(np.uint8)
    tgt = disks[:, ::2, ::2] if half_res_target else disks
    return torch.from_numpy(u8), torch.from_numpy(np.ascontiguousarray(tgt))
