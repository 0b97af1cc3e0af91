"""One launch each of the kernels either side of the hot path (for ncu captures): augmentation, Tofts forward + fit, evaluation
metrics.   python tools/aux_probe.py"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stf_unet_b200.augment import PairedAugment          # noqa: E402
from stf_unet_b200.metrics import EvalMetrics            # noqa: E402
from stf_unet_b200.pk_fitting import ToftsModelFitter    # noqa: E402
from stf_unet_b200.synthetic import synthetic_dce_batch_u8  # noqa: E402

dev = torch.device("cuda")
u8, tgt = synthetic_dce_batch_u8(16, 8, 256, 256, seed=1)
masks = torch.from_numpy(np.kron(tgt.numpy().astype(np.uint8), np.ones((2, 2), dtype=np.uint8)))
aug = PairedAugment(train=True, target_stride=2, rng=random.Random(0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record()
    x, t = aug(u8.to(dev), masks.to(dev))
    e1.record()
    torch.cuda.synchronize()
print(f"augment 16 x 8 x 256^2 -> 224^2: {e0.elapsed_time(e1) * 1e3:.1f} us (incl. host planning + table upload)")
f = ToftsModelFitter(device=dev)
px = torch.rand(60000, 8, device=dev) * 0.5
for i in range(2):
    e0.record()
    k, e, v = f.fit_pixels(px)
    e1.record()
    torch.cuda.synchronize()
print(f"tofts fit 60000 pixels x 100 epochs: {e0.elapsed_time(e1):.2f} ms")
out = f.extended_tofts_model_batch(f.time_points, k, e, v)
m = EvalMetrics(2, ignore_index=255, device=dev)
logits = torch.randn(16, 2, 128, 128, device=dev)
m.update(logits, tgt.to(dev), want_mask=True)
torch.cuda.synchronize()
print("ok", float(out.mean()))
