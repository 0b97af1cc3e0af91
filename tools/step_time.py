"""Time the CUDA-graph training step (fwd + CE/Dice + bwd, no optimizer) — a quick A/B harness for kernel changes.

    python tools/step_time.py [--iters 20] [--batch 16] [--T 8] [--hw 256] [--fp32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stf_unet_b200 as S                       # noqa: E402
from stf_unet_b200.graph import GraphedStep     # noqa: E402
from stf_unet_b200.synthetic import synthetic_dce_batch   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--fp32", action="store_true", help="fp32 mode (no autocast): split-precision tensor-core path unless STFB_NO_SPLIT_FP32=1")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = S.STFLSTMUNet(1, 2, a.T).to(dev)
    x, t = synthetic_dce_batch(a.batch, a.T, a.hw, a.hw, seed=1234)
    x, t = x.to(dev), t.to(dev)
    g = GraphedStep(model, S.criterion, x, t, autocast_dtype=None if a.fp32 else torch.bfloat16)
    for _ in range(3):
        g(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(a.iters):
            g(x, t)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.iters)
    print(f"step_time: {best:.3f} ms/step  ({a.batch / best * 1e3:.1f} slices/s)  loss {g.loss.item():.5f}  "
          f"launches/replay {g.launches_per_replay}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
