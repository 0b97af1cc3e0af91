"""Kernel timeline of the CUDA-graph training step (torch.profiler / CUPTI): every kernel of two replays with its stream, start
and duration -> gpurun_out/timeline.csv, plus a per-stream summary.  The per-launch ncu list is serialised; this one shows
what actually overlaps and where the main chain waits.

    python tools/timeline.py [--batch 16] [--T 8] [--hw 256]
"""
import argparse
import csv
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stf_unet_b200 as S                       # noqa: E402
from stf_unet_b200.graph import GraphedStep     # noqa: E402
from stf_unet_b200.synthetic import synthetic_dce_batch   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = S.STFLSTMUNet(1, 2, a.T).to(dev)
    opt = S.FlatAdamW(model, lr=1e-3, weight_decay=1e-4)
    x, t = synthetic_dce_batch(a.batch, a.T, a.hw, a.hw, seed=1234)
    x, t = x.to(dev), t.to(dev)
    g = GraphedStep(model, S.criterion, x, t)
    for _ in range(5):
        g(x, t)
        opt.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            g(x, t)
            opt.step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range is not None:
            rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, "device_index", 0), ev.name))
    rows.sort()
    t0 = rows[0][0] if rows else 0
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    # stream ids: torch's FunctionEvent does not expose them; take them from the chrome trace instead
    trace = a.out.replace(".csv", ".json")
    prof.export_chrome_trace(trace)
    import json
    ev = json.load(open(trace))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["start_us", "dur_us", "stream", "name"])
        for e in ks:
            w.writerow([round(e["ts"] - t0, 3), round(e["dur"], 3), e.get("args", {}).get("stream", e.get("tid")), e["name"][:160]])
    os.remove(trace)
    print(f"timeline: {len(ks)} device activities over {(ks[-1]['ts'] + ks[-1]['dur'] - t0) / 1e3:.3f} ms -> {a.out}")


if __name__ == "__main__":
    main()
