"""Where do per-parameter fill kernels come from? (profiles one graph capture and one optimizer step)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stf_unet_b200 as S
from stf_unet_b200.synthetic import synthetic_dce_batch
from stf_unet_b200.graph import GraphedStep
from torch.profiler import profile, ProfilerActivity

dev = "cuda"
model = S.STFLSTMUNet(1, 2, 8).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
x, t = synthetic_dce_batch(2, 8, 64, 64, seed=1, half_res_target=True)
x, t = x.to(dev), t.to(dev)


def report(prof, title):
    print("====", title)
    rows = [e for e in prof.key_averages(group_by_stack_n=6) if e.key in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like")]
    rows.sort(key=lambda e: -e.count)
    for e in rows[:8]:
        print(e.key, e.count)
        for s in e.stack[:6]:
            print("     ", s)


with profile(activities=[ProfilerActivity.CPU], with_stack=True) as prof:
    gs = GraphedStep(model, S.criterion, x, t, warmup=1)
report(prof, "GraphedStep construction (1 warm-up + capture)")
gs(x, t); opt.step()
with profile(activities=[ProfilerActivity.CPU], with_stack=True) as prof:
    gs(x, t)
    opt.step()
report(prof, "replay + opt.step()")
