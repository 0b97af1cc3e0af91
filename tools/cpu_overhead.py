"""How long does the host take to ENQUEUE one training step (no sync) vs. how long the GPU takes to run it?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stf_unet_b200 as S
from stf_unet_b200 import _lib
import bench

dev = "cuda"
model = S.STFLSTMUNet(1, 2, 8).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
x, t = bench.make_batch(0)
x, t = x.to(dev), t.to(dev)

def step():
    model.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = S.criterion(model(x), t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
for trial in range(3):
    n0 = _lib.launch_count()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0):7.2f} ms   enqueue+drain {1e3 * (t2 - t0):7.2f} ms   launches {_lib.launch_count() - n0}")
