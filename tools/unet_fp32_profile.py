"""Per-family CUDA-event profile of one eager fp32 UNet(1, 2, 64) train step at batch 4 (bench.py --workload unet's step)."""
import sys
import torch
sys.path.insert(0, ".")
import stf_unet_b200 as S
from stf_unet_b200 import ops
from stf_unet_b200.synthetic import synthetic_dce_batch

dev = torch.device("cuda", 0)
x, t = synthetic_dce_batch(4, 1, 256, 256, seed=1234, half_res_target=False)
x, t = x[:, 0].contiguous().to(dev), t.to(dev)
torch.manual_seed(0)
model = S.UNet(1, 2, 64).to(dev).train()


def step():
    loss = S.criterion(model(x), t)
    model.zero_grad(set_to_none=True)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
if "--ncu" in sys.argv:          # ncu --profile-from-start off: exactly one warm step inside the profiler range
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
prof = ops.KernelProfiler()
ops.set_profiler(prof)
step()
ops.set_profiler(None)
fam = prof.summary()
tot = sum(d["ms"] for d in fam.values())
print(f"families (serialised, one step): total {tot:.3f} ms")
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0
    gb = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0
    print(f"  {k:34s} {d['ms']:8.3f} ms  n={d['n']:4d}  {tf:8.1f} TFLOP/s  {gb:8.1f} GB/s")
print("top launches:")
for ms, n, tf, family, tag in prof.top(30):
    print(f"  {ms:7.3f} ms x{n:3d} {tf:7.1f} TF  {family:30s} {tag}")
