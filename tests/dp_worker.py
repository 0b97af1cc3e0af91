"""Worker of tests/test_dp_gpu.py (one process per GPU under torch.distributed.run, NCCL).

SURVEY.md section 8(e): "N-GPU gradients == mean of the N single-GPU gradients on the same shards".  Every rank computes
the gradient of EVERY shard on its own GPU without data parallelism (the expectation), then the data-parallel gradient of its
own shard three ways -- exchange after the backward pass, per-segment exchange overlapped with it, and the overlapped
exchange captured inside the CUDA graph of the step -- and compares the flat gradient buffers."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stf_unet_b200 as S  # noqa: E402
from stf_unet_b200 import parallel  # noqa: E402
from stf_unet_b200.graph import GraphedStep  # noqa: E402
from stf_unet_b200.synthetic import synthetic_dce_batch  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, T, HW = 4, 4, 128
    shards = [tuple(t.to(dev) for t in synthetic_dce_batch(B, T, HW, HW, seed=500 + r)) for r in range(world)]
    torch.manual_seed(0)
    base = S.STFLSTMUNet(1, 2, T).to(dev)
    sd = {k: v.detach().clone() for k, v in base.state_dict().items()}
    ok = True
    # bf16: the BatchNorm statistics fused into the conv epilogues (and the one-launch BatchNorm backward) meet through fp32
    # red.adds whose order differs from run to run; on these cold random weights train-mode BatchNorm amplifies that noise to
    # ~20 % of a gradient (measured: the SAME shard computed twice differs by 2.3e-1), which would hide any exchange bug.  The
    # data-parallel check therefore runs the deterministic routes (what test_graphed_step_matches_eager[False] also pins).
    from stf_unet_b200 import engine, ops
    engine.USE_FUSED_BN_STATS = False
    ops.USE_FUSED_BN_BWD = False
    for tag, dt, tol in (("fp32", None, 2e-5), ("bf16", torch.bfloat16, 2e-3)):
        def fresh():
            m = S.STFLSTMUNet(1, 2, T).to(dev)
            m.load_state_dict(sd)
            return m.train()

        def step(net, x, t):
            for p in net.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt is not None):
                loss = S.criterion(net(x), t)
            loss.backward()
            return loss

        # expectation: mean over shards of the single-GPU gradients (fresh running stats per shard, like separate replicas)
        expect = None
        for x, t in shards:
            m = fresh()
            step(m, x, t)
            g = m._last_flat_grad.detach().clone()
            expect = g if expect is None else expect + g
        expect /= world
        print(f"DP_PROGRESS rank={rank} {tag}: expectation done", flush=True)
        x, t = shards[rank]
        results = {}
        for mode in ("after", "overlap"):
            m = fresh()
            net = parallel.DataParallel(m, overlap=(mode == "overlap"))
            assert (m.__dict__.get("_grad_segment_hook") is not None) == (mode == "overlap")
            step(net, x, t)
            torch.cuda.synchronize()
            results[mode] = rel(m._last_flat_grad, expect)
            print(f"DP_PROGRESS rank={rank} {tag}: {mode} {results[mode]:.3e}", flush=True)
        m = fresh()
        net = parallel.DataParallel(m, overlap=True)
        g = GraphedStep(m, S.criterion, x, t, autocast_dtype=dt)
        assert g._in_graph_comm, "the collectives were not captured inside the graph"
        for _ in range(2):
            g(x, t)
        torch.cuda.synchronize()
        results["graph"] = rel(g.flat_grad, expect)   # replays move the running stats, not the weights: same gradient every time
        # every rank must hold the SAME averaged gradient
        mine = g.flat_grad.detach().clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        results["rank_spread"] = rel(mine, other) if rank != 0 else 0.0
        line = " ".join(f"{k}={v:.3e}" for k, v in results.items())
        good = all(v < tol for k, v in results.items() if k != "rank_spread") and results["rank_spread"] == 0.0
        print(f"DP_CHECK rank={rank} world={world} {tag}: {line} tol={tol:.0e} {'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
        # a communicator can only be torn down after every CUDA graph that captured its collectives is gone
        del g, net, m
        import gc
        gc.collect()
        torch.cuda.synchronize()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = int(flag) == 1
    if rank == 0:
        print("DP_OK" if good else "DP_FAIL", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
