"""CPU (gloo, world_size 2) coverage of the data-parallel plumbing: flat-gradient bucketing / averaging, the initial
state broadcast, and the shard arithmetic of bench.py.  The kernels themselves need a GPU; the exchange logic does not."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stf_unet_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return dict(ret)


def _allreduce_case(rank, world):
    torch.manual_seed(rank)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    parallel.allreduce_mean_(flat, bucket_elems=128)           # 8 buckets, last one ragged
    return flat.tolist()


def test_bucketed_allreduce_is_the_mean_over_ranks():
    out = _run(_allreduce_case)
    expect = (torch.arange(1000, dtype=torch.float32) * 1.5).tolist()
    assert out[0] == expect and out[1] == expect


def _broadcast_case(rank, world):
    torch.manual_seed(100 + rank)                               # different init per rank
    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
    m[1].running_mean.fill_(float(rank))
    dp = parallel.DataParallel(m)
    flat = torch.full((10,), float(rank + 1))
    dp._on_grads(flat)                                          # what ModelFunction.backward calls after the tape
    return m[0].weight.flatten().tolist(), m[1].running_mean.tolist(), flat.tolist(), hasattr(m, "_grad_ready_hook")


def test_dataparallel_broadcasts_state_and_averages_grads():
    out = _run(_broadcast_case)
    assert out[0][0] == out[1][0]                               # rank 0's weights everywhere
    assert out[1][1] == [0.0] * 4                               # buffers too
    assert out[0][2] == [1.5] * 10 and out[1][2] == [1.5] * 10  # mean of (1, 2)
    assert out[0][3] and out[1][3]


def test_bucket_ranges_cover_exactly():
    assert parallel.bucket_ranges(0, 16) == []
    r = parallel.bucket_ranges(100, 32)
    assert r == [(0, 32), (32, 64), (64, 96), (96, 100)]
    assert parallel.bucket_ranges(5, 0) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]


def test_bench_shards_are_disjoint_per_rank():
    import bench
    x0, t0 = bench.make_batch(0, batch=2, T=2, hw=32)
    x1, t1 = bench.make_batch(1, batch=2, T=2, hw=32)
    assert x0.shape == (2, 2, 1, 32, 32) and t0.shape == (2, 16, 16)
    assert not torch.equal(x0, x1)
    cfg = bench.workload_config(8)
    assert cfg["global_batch"] == 128 and cfg["parallelism"] == "dp8"


@pytest.mark.parametrize("impl", ["reference"])
def test_bench_reference_arm_prints_contract_line(impl, capsys, monkeypatch):
    import json
    import bench
    monkeypatch.setattr(bench, "HW", 32)
    monkeypatch.setattr(bench, "T_PHASES", 2)

    class A:
        steps, warmup, gpus, workload = 1, 0, 1, "train"
    bench.run_reference(A, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "slices/s" and line["value"] > 0
    # the unmodified reference when it is installed (baseline/_ref or the build container's /root/reference), else the port
    want = "reference" if bench.reference_root() is not None else "port"
    assert line["cpu_baseline"]["kind"] == want and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "CUDA graph" not in line["config"]["launch"]         # the arm's own launch description, not the GPU arm's
    assert line["config"]["workload"] == bench.workload_config(1)["workload"]
    bench.run_reference(A, 1)                                   # non-zero ranks print nothing
    assert capsys.readouterr().out == ""


def test_bench_reference_arm_falls_back_to_the_oracle_port(capsys, monkeypatch):
    """Without baseline/_ref (a checkout where tools/install_ref.sh never ran) the arm still answers: kind = "port"."""
    import json
    import bench
    monkeypatch.setattr(bench, "HW", 32)
    monkeypatch.setattr(bench, "T_PHASES", 2)
    monkeypatch.setattr(bench, "_REF", False)

    class A:
        steps, warmup, gpus, workload = 1, 0, 1, "train"
    bench.run_reference(A, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
