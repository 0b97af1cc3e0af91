"""Data parallelism on hardware (needs >= 2 GPUs; the driver's 1-GPU `pytest -m gpu` skips it -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu -q`; the last run's output is committed under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_gradients_equal_mean_of_shard_gradients():
    """N-GPU averaged gradients == mean of the per-shard single-GPU gradients, for the exchange after
    the backward pass, the per-segment exchange overlapped with it, and the overlapped exchange captured in the CUDA graph."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("DP_CHECK") == 2 * world
