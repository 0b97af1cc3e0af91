#!/bin/bash
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_tofts_gpu.py -x -q -s -m gpu 2>&1 | tail -25
STFB_DP_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n2_after.json 2> gpurun_out/r02_bench_n2_after.err; tail -2 gpurun_out/r02_bench_n2_after.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n2_overlap.json 2> gpurun_out/r02_bench_n2_overlap.err; tail -2 gpurun_out/r02_bench_n2_overlap.err
python - <<'PY'
import json
for f in ("after","overlap"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_n2_{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
