#!/bin/bash
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_worker.py > gpurun_out/r02_dp_worker_n2.log 2>&1
echo "dp_worker rc=$?"; grep -E "DP_|Error|error" gpurun_out/r02_dp_worker_n2.log | tail -20
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n2_overlap.json 2> gpurun_out/r02_bench_n2_overlap.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n2_overlap.err
