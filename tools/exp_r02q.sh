#!/bin/bash
# 2 GPUs: headline bench with the in-graph overlapped exchange (final code), twice
set -x
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_n2_final.json; grep "loss_check" gpurun_out/r02_bench_n2_final.err | cut -c1-400
done
