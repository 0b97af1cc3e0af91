#!/bin/bash
set -x
timeout 500 python -m pytest tests/test_dp_gpu.py -x -q -s -m gpu 2>&1 | grep -E "DP_|passed|failed|Error" | tail -14
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n2_final.err
STFB_DP_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n2_after_final.json 2> gpurun_out/r02_bench_n2_after_final.err
echo "bench rc=$?"
for f in n2_final n2_after_final; do python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
