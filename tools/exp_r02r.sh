#!/bin/bash
set -x
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n8.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]); print('n8', d['value'], d['ms_per_step'], d['e2e']['value'])"
