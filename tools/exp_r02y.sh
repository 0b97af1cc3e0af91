#!/bin/bash
set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -5
timeout 300 python bench.py --workload unet --steps 20 --warmup 5 > gpurun_out/r02_bench_unet_split.json 2> gpurun_out/r02_bench_unet_split.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_unet_split.json').read().strip().splitlines()[-1]); print(d['modes'], d['gpu_launches'])"
timeout 300 python bench.py --workload infer --steps 20 --warmup 5 > gpurun_out/r02_bench_infer_b.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_infer_b.json
