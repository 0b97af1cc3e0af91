#!/bin/bash
set -x
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_train_c.json 2>/dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/r02_bench_train_c.json
timeout 300 python bench.py --workload unet --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('UNET', {k:v['ms_per_step'] for k,v in d['modes'].items()})"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
