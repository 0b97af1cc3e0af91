#!/bin/bash
set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
timeout 300 python bench.py --workload infer --steps 20 --warmup 5 > gpurun_out/r02_bench_infer_final.json 2>/dev/null
timeout 300 python bench.py --workload volume --steps 5 --warmup 3 > gpurun_out/r02_bench_volume_final.json 2>/dev/null
timeout 600 python bench.py --workload train512 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_train512_final.json 2>/dev/null
timeout 600 python bench.py --workload unet --steps 10 --warmup 3 > gpurun_out/r02_bench_unet_final.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for f in final infer_final volume_final train512_final unet_final reference_arm; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d.get('value'), d.get('unit'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('model_frac_of_bf16_peak'))"; done
