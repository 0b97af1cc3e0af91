#!/bin/bash
# programmatic dependent launch by kernel family (STFB_PDL bit mask: 1 = convolutions, 2 = BatchNorm kernels): A/B of the headline step and of inference
set -x
one() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'])"; }
for m in 0 1 2 0 1; do
  STFB_PDL=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one PDL$m
done
for m in 0 3 1; do
  STFB_PDL=$m timeout 200 python bench.py --workload infer --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | one INFER_PDL$m
done
