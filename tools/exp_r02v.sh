#!/bin/bash
set -x
for cfg in "STFB_BN_OCC=0" "STFB_BN_OCC=1" "STFB_BN_OCC=0" "STFB_BN_OCC=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
for cfg in "STFB_BN_OCC=0" "STFB_BN_OCC=1"; do
  env $cfg timeout 300 python bench.py --workload infer --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -6
