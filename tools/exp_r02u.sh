#!/bin/bash
# A/B: occupancy-aware BN choice and fp32 side streams on the UNet fp32 / bf16 step; then the fp32 model tests
set -x
for cfg in "STFB_BN_OCC=0 STFB_NO_FP32_SIDE_STREAMS=1" "STFB_BN_OCC=1 STFB_NO_FP32_SIDE_STREAMS=1" "STFB_BN_OCC=1 STFB_NO_FP32_SIDE_STREAMS=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --workload unet --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:(v['ms_per_step'],v['loss']) for k,v in d['modes'].items()})"
done
timeout 600 python -m pytest tests/test_split_gpu.py tests/test_models_gpu.py -q --timeout 200 -k "split or fp32 or float32 or pk_maps or unet or graphed" 2>&1 | tail -8
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | cut -c1-200
