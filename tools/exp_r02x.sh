#!/bin/bash
# Same-box A/B harness: `mkdir _ab_old && git archive <commit> | tar -x -C _ab_old && (cd _ab_old && python -m stf_unet_b200.build)`
# puts an older commit (with its own library) beside the working tree; the runs alternate so both see the same box and clocks.
set -x
one() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'])"; }
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -3
for i in 1 2; do
  (cd _ab_old && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one OLD)
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one NEW
done
timeout 300 python bench.py --workload unet --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('UNET', {k:v['ms_per_step'] for k,v in d['modes'].items()})"
