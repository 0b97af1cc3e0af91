#!/bin/bash
# same-box A/B of the headline step: the commit before the split-precision work (_ab_old), its Python over the new library (_ab_mix), the working tree
set -x
one() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'])"; }
for i in 1 2; do
  (cd _ab_old && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one OLD)
  (cd _ab_mix && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one OLDPY_NEWLIB)
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | one NEW
done
