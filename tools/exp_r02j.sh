#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm or ce_dice or conv_simt or head or errors" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_models_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -4
STFB_NO_LSTM_SEQ=1 timeout 200 python tools/step_time.py --iters 20
timeout 200 python tools/step_time.py --iters 20
timeout 300 python bench.py --workload infer --steps 10 --warmup 3 2>/dev/null | cut -c1-230
