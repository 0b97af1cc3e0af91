#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -6
STFB_NO_PACK_OVERLAP=1 timeout 200 python tools/step_time.py --iters 20
timeout 200 python tools/step_time.py --iters 20
