#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -6
timeout 200 python tools/step_time.py --iters 20
STFB_FUSED_LSTM_BWD=1 timeout 200 python tools/step_time.py --iters 20
STFB_FUSED_LSTM_BWD=1 timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "train or graphed" 2>&1 | tail -4
