#!/bin/bash
set -x
timeout 400 python -m pytest tests/test_dp_gpu.py -x -q -s -m gpu 2>&1 | grep -E "DP_CHECK|DP_OK|DP_FAIL|passed|failed" | tail -8
