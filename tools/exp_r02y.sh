#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu --timeout 120 -k "programmatic or conv_fused_bn" 2>&1 | tail -5
