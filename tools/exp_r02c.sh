#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "conv_tcgen05_matches_reference or conv_fused_bn_statistics or ce_dice" 2>&1 | tail -5
timeout 120 python tools/kernel_probe.py conv --iters 20
STFB_HALO_PAIR=0 timeout 200 python tools/step_time.py --iters 20
timeout 200 python tools/step_time.py --iters 20
timeout 900 python -m pytest tests/test_dropin_gpu.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -s -k "bench_size or config3" 2>&1 | tail -25
