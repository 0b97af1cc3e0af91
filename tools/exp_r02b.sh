#!/bin/bash
# round-2 experiment B: CTA-pair (cta_group::2) halo conv kernel: parity, isolated timing, in-step timing
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "conv_tcgen05_matches_reference or conv_fused_bn_statistics" 2>&1 | tail -15
STFB_HALO_PAIR=0 timeout 120 python tools/kernel_probe.py conv --iters 20
timeout 120 python tools/kernel_probe.py conv --iters 20
STFB_HALO_PAIR=0 timeout 200 python tools/step_time.py --iters 20
timeout 200 python tools/step_time.py --iters 20
