#!/bin/bash
# round-2 experiment A: pixel-tile pairing (MT=2) per layer width, in isolation and inside the graph step
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; lscpu | grep "Model name"
python tools/kernel_probe.py conv --iters 20
STFB_HALO_MT=2 python tools/kernel_probe.py conv --iters 20
python tools/kernel_probe.py wgrad --iters 20
python tools/step_time.py --iters 20
STFB_HALO_MT=2 STFB_HALO_MT_BN=256 python tools/step_time.py --iters 20
STFB_HALO_MT=2 STFB_HALO_MT_BN=128,256 python tools/step_time.py --iters 20
STFB_HALO_MT=2 python tools/step_time.py --iters 20
