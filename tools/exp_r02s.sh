#!/bin/bash
# split-precision (bf16x3) fp32 path: operator tests, fp32 model parity, UNet fp32 bench
set -x
timeout 400 python -m pytest tests/test_split_gpu.py -q -s --timeout 120 2>&1 | grep -v "^$" | tail -60
timeout 500 python -m pytest tests/test_models_gpu.py -q -s --timeout 200 -k "fp32 or float32 or pk_maps" 2>&1 | grep -i "rel\|passed\|failed\|error\|assert" | cut -c1-250 | tail -40
timeout 300 python bench.py --workload unet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_unet_split.json 2> gpurun_out/r02_bench_unet_split.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_unet_split.err | cut -c1-300; cut -c1-300 gpurun_out/r02_bench_unet_split.json
