#!/bin/bash
# final round-2 bench lines (1 GPU)
set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "rc=$?"; cut -c1-220 gpurun_out/r02_bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout 300 python bench.py --workload infer --steps 20 --warmup 5 > gpurun_out/r02_bench_infer_final.json 2>/dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/r02_bench_infer_final.json
timeout 300 python bench.py --workload volume --steps 10 --warmup 3 > gpurun_out/r02_bench_volume_final.json 2>/dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/r02_bench_volume_final.json
timeout 400 python bench.py --workload train512 --steps 10 --warmup 3 > gpurun_out/r02_bench_train512_final.json 2>/dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/r02_bench_train512_final.json
timeout 400 python bench.py --workload unet --steps 20 --warmup 5 > gpurun_out/r02_bench_unet_final.json 2>/dev/null; echo "rc=$?"; cut -c1-200 gpurun_out/r02_bench_unet_final.json
timeout 300 python tools/step_time.py --fp32 --iters 10 2>&1 | tail -1
STFB_NO_SPLIT_FP32=1 timeout 300 python tools/step_time.py --fp32 --iters 5 2>&1 | tail -1
