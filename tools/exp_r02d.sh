#!/bin/bash
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err
timeout 600 python bench.py --workload unet --steps 10 --warmup 3 > gpurun_out/r02_bench_unet_d.json 2> gpurun_out/r02_bench_unet_d.err; tail -3 gpurun_out/r02_bench_unet_d.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
