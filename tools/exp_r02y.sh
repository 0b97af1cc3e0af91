#!/bin/bash
set -x
timeout 200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "rc=$?"; cut -c1-260 gpurun_out/r02_bench_default.json; tail -2 gpurun_out/r02_bench_default.err | cut -c1-200
