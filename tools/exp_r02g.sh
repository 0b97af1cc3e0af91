#!/bin/bash
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
STFB_NO_FWD_SPLIT=1 timeout 200 python tools/step_time.py --iters 20
timeout 200 python tools/step_time.py --iters 20
timeout 300 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r02_bench_infer_g.json 2> gpurun_out/r02_bench_infer_g.err; tail -2 gpurun_out/r02_bench_infer_g.err; cat gpurun_out/r02_bench_infer_g.json | cut -c1-400
