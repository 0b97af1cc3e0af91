#!/bin/bash
# final-code evidence: ncu metrics of every kernel of one graph replay of the headline step + one full capture of the layer-1 pair kernel
set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
STEP="python tools/step_time.py --iters 1"
timeout 200 $STEP > gpurun_out/plain_step.log 2>&1 || exit 1
tail -1 gpurun_out/plain_step.log
timeout 600 ncu --metrics $M --clock-control none --graph-profiling node -s 1500 -c 560 --csv --log-file gpurun_out/r02_step_metrics_final.csv $STEP > gpurun_out/ncu1.log 2>&1
python tools/step_metrics_summary.py gpurun_out/r02_step_metrics_final.csv > gpurun_out/r02_step_kernels_final.txt 2>&1
head -40 gpurun_out/r02_step_kernels_final.txt | cut -c1-200
timeout 420 ncu --set full --clock-control none --import-source on -k regex:"conv_halo2" -s 2 -c 1 -o /tmp/r02_halo2_c64 -f $STEP > gpurun_out/ncu2.log 2>&1
python tools/ncu_summary.py /tmp/r02_halo2_c64.ncu-rep --stalls 12 > gpurun_out/r02_ncu_conv_halo2_c64.txt 2>&1
rm -f /tmp/r02_halo2_c64.ncu-rep gpurun_out/r02_step_metrics_final.csv
head -30 gpurun_out/r02_ncu_conv_halo2_c64.txt | cut -c1-200
