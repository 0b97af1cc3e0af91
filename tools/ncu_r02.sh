#!/bin/bash
# round-2 ncu captures (one gpurun call; every ncu command follows a plain run of the same command that exited 0).
# Reports are summarised ON the box (tools/ncu_summary.py) and deleted: gpurun_out/ may only bring back 64 MiB.
set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
STEP="python tools/step_time.py --iters 1"
mkdir -p gpurun_out/ncu
timeout 200 $STEP > gpurun_out/ncu/plain_step.log 2>&1 || exit 1
# (1) every node of one graph replay with the roofline metrics (cold, serialised): skip the three eager warm-up steps
timeout 600 ncu --metrics $M --clock-control none --graph-profiling node -s 1500 -c 560 --csv --log-file gpurun_out/ncu/r02_step_metrics.csv $STEP > gpurun_out/ncu/ncu1.log 2>&1
cap() {   # name, kernel regex, skip, count [, extra args]
  timeout 420 ncu --set full --clock-control none --import-source on --graph-profiling node $5 -k regex:"$2" -s $3 -c $4 -o /tmp/$1 -f $STEP > gpurun_out/ncu/$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep --stalls 12 > gpurun_out/ncu/$1.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap r02_ncu_conv_halo2_l1 conv_halo2 186 1
cap r02_ncu_conv_halo2_l2 conv_halo2 192 1
cap r02_ncu_conv_halo2_l3 conv_halo2 200 1
cap r02_ncu_wgrad_instep "wgrad_(halo|tc)_kernel" 190 4
cap r02_ncu_lstm_seq "lstm_seq64" 3 1
cap r02_ncu_lstm_step "conv_tc_kernel<256, 3" 80 2 "--kernel-name-base demangled"
cap r02_ncu_bn_elementwise "bn_bwd_fused|bn_apply_kernel|lstm_cell_bwd|bn_relu_maxpool" 400 4
cap r02_ncu_misc "colreduce|ce_dice_fwd|ce_dice_bwd|adamw_flat|wgrad_scatter_batched|pointwise_narrow" 60 8
AUX="python tools/aux_probe.py"
timeout 200 $AUX > gpurun_out/ncu/aux_probe.log 2>&1 && timeout 420 ncu --set full --clock-control none -k regex:"augment|tofts|eval_metrics" -s 3 -c 4 -o /tmp/r02_aux -f $AUX > gpurun_out/ncu/r02_aux.log 2>&1
python tools/ncu_summary.py /tmp/r02_aux.ncu-rep > gpurun_out/ncu/r02_ncu_aux.txt 2>&1; rm -f /tmp/r02_aux.ncu-rep
tail -2 gpurun_out/ncu/*.log | cut -c1-200
du -sh gpurun_out
