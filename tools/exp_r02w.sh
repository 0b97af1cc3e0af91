#!/bin/bash
# fp32 UNet step on bf16x3 operands: family profile, ncu metrics of every launch of one step, one full capture of the pair kernel
set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 300 python tools/unet_fp32_profile.py > gpurun_out/r02_unet_fp32_families.txt 2>&1 || exit 1
tail -50 gpurun_out/r02_unet_fp32_families.txt | cut -c1-200
timeout 300 python bench.py --workload unet --steps 20 --warmup 5 > gpurun_out/r02_bench_unet_split.json 2> gpurun_out/r02_bench_unet_split.err
echo "bench rc=$?"; cut -c1-250 gpurun_out/r02_bench_unet_split.json
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_unet_fp32_metrics.csv python tools/unet_fp32_profile.py --ncu > gpurun_out/ncu_unet1.log 2>&1
python tools/step_metrics_summary.py gpurun_out/r02_unet_fp32_metrics.csv > gpurun_out/r02_unet_fp32_kernels.txt 2>&1
cat gpurun_out/r02_unet_fp32_kernels.txt | cut -c1-220
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_halo2" -s 2 -c 1 -o /tmp/r02_split_halo2 -f python tools/unet_fp32_profile.py --ncu > gpurun_out/ncu_unet2.log 2>&1
python tools/ncu_summary.py /tmp/r02_split_halo2.ncu-rep --stalls 12 > gpurun_out/r02_ncu_conv_halo2_bf16x3.txt 2>&1
rm -f /tmp/r02_split_halo2.ncu-rep
head -60 gpurun_out/r02_ncu_conv_halo2_bf16x3.txt | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_train_b.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_train_b.json
