#!/bin/bash
# Profiling passes committed under profiles/ (run on the GPU box through gpurun; outputs in gpurun_out/):
#  1. launch list of a bench run (per-launch gpu__time_duration, the "share of the step" evidence)
#  2. metrics pass over every heavy kernel of ONE forward (time, DRAM bytes, unit throughputs, tensor pipe)
#  3. `--set full` capture of the dominant kernels (kernel map L0, conv0, three convolutions), raw page as CSV
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --metrics $M --clock-control none -s 137 -c 137 --csv --log-file gpurun_out/forward_metrics_r1.csv python tools/profile_forward.py 2 > gpurun_out/forward_metrics.log 2>&1
ncu --set full --clock-control none -k regex:"k_kernel_map_blk3|k_conv0_const|k_conv_umma6" -s 30 -c 10 -o /tmp/top_r1 python tools/profile_forward.py 2 > gpurun_out/top_full.log 2>&1
ncu -i /tmp/top_r1.ncu-rep --page raw --csv > gpurun_out/top_full_r1.csv 2>/dev/null
ls -la gpurun_out | tail -8
