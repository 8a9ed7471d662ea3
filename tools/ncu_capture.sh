#!/bin/bash
# Profiling passes committed under profiles/ (run on the GPU box through gpurun; outputs in gpurun_out/):
#  1. launch list of a bench run (per-launch gpu__time_duration, the "share of the step" evidence)
#  2. metrics pass over every kernel of ONE forward (time, DRAM bytes, unit throughputs, tensor pipe)
#  3. `--set full` capture of the dominant kernels of the forward, and of one wide layer of the width sweep
R=${1:-r2}
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
# the two timed device-resident steps of a single-lane bench run (NL launches per forward; 3 warm-up forwards before)
NL=${NL:-56}
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * NL)) -c $((2 * NL)) --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --no-alt > gpurun_out/${R}_launches_bench.log 2>&1
ncu --metrics $M --clock-control none -s $NL -c $NL --csv --log-file gpurun_out/${R}_forward_metrics.csv python tools/profile_forward.py 2 > gpurun_out/${R}_forward_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_kernel_map_blk3|k_conv_umma6|k_onesweep_pass|k_tile_masks_perm" -s 30 -c 30 -o /tmp/top_${R} python tools/profile_forward.py 2 > gpurun_out/${R}_top_full.log 2>&1
ncu -i /tmp/top_${R}.ncu-rep --page raw --csv > gpurun_out/${R}_top_full.csv 2>/dev/null
# one wide layer (PLANES x4, block5.conv2: 256 -> 256 on level 3): the 8th 81-offset layer of the sweep, 6 launches each, after 8 forwards
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv_umma6<\(int\)256" -s 12 -c 2 -o /tmp/wide_${R} python tools/width_sweep.py --widths 4 --steps 1 > gpurun_out/${R}_wide_full.log 2>&1
ncu -i /tmp/wide_${R}.ncu-rep --page raw --csv > gpurun_out/${R}_wide_full.csv 2>/dev/null
cp /tmp/wide_${R}.ncu-rep gpurun_out/ 2>/dev/null
python tools/ncu_top_md.py gpurun_out/${R}_top_full.csv "Kernel-map, sort, slice and convolution kernels of ONE forward, full ncu capture (${R})" > gpurun_out/${R}_top_kernels_ncu_full.md
python tools/ncu_top_md.py gpurun_out/${R}_wide_full.csv "One wide layer of the width sweep (PLANES x4, 256 output channels), full ncu capture (${R})" > gpurun_out/${R}_wide_layer_ncu_full.md
python tools/make_profile_tables.py gpurun_out/${R}_forward_metrics.csv gpurun_out/${R}_forward_kernels_ncu.md gpurun_out/traffic.json > /dev/null
ls -la gpurun_out | tail -12
