R=r2; NL=56
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * NL)) -c $((2 * NL)) --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --no-alt > gpurun_out/${R}_launches_bench.log 2>&1
timeout 120 ncu --metrics $M --clock-control none -s $NL -c $NL --csv --log-file gpurun_out/${R}_forward_metrics.csv python tools/profile_forward.py 2 > gpurun_out/${R}_forward_metrics.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_kernel_map_blk3|k_conv_umma6|k_onesweep_pass|k_tile_masks_perm" -s 30 -c 30 -o /tmp/top_${R} python tools/profile_forward.py 2 > gpurun_out/${R}_top_full.log 2>&1
ncu -i /tmp/top_${R}.ncu-rep --page raw --csv > gpurun_out/${R}_top_full.csv 2>/dev/null
python tools/ncu_top_md.py gpurun_out/${R}_top_full.csv "Kernel-map, sort, slice and convolution kernels of ONE forward, full ncu capture (${R})" > gpurun_out/${R}_top_kernels_ncu_full.md
python tools/make_profile_tables.py gpurun_out/${R}_forward_metrics.csv gpurun_out/${R}_forward_kernels_ncu.md gpurun_out/traffic.json > /dev/null
wc -l gpurun_out/${R}_top_kernels_ncu_full.md gpurun_out/${R}_forward_kernels_ncu.md
