#!/bin/bash
# Round-end measurement suite on ONE B200 (through gpurun; everything lands in gpurun_out/ and is copied to profiles/ by hand):
# GPU tests, the default bench line (with cpu_baseline and the TF32 alt backend), the reference arm, the streamed
# configuration, the precision probe and the three compute-sanitizer passes.
R=${1:-r2}
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${R}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
timeout 600 python bench.py --config 3 > gpurun_out/${R}_stream_config3.json 2> gpurun_out/${R}_stream_config3.err
timeout 600 python tools/precision_probe.py > gpurun_out/${R}_precision_probe.txt 2>&1
bash tools/sanitize.sh > gpurun_out/${R}_sanitize.log 2>&1
tail -c 600 gpurun_out/${R}_bench_n1.json; cat gpurun_out/${R}_pytest_gpu.txt; tail -12 gpurun_out/${R}_sanitize.log
