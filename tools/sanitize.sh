#!/bin/bash
# compute-sanitizer passes SURVEY.md §5 asks for (run on the GPU box through gpurun; logs in gpurun_out/, summaries are
# copied to profiles/): memcheck, racecheck (shared-memory hazards: hash/scan/sort/conv kernels) and initcheck
# (reads of never-written global memory) over the bit-exact map test, layer-level convolutions and one fused forward
# with the shape sort forced on.
set -x
SEL="tests/test_gpu_parity.py::test_voxelize_and_maps_bit_exact tests/test_gpu_parity.py::test_edge_inputs tests/test_gpu_conv.py::test_sparse_conv_fp16_rows tests/test_gpu_conv.py::test_two_segment_input_rows tests/test_gpu_conv.py::test_split_precision_options_of_the_fp16_path tests/test_gpu_conv.py::test_pattern_sorted_processing_order_does_not_change_results tests/test_rosio.py"
for tool in memcheck racecheck initcheck; do
  timeout 540 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log \
    python -m pytest $SEL -m gpu -q -x -k "not hdl-32 or sorted" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  tail -3 gpurun_out/sanitizer_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -2
done
