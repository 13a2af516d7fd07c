#!/bin/bash
# compute-sanitizer over the dense path (random field), the sparse paths (listed tiles, zero fill + touched-block
# backward: sphere) and the smallest / odd-shaped parity cases
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_step.py --size 40 --steps 1 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool dense exit $?" >> gpurun_out/sanitize_$tool.log
  for alg in mc dmc; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/kernel_breakdown.py sphere 72 $alg >> gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool sphere $alg exit $?" >> gpurun_out/sanitize_$tool.log
  done
  grep -E "exit|SUMMARY" gpurun_out/sanitize_$tool.log
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_tile_variants_gpu.py -q -x -k "ragged_5x9x70 or tiny_2x2x2 or thin_1x7x33 or ragged_31x2x30 or tiles_agree" > gpurun_out/sanitize_pytest.log 2>&1; echo "pytest-memcheck exit $?" >> gpurun_out/sanitize_pytest.log; tail -n 4 gpurun_out/sanitize_pytest.log
# the sparse backward (zero fill + touched-block list + persistent grid) forced on the dense input
for tool in memcheck racecheck initcheck; do
  DISO_BWD_SPARSE=1 timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_step.py --size 40 --steps 1 > gpurun_out/sanitize_sparsebwd_$tool.log 2>&1
  echo "$tool forced sparse backward exit $?" >> gpurun_out/sanitize_sparsebwd_$tool.log
  grep -E "exit|SUMMARY" gpurun_out/sanitize_sparsebwd_$tool.log
done
