#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_step.py --size 40 --steps 1 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  tail -n 4 gpurun_out/sanitize_$tool.log
done
# odd sizes / fp64 through the test-suite's smallest cases under memcheck
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "ragged_5x9x70 or tiny_2x2x2 or thin_1x7x33 or ragged_31x2x30" > gpurun_out/sanitize_pytest.log 2>&1; echo "pytest-memcheck exit $?" >> gpurun_out/sanitize_pytest.log; tail -n 4 gpurun_out/sanitize_pytest.log
