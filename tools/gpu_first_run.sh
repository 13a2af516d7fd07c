#!/bin/bash
# First GPU session: golden generation with the reference, parity tests, smoke.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden exit $?" >> gpurun_out/golden.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_oracle_golden.py 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/golden.log; cat gpurun_out/smoke.log | tail -5; tail -30 gpurun_out/pytest_gpu.log
