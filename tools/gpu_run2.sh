#!/bin/bash
mkdir -p gpurun_out
python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden exit $?" >> gpurun_out/golden.log
cp gpurun_out/golden/*.npz tests/golden/
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
for sz in 128 256; do timeout 600 python bench.py --size $sz --steps 5 --warmup 3 > gpurun_out/bench_$sz.json 2> gpurun_out/bench_$sz.err; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err
tail -3 gpurun_out/golden.log; tail -30 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_*.err; cat gpurun_out/bench_512.json
