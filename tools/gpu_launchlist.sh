#!/bin/bash
TAG=${1:-r1_final}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/launches_all_$TAG.log 2>&1
tail -n 2 gpurun_out/launches_all_$TAG.log | cut -c1-200
