#!/bin/bash
mkdir -p gpurun_out
python tools/bench_configs.py > gpurun_out/configs.md 2> gpurun_out/configs.err; tail -n 3 gpurun_out/configs.err; cat gpurun_out/configs.md
ncu --set full --clock-control none --import-source on -k "regex:mc_backward" -s 1 -c 1 -f -o gpurun_out/prof_bwd512 python tools/profile_step.py --size 512 --steps 2 --alg mc > gpurun_out/prof_bwd512.log 2>&1; tail -n 2 gpurun_out/prof_bwd512.log
