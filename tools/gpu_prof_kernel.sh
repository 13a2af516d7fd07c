#!/bin/bash
# usage: gpu_prof_kernel.sh <kernel-regex> <tag> [size] [alg]
RE=$1; TAG=$2; SIZE=${3:-256}; ALG=${4:-both}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s 1 -c 1 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --size $SIZE --steps 2 --alg $ALG > gpurun_out/prof_$TAG.log 2>&1
tail -n 3 gpurun_out/prof_$TAG.log
