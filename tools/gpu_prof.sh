#!/bin/bash
# usage: gpu_prof.sh <tag>   -> gpurun_out/launches_<tag>.csv (512^3) and gpurun_out/prof_<tag>.ncu-rep (256^3, full set)
TAG=${1:-base}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:sign_pack|classify_scan|edge_verts|mc_tris|mc_backward|dmc_|quad_' --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --size 512 --steps 2 > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:sign_pack|classify_scan|edge_verts|mc_tris|mc_backward|dmc_|quad_' -s 12 -c 12 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --size 256 --steps 2 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log gpurun_out/prof_$TAG.log; ls -la gpurun_out/*.ncu-rep
