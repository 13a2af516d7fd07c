#!/bin/bash
bash tools/gpu_quick.sh
bash tools/gpu_prof_kernel.sh "$1" "$2" 256 "${3:-both}"
