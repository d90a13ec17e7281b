#!/bin/bash
# Dev script (GPU box): one ncu --set full capture of the megakernel on a workload.  usage: gpu_prof.sh <name> <bench args...>
name=$1; shift
mkdir -p gpurun_out/prof
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -o gpurun_out/prof/$name -f python bench.py --steps 1 --warmup 3 --no-baselines "$@" > gpurun_out/prof/$name.log 2>&1
tail -3 gpurun_out/prof/$name.log
