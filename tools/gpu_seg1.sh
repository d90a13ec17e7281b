#!/bin/bash
# Dev script (GPU box): one ncu full capture + SASS segment breakdown.  usage: gpu_seg1.sh <tag> <bench args...>
tag=$1; shift
mkdir -p gpurun_out/$tag
timeout 600 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega python bench.py --steps 1 --warmup 3 --no-baselines "$@" > gpurun_out/$tag/ncu.log 2>&1
python tools/ncu_segments.py gpurun_out/$tag/mega.ncu-rep 0.002 > gpurun_out/$tag/segments.txt 2>&1
python tools/ncu_summary.py gpurun_out/$tag/mega.ncu-rep > gpurun_out/$tag/summary.txt 2>&1
