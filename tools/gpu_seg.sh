#!/bin/bash
# Dev script (GPU box): ncu full captures of c2 and c3 + SASS segment breakdown (tools/ncu_segments.py).  usage: gpu_seg.sh <tag>
tag=${1:-seg}
mkdir -p gpurun_out/$tag
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c2 python bench.py --steps 1 --warmup 3 --spp 256 --no-baselines > gpurun_out/$tag/ncu_full_c2.log 2>&1
python tools/ncu_segments.py gpurun_out/$tag/mega_c2.ncu-rep 0.002 > gpurun_out/$tag/segments_c2.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c3 python bench.py --steps 1 --warmup 3 --workload c3 --spp 32 --no-baselines > gpurun_out/$tag/ncu_full_c3.log 2>&1
python tools/ncu_segments.py gpurun_out/$tag/mega_c3.ncu-rep 0.002 > gpurun_out/$tag/segments_c3.txt 2>&1
ls -la gpurun_out/$tag
