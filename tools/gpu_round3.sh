#!/bin/bash
# Dev script (GPU box): GPU tests, smoke, the default bench line (both arms), the ncu launch list and one full capture of the config-2 kernel.
#   usage: gpu_round3.sh <tag>
tag=${1:-r02p}
mkdir -p gpurun_out/$tag
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/$tag/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/$tag/bench_c2.json 2> gpurun_out/$tag/bench_c2.err; tail -c 1500 gpurun_out/$tag/bench_c2.json; tail -8 gpurun_out/$tag/bench_c2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$tag/bench_reference_arm.json 2>> gpurun_out/$tag/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$tag/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/$tag/ncu_bench_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c2 python bench.py --steps 1 --warmup 3 --spp 256 --no-baselines > gpurun_out/$tag/ncu_full_c2.log 2>&1
ls -la gpurun_out/$tag
