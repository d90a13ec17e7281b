#!/bin/bash
# Dev script (GPU box): GPU tests, smoke, the three bench lines, ncu launch list + full captures.  usage: gpu_round.sh <tag>
tag=${1:-r01}
set -x
mkdir -p gpurun_out/$tag
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/$tag/bench_c2.json 2> gpurun_out/$tag/bench_c2.err; tail -c 4000 gpurun_out/$tag/bench_c2.json; tail -5 gpurun_out/$tag/bench_c2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$tag/bench_reference_arm.json 2>> gpurun_out/$tag/bench_c2.err; cat gpurun_out/$tag/bench_reference_arm.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload c3 --no-baselines > gpurun_out/$tag/bench_c3.json 2>> gpurun_out/$tag/bench_c2.err; cat gpurun_out/$tag/bench_c3.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload c4 --no-baselines > gpurun_out/$tag/bench_c4.json 2>> gpurun_out/$tag/bench_c2.err; cat gpurun_out/$tag/bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$tag/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/$tag/ncu_bench_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c2 python bench.py --steps 1 --warmup 3 --spp 256 --no-baselines > gpurun_out/$tag/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c3 python bench.py --steps 1 --warmup 3 --workload c3 --spp 32 --no-baselines > gpurun_out/$tag/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c4 python bench.py --steps 1 --warmup 3 --workload c4 --spp 8 --no-baselines > gpurun_out/$tag/ncu_full_c4.log 2>&1
ls -la gpurun_out/$tag
