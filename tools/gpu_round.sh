#!/bin/bash
# Dev script (GPU box): golden vectors, GPU tests, smoke, bench, ncu launch list + full capture.
set -x
mkdir -p gpurun_out/golden gpurun_out/prof
python tests/golden/make_golden_gpu.py gpurun_out/golden 2>&1 | tail -2
cp gpurun_out/golden/gpu_golden.npz tests/golden/gpu_golden.npz
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 300 python bench.py --steps 3 --warmup 3 --workload c3 --no-baselines > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c3.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload c4 --no-baselines > gpurun_out/bench_c4.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/prof/ncu_bench_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -o gpurun_out/prof/mega_c2 python bench.py --steps 2 --warmup 3 --no-baselines > gpurun_out/prof/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -o gpurun_out/prof/mega_c3 python bench.py --steps 2 --warmup 3 --workload c3 --spp 16 --no-baselines > gpurun_out/prof/ncu_full_c3.log 2>&1
ls -la gpurun_out/prof
