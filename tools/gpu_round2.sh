#!/bin/bash
# Dev script (GPU box): GPU tests, smoke, config-2 golden digests, the default bench line, the reference arm, ncu captures.
#   usage: gpu_round2.sh <tag> [skip-tests]
tag=${1:-r02a}
mkdir -p gpurun_out/$tag gpurun_out/golden
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/$tag/pytest_gpu.txt
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tests/golden/make_golden_c2.py gpurun_out/golden 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/$tag/bench_c2.json 2> gpurun_out/$tag/bench_c2.err; tail -c 6000 gpurun_out/$tag/bench_c2.json; tail -8 gpurun_out/$tag/bench_c2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$tag/bench_reference_arm.json 2>> gpurun_out/$tag/bench_c2.err; cat gpurun_out/$tag/bench_reference_arm.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c3 python bench.py --steps 1 --warmup 3 --workload c3 --spp 32 --no-baselines > gpurun_out/$tag/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/$tag/mega_c4 python bench.py --steps 1 --warmup 3 --workload c4 --spp 8 --no-baselines > gpurun_out/$tag/ncu_full_c4.log 2>&1
ls -la gpurun_out/$tag
