#!/bin/bash
# Dev script (GPU box): parity tests, then quick bench sweeps of the megakernel forms.
mkdir -p gpurun_out/exp
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
b() { name=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-baselines "$@" > gpurun_out/exp/$name.json 2> gpurun_out/exp/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp/$name.json"))
    print("$name", round(d["value"],1), "Mpaths/s", round(d["ms_per_step"],2), "ms  frac", round(d["roofline"]["frac"],4), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/exp/$name.err").read()[-600:])
PY
}
b c2_ww1 --workload c2 --mega-kind 1 --trace-rounds 1
b c2_ww2 --workload c2 --mega-kind 1 --trace-rounds 2
b c2_ww3 --workload c2 --mega-kind 1 --trace-rounds 3
b c2_ww4 --workload c2 --mega-kind 1 --trace-rounds 4
b c2_pair --workload c2 --mega-kind 2
b c3_pair --workload c3 --spp 64
b c3_ww --workload c3 --spp 64 --mega-kind 1
b c4_pair --workload c4 --spp 16
b c4_chunk --workload c4 --spp 16 --chunk 2048
