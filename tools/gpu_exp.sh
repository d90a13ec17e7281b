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
for spec in "$@"; do
  name=${spec%%:*}; args=${spec#*:}
  b $name $args
done
