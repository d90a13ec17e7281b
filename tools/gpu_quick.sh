#!/bin/bash
# Dev script (GPU box): selected GPU tests, then quick device-timed bench lines.  usage: gpu_quick.sh <tag> "<pytest -k expr or ''>" spec...
#   spec = name:"bench args"
tag=$1; kexpr=$2; shift 2
mkdir -p gpurun_out/$tag
if [ -n "$kexpr" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$kexpr" 2>&1 | tail -8 | tee gpurun_out/$tag/pytest.txt
fi
for spec in "$@"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout 300 python bench.py --steps 3 --warmup 3 --no-baselines $args > gpurun_out/$tag/$name.json 2> gpurun_out/$tag/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag/$name.json"))
    print("$name", round(d["value"],1), "Mpaths/s", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],4), d["kernel"]["form"], "verified", d["verified"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/$tag/$name.err").read()[-800:])
PY
done
