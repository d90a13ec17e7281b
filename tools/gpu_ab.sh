#!/bin/bash
# Dev script (GPU box): A/B the product library against experimental builds.  usage: gpu_ab.sh "<bench args>" lib1.so lib2.so ...
args=$1; shift
for l in "$@"; do
  ATX_LIB=$l timeout 200 python bench.py --steps 3 --warmup 3 --no-baselines $args | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
done
