#!/bin/bash
# Dev script (GPU box): A/B the product library against experimental builds.  usage: gpu_ab.sh "<bench args>" lib1.so lib2.so ...
args=$1; shift
for l in "$@"; do
  ATX_LIB=$l timeout 200 python bench.py --steps 3 --warmup 3 --no-baselines $args 2>/tmp/ab.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$l', '$args', round(d['value'],1), round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],4), d['kernel']['form'], 'verified', d['verified'])
except Exception as e:
    print('$l failed', e); print(open('/tmp/ab.err').read()[-600:])"
done
