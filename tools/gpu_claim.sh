#!/bin/bash
# Dev script (GPU box): config 2 at several frames per launch (strong-scaling shares are short launches)
for spp in 1024 256 128 64; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-baselines --workload c2 --spp $spp | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spp $spp', round(d['ms_per_step'],4), 'ms', round(d['value'],1), d['verified'], d['verification'])"
done
