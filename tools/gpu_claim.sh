#!/bin/bash
# Dev script (GPU box): pixel-claim batch size of the warp-queue form against frames per launch on config 2
for spp in 1024 128; do for ct in 0 1 2 4 8; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-baselines --workload c2 --spp $spp --claim-threshold $ct | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spp $spp claim $ct', round(d['ms_per_step'],4), 'ms', round(d['value'],1))"
done; done
