#!/bin/bash
# Dev script (multi-GPU box): the spp-split bench at N ranks.  usage: gpu_scale.sh <tag> "<N list>" "<workloads>"
tag=$1; ns=$2; wls=$3
mkdir -p gpurun_out/$tag
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in $wls; do for n in $ns; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 2 --warmup 3 --no-baselines --workload $wl > gpurun_out/$tag/scale_${wl}_n$n.json 2> gpurun_out/$tag/scale_${wl}_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 2 --warmup 3 --no-baselines --workload $wl > gpurun_out/$tag/scale_${wl}_n$n.json 2> gpurun_out/$tag/scale_${wl}_n$n.err
  fi
  tail -1 gpurun_out/$tag/scale_${wl}_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['n_gpus'], round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d['scaling'], 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/$tag/scale_${wl}_n$n.err
done; done
