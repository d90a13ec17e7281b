#!/bin/bash
# Dev script (multi-GPU box): the N>1 correctness test, then the bench at N ranks with both reduce transports.  usage: gpu_mgpu.sh <tag> <N> [extra bench args]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out/$tag
export ATX_P2P_TIMEOUT_MS=20000
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
for red in auto nccl tiles; do
  extra="--reduce $red --no-baselines"; [ "$red" = "tiles" ] && extra="--split tiles --no-baselines"; [ "$red" = "auto" ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 $extra "$@" > gpurun_out/$tag/bench_n${n}_$red.json 2> gpurun_out/$tag/bench_n${n}_$red.err
  tail -1 gpurun_out/$tag/bench_n${n}_$red.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$red', d['n_gpus'], round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms', d['kernel']['reduce'], 'verified', d['verified'], d['verification'])
for k,v in (d.get('strong_scaling') or {}).items(): print('   ', k, round(v['ms_per_step'],3), 'ms', round(v['value'],1), v['reduce'], v['verified'])
" || tail -20 gpurun_out/$tag/bench_n${n}_$red.err
done
