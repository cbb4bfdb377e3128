#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
grep -v "^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -25 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism'])
except Exception as e: print('no json', e)
PY
