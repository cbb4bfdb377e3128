#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -${TAIL:-8}
for prec in ${PRECS:-tf32}; do
echo "=== bench $prec"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision $prec ${BENCH_ARGS:---no-cpu-baseline} --layers-out gpurun_out/layers_$prec.json > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$prec.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['other_class']['bound'], d['roofline']['other_class']['frac'], d['clocks'])"; tail -3 gpurun_out/bench_$prec.err
done
