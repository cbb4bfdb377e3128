#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_conv.py -m gpu -q --timeout 120 2>&1 | tail -4
for prec in ${PRECS:-tf32}; do
echo "=== bench $prec"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu-baseline --layers-out gpurun_out/layers_$prec.json > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$prec.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['hbm']['achieved'])"; tail -3 gpurun_out/bench_$prec.err
done
