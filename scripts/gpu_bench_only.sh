#!/bin/bash
mkdir -p gpurun_out
for prec in ${PRECS:-tf32}; do
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision $prec --no-cpu-baseline --layers-out gpurun_out/layers_${prec}${TAG}.json > gpurun_out/bench_${prec}${TAG}.json 2> gpurun_out/bench_${prec}${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${prec}${TAG}.json')); print('$prec$TAG value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'TF', d['roofline']['achieved'])"; tail -3 gpurun_out/bench_${prec}${TAG}.err
done
