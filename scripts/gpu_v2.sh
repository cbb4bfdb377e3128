#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_conv.py -m gpu -q --timeout 60 -k "tf32 or fused" 2>&1 | tail -15
echo "=== bench tf32"
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --precision tf32 --no-cpu-baseline --layers-out gpurun_out/layers_tf32.json > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['hbm']['achieved'])"; tail -3 gpurun_out/bench_tf32.err
