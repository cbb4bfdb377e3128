#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_function.py -m gpu -q --timeout 120 2>&1 | tail -5
echo "=== bench tf32"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision tf32 --layers-out gpurun_out/layers_tf32.json > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
tail -c 3000 gpurun_out/bench_tf32.json; tail -5 gpurun_out/bench_tf32.err
echo "=== bench fp32 (3xTF32)"
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline --layers-out gpurun_out/layers_fp32.json > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
tail -c 1500 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
echo "=== ncu launch list"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches_tf32.csv
