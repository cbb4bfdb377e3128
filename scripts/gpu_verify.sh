#!/bin/bash
# full state check: smoke, the whole -m gpu suite, bench (tf32 with layer table + cpu baseline), ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== bench tf32"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision tf32 --layers-out gpurun_out/layers_tf32.json > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
tail -c 2500 gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
echo "=== ncu launch list"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_tf32.csv
