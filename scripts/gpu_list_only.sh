#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG:-x}.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_ncu_${TAG:-x}.json > gpurun_out/ncu_bench_${TAG:-x}.log 2>&1
wc -l gpurun_out/launches_${TAG:-x}.csv
