#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 0 -c 8 -o gpurun_out/prof_conv_v1 -f python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
