#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 0 -c 7 -o gpurun_out/prof_conv_v2 -f python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-200
