#!/bin/bash
# ncu --set full (+source) of the first conv_tc2 launches of one ResNet-50 step
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s ${SKIP:-0} -c ${COUNT:-5} -o gpurun_out/${OUT:-prof_tc2} -f python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-fp32-variant > gpurun_out/ncu_full_tc2.log 2>&1
tail -2 gpurun_out/ncu_full_tc2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
