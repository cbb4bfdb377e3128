#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:eltwise' -c 3 -o gpurun_out/prof_elt -f python bench.py --config bandwidth --steps 1 --warmup 3 > gpurun_out/ncu_elt.log 2>&1
ls -la gpurun_out/*.ncu-rep
CONFIGS="bandwidth" STEPS=10 bash scripts/gpu_r2_configs.sh 2>&1 | head -12
