#!/bin/bash
# ncu --set full of launches of kernel regex $K: skip $S, count $C -> gpurun_out/$OUT.ncu-rep
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:${K} -s ${S:-0} -c ${C:-2} -o gpurun_out/${OUT:-prof_k} -f python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-fp32-variant > gpurun_out/ncu_k.log 2>&1
tail -2 gpurun_out/ncu_k.log | cut -c1-200; ls -la gpurun_out/${OUT:-prof_k}.ncu-rep
