#!/bin/bash
# Round-2 state check: full -m gpu suite, smoke, bench (tf32 headline + fp32-exact), compute-sanitizer memcheck + racecheck
# on smoke().  Every step has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== full gpu suite"; date +%s
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; date +%s
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
if [ "${BENCH:-1}" = 1 ]; then
echo "=== bench tf32"; date +%s
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_tf32.json > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
tail -c 2500 gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
fi
if [ "${SANITIZE:-0}" = 1 ]; then
echo "=== compute-sanitizer memcheck (smoke)"; date +%s
B2J_SMOKE_PRECISIONS=fp32,tf32,simt timeout -s KILL 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
echo "=== compute-sanitizer racecheck (smoke)"; date +%s
B2J_SMOKE_PRECISIONS=fp32,tf32,simt timeout -s KILL 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
fi
date +%s
