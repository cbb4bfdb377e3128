#!/bin/bash
# One-call state check for the end of a round: pooling micro-benchmark, tests of the non-default kernel switches, bench (tf32 +
# layer table + cpu baseline), the reference arm, ncu launch list with DRAM bytes, then smoke() and the whole -m gpu suite.  Every step has its own timeout; order = priority.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== pooling micro-benchmark (32- vs 64-bit index math) + pooling tests with the 64-bit instantiation"; date +%s
timeout -s KILL 240 python scripts/pool_ab.py B2J_POOL_IDX64 > gpurun_out/pool_ab.log 2>&1; head -5 gpurun_out/pool_ab.log
B2J_POOL_IDX64=1 timeout -s KILL 300 python -m pytest tests/test_reduce_window.py "tests/test_elegy_models.py::test_c3_pool_sweep_batch256" -q --timeout 120 2>&1 | tail -3 | tee gpurun_out/pytest_pool_idx64.log
echo "=== conv / model tests with programmatic dependent launch on"; date +%s
B2J_PDL=1 timeout -s KILL 300 python -m pytest tests/test_conv.py tests/test_elegy_models.py -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/pytest_pdl.log
echo "=== experiment: CTA pairs for the 64-wide tiles too (B2J_TC2_CG=4: halves the weight-tile share of the L2 -> SM traffic)"; date +%s
for mode in 0 4 0 4; do
  B2J_TC2_CG=$mode timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_cg$mode.json > gpurun_out/bench_cg$mode.json 2> gpurun_out/bench_cg$mode.err
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_cg$mode.json')); L = json.load(open('gpurun_out/layers_cg$mode.json'))['layers']
print('B2J_TC2_CG=$mode step', round(d['ms_per_step'], 4), 'ms; N<=64 layers:', [round(l['ms'], 4) for l in L if l['N'] <= 64])
PY
done
B2J_TC2_CG=4 timeout -s KILL 300 python -m pytest tests/test_conv.py tests/test_elegy_models.py -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/pytest_cg4.log
echo "=== bench tf32"; date +%s
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --precision tf32 --layers-out gpurun_out/layers_tf32.json > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
tail -c 3000 gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
echo "=== bench --impl reference"; date +%s
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
echo "=== ncu launch list"; date +%s
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 3 --precision tf32 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_ncu.json > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_tf32.csv
echo "=== smoke + full gpu suite"; date +%s
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
date +%s
