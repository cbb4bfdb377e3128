#!/bin/bash
# One-call state check for the end of a round: pooling-kernel A/B, the new tests, bench (tf32 + layer table + cpu baseline),
# ncu launch list with DRAM bytes, then the whole -m gpu suite.  Every step has its own timeout; order = priority.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== pool A/B"; date +%s
timeout -s KILL 240 python scripts/pool_ab.py > gpurun_out/pool_ab.log 2>&1; cat gpurun_out/pool_ab.log | head -8
# adopt the pair kernel for the rest of the call when it is >= 3 % faster on the stem pool and bit-identical everywhere
export B2J_POOL_PAIR=$(python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/pool_ab.log').read().strip().splitlines()[-1])
    k = 'max 3x3 s2 SAME [256,112,112,64]'
    same = all(d['0'][n]['checksum'] == d['1'][n]['checksum'] for n in d['0'])
    print(1 if same and d['1'][k]['ms'] < 0.97 * d['0'][k]['ms'] else 0)
except Exception:
    print(0)
PY
)
echo "B2J_POOL_PAIR=$B2J_POOL_PAIR" | tee gpurun_out/pool_choice.txt
echo "=== pool tests with both kernels + the full-size C4 test"; date +%s
B2J_POOL_PAIR=1 timeout -s KILL 300 python -m pytest tests/test_reduce_window.py "tests/test_elegy_models.py::test_c3_pool_sweep_batch256" -q --timeout 120 2>&1 | tail -5 | tee gpurun_out/pytest_pool_pair.log
B2J_POOL_PAIR=0 timeout -s KILL 300 python -m pytest tests/test_reduce_window.py -q --timeout 120 2>&1 | tail -3 | tee gpurun_out/pytest_pool_one.log
B2J_POOL_PAIR=1 timeout -s KILL 400 python -m pytest "tests/test_elegy_models.py::test_c4_resnet50_batch256_full_size" -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_c4.log
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
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_elegy_models.py::test_c4_resnet50_batch256_full_size 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
date +%s
