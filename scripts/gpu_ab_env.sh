#!/bin/bash
# A/B of one environment switch of libb2jax.so (default B2J_PDL): parity tests with it on, then bench.py without / with it.
# usage: gpu_ab_env.sh [VAR]     (scripts run through gpurun: bash scripts/gpu_ab_env.sh B2J_PDL)
VAR=${1:-B2J_PDL}
mkdir -p gpurun_out
echo "== tests, default environment"
timeout -s KILL 400 python -m pytest tests/test_conv.py tests/test_elegy_models.py tests/test_function.py tests/test_train_step.py -m gpu -q --timeout 300 2>&1 | tail -4
echo "== tests, $VAR=1"
env $VAR=1 timeout -s KILL 400 python -m pytest tests/test_conv.py tests/test_elegy_models.py tests/test_function.py tests/test_train_step.py -m gpu -q --timeout 300 2>&1 | tail -4
for mode in 0 1 0 1; do
  env $VAR=$mode timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_${VAR}_$mode.json > gpurun_out/bench_${VAR}_$mode.json 2> gpurun_out/bench_${VAR}_$mode.err
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_${VAR}_$mode.json'))
print('$VAR=$mode value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
for r in d['roofline']['bandwidth_kernels']['launches'][:3]:
    print('   ', r['op'], round(r['ms'], 4), 'ms', round(r['gbs']), 'GB/s', round(r['frac'], 3))
PY
  tail -2 gpurun_out/bench_${VAR}_$mode.err
done
