#!/bin/bash
# quick iteration: selected tests (TESTS=...) + optional short bench (BENCH=1)
mkdir -p gpurun_out
echo "=== tests: ${TESTS:-tests/test_golden_fixtures.py tests/test_conv.py}"; date +%s
timeout -s KILL 900 python -m pytest ${TESTS:-tests/test_golden_fixtures.py tests/test_conv.py} -m gpu -q --timeout 300 ${PYTEST_ARGS:-} 2>&1 | tail -${TAIL:-40} | tee gpurun_out/pytest_quick.log
if [ "${BENCH:-0}" = 1 ]; then
echo "=== bench"; date +%s
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} --layers-out gpurun_out/layers_quick.json > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_quick.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'dtype') if k in d}, 'e2e', d.get('e2e', {}).get('value'))
    for k in ('fp32_exact_ms_per_step', 'fp32_exact_images_per_s', 'prologue_replayed_ms_per_step', 'tf32_peak_measured_tflops'):
        if k in d: print(k, d[k])
    r = d.get('roofline', {})
    print('roofline', {k: r.get(k) for k in ('bound', 'achieved', 'peak', 'frac')}, 'all', r.get('all_launches', {}).get('frac'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
fi
date +%s
