#!/bin/bash
# 3xTF32 A/B: tests, then fp32 bench per environment setting in ENVS (separated by ';')
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest ${TESTS:-tests/test_golden_fixtures.py tests/test_conv.py} -m gpu -q --timeout 300 -x 2>&1 | tail -5
IFS=';' read -ra E <<< "${ENVS:-B2J_NOP=1}"
i=0
for PREC in ${PRECS:-fp32}; do
for e in "${E[@]}"; do
  echo "=== bench $PREC [$e]"
  env $e timeout -s KILL 600 python bench.py --precision $PREC --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-variant --layers-out gpurun_out/layers_ab$i.json > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_ab$i.json')); print('ms_per_step', d['ms_per_step'], 'value', d['value'])
    L = json.load(open('gpurun_out/layers_ab$i.json'))
    print(' '.join('%d:%.3f' % (i, l['ms']) for i, l in enumerate(L['layers'])))
except Exception as ex:
    print('failed', ex); print(open('gpurun_out/bench_ab$i.err').read()[-2000:])
PY
  i=$((i+1))
done
done
