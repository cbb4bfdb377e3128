#!/bin/bash
# bench.py --config c1|c2|c3|bandwidth (CONFIGS=...), outputs under gpurun_out/
mkdir -p gpurun_out
for c in ${CONFIGS:-bandwidth c1 c2 c3}; do
  echo "=== bench --config $c"; date +%s
  timeout -s KILL 600 python bench.py --config $c --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$c.json'))
    print(d['metric'], d['value'], d['unit'])
    for r in d['rows']:
        if 'kernels' not in r:
            print('  ', r); continue
        print('  %-100s graph %.4f ms  e2e %.3f ms' % (r['name'][:100], r['graph_ms'], r['e2e_ms_host_inputs']))
        for k in r['kernels']:
            print('      %-60s %.4f ms  %s GB/s %s TF  frac %s  %s' % (k['kernel'][:60], k['ms'], ('%.0f' % k['gbs']) if k.get('gbs') else '-', ('%.0f' % k['tflops']) if k.get('tflops') else '-', ('%.2f' % k['frac']) if k.get('frac') else '-', k.get('bound', k.get('path', ''))))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_$c.err').read()[-3000:])
PY
done
date +%s
