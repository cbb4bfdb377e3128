#!/bin/bash
# fp32-exact (3xTF32) mode: per-layer table + ncu full capture of a few conv_tc2 X3 launches
mkdir -p gpurun_out
echo "=== bench fp32 layers"; date +%s
timeout -s KILL 600 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_fp32.json > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/layers_fp32.json'))
print('step', d['step_ms_graph'], 'conv', d['conv_ms'], d['other_ms'])
for l in d['layers']:
    print('%-8s M=%8d N=%5d K=%5d ms=%.4f roof=%.4f frac=%.2f %s TF=%.0f GBs=%.0f' % (l['path'], l['M'], l['N'], l['K'], l['ms'], l['roofline_ms'], l['roofline_ms'] / l['ms'], l['bound'], l['tflops'], l['gbs']))
PY
for spec in ${SPECS:-16:2}; do
  s=${spec%%:*}; c=${spec##*:}
  echo "=== ncu full, launches $s +$c"; date +%s
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:conv_tc2|conv_patch' -s $((54 + s)) -c $c -o gpurun_out/prof_x3_l$s -f python bench.py --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_x3_l$s.log 2>&1
done
ls -la gpurun_out/*.ncu-rep; date +%s
