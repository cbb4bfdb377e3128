#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of bench.py into
  * a per-launch table of ONE steady-state step (markdown, for profiles/), and
  * profiles/r02_dram_traffic.json: measured DRAM bytes per launch of the tensor-bound / HBM-bound conv_tc2 launches,
    which bench.py reports as roofline.traffic.
usage: ncu_step_summary.py launches.csv layers.json out.md out.json"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
launches = {}
order = []
for r in rows[1:]:
    k = int(r[iID])
    if k not in launches:
        launches[k] = {'name': r[iK].split('(')[0].replace('void ', '').replace('b2j::', '')}
        order.append(k)
    launches[k][r[iM]] = float(r[iV].replace(',', ''))
seq = [launches[k] for k in order]
# a steady-state replay of the per-call graph: starts at the stem -- its re-layout (round 1) or the raw-row kernel
# conv_tc2_kernel<64, A_ROWS(2) / A_ROWS_U8(3), ...> (round 2) -- and ends before the next one / a weight_prep launch
def is_start(l):
    n = l['name']
    return n.startswith('relayout') or 'conv_tc2_kernel<64, 2,' in n or 'conv_tc2_kernel<64, 3,' in n
starts = [i for i, l in enumerate(seq) if is_start(l)]
pick = starts[min(3, len(starts) - 1)]
step = []
for l in seq[pick:]:
    if step and (is_start(l) or l['name'].startswith('weight_prep')):
        break
    step.append(l)
layers = json.load(open(sys.argv[2]))['layers']
convs = [l for l in step if l['name'].startswith(('conv_tc2', 'conv_patch'))]
assert len(convs) == len(layers), (len(convs), len(layers))
tot = sum(l['gpu__time_duration.sum'] for l in step)
with open(sys.argv[3], 'w') as f:
    f.write('| # | kernel | us (ncu, serialised, cold L2) | share | DRAM read MB | DRAM write MB | algorithmic MB | bound |\n|---|---|---|---|---|---|---|---|\n')
    ci = 0
    for i, l in enumerate(step):
        alg, bound = '', ''
        if l['name'].startswith(('conv_tc2', 'conv_patch')):
            alg, bound = f"{layers[ci]['mbytes']:.0f}", layers[ci]['bound']
            l['alg_mb'], l['bound'] = layers[ci]['mbytes'], layers[ci]['bound']
            ci += 1
        f.write(f"| {i} | {l['name'][:60]} | {l['gpu__time_duration.sum'] / 1e3:.1f} | {l['gpu__time_duration.sum'] / tot * 100:.1f}% | "
                f"{l.get('dram__bytes_read.sum', 0) / 1e6:.0f} | {l.get('dram__bytes_write.sum', 0) / 1e6:.0f} | {alg} | {bound} |\n")
    f.write(f'\nstep total under ncu: {tot / 1e6:.3f} ms over {len(step)} launches; conv_tc2 + conv_patch share '
            f"{sum(l['gpu__time_duration.sum'] for l in convs) / tot * 100:.1f}%\n")
out = {}
for cls in ('tensor', 'hbm'):
    sel = [l for l in convs if l['bound'] == cls]
    if sel:
        dram = sum(l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0) for l in sel)
        out[cls] = {'dram_bytes_per_launch': dram / len(sel), 'algorithmic_bytes_per_launch': sum(l['alg_mb'] for l in sel) * 1e6 / len(sel),
                    'launches': len(sel),
                    'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the %d %s-bound conv_tc2 launches of one '
                            'steady-state step (ncu --metrics pass, %s)' % (len(sel), cls, sys.argv[1].split('/')[-1])}
json.dump(out, open(sys.argv[4], 'w'), indent=1)
print(json.dumps(out, indent=1))
