#!/usr/bin/env python
"""Side-by-side per-layer times of several bench.py --layers-out files: cmp_layers.py a.json b.json ..."""
import json, sys
tabs = [json.load(open(f)) for f in sys.argv[1:]]
print('graph ms/step:', [round(t['step_ms_graph'], 3) for t in tabs])
keys = []
rows = {}
for ti, t in enumerate(tabs):
    for l in t['layers']:
        k = (l['M'], l['N'], l['K'], l['op'].count('+add+add') > 0)
        if k not in rows:
            rows[k] = [[] for _ in tabs]; keys.append(k)
        rows[k][ti].append(l['ms'])
for k in keys:
    ms = [sum(v) / max(len(v), 1) for v in rows[k]]
    best = min(range(len(ms)), key=lambda i: ms[i])
    print(f"x{len(rows[k][0])} M={k[0]:8d} N={k[1]:5d} K={k[2]:5d} res={int(k[3])}  " + '  '.join(f'{m:.3f}' for m in ms) + f'   best={best}')
