#!/usr/bin/env python
"""Print bench.py's --layers-out table against the per-layer roofline (measured peaks)."""
import json, sys
d = json.load(open(sys.argv[1]))
peak_t = float(sys.argv[2]) if len(sys.argv) > 2 else 699.3
peak_b = 6533.5
print('graph ms/step', round(d['step_ms_graph'], 3), 'profiled sum', round(d['step_ms_profiled_sum'], 3), 'conv ms', round(d['conv_ms'], 3))
tot = 0
seen = {}
for l in d['layers']:
    extra = 4 * l['M'] * l['N'] / 1e6 if '+add+add' in l['op'] else 0.0      # the residual read is real traffic too
    ideal = max(l['gflop'] / peak_t, (l['mbytes'] + extra) / peak_b)
    tot += ideal
    key = (l['M'], l['N'], l['K'], l['op'])
    seen.setdefault(key, []).append((l['ms'], ideal, l))
for (M, N, K, op), v in seen.items():
    ms = sum(x[0] for x in v) / len(v); ideal = v[0][1]; l = v[0][2]
    print(f"x{len(v)} M={M:8d} N={N:5d} K={K:5d} {ms:.3f} ms ideal {ideal:.3f} x{ms/ideal:.2f} {l['gflop']/ms:.0f} TF {op[20:]}")
print('ideal conv total', round(tot, 3))
for k, v in sorted(d['other_ms'].items(), key=lambda kv: -kv[1]):
    print(f"{v:.3f} {k[:100]}")
