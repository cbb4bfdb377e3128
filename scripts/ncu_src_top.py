#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions with their stall reasons.
usage: ncu_src_top.py src.csv [N [section]]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = starts[sec]
end = starts[sec + 1] - 1 if sec + 1 < len(starts) else len(rows)
print('section', sec, 'of', len(starts), rows[hi - 1][1][:90] if hi else '')
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
ic = [hdr.index(c) for c in cols]
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {c: sum(int(r[j]) for r in data) for c, j in zip(cols, ic)}
print('by reason:', ' '.join(f'{c[6:]}={v}' for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:n]
for i in sorted(top):
    r = data[i]
    st = ' '.join(f'{c[6:]}={r[j]}' for c, j in zip(cols, ic) if int(r[j]) > 0)
    print(f'{i:5d} {int(r[iS]):6d} {int(r[iEx]):9d}  {r[iSrc].strip()[:72]:72s} | {st}')
