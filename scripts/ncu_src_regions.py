#!/usr/bin/env python
"""Samples per code region of an `ncu --page source --csv` dump: prints every instruction with >= N samples plus the marker
instructions (TMA / MMA / TMEM / barriers), so that the warp roles can be told apart.  usage: ncu_src_regions.py src.csv [N]"""
import csv, gzip, sys
opener = (lambda p: gzip.open(p, 'rt', errors='replace')) if sys.argv[1].endswith('.gz') else (lambda p: open(p, errors='replace'))
rows = list(csv.reader(opener(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
iS, iSrc = hdr.index('# Samples'), hdr.index('Source')
marks = ('UTMALDG', 'UTCHMMA', 'STTM', 'LDTM', 'UTCBAR', 'BAR.SYNC', 'SYNCS.ARRIVE', 'SYNCS.PHASECHK', 'UTMASTG', 'UBLKCP', 'EXIT', 'STG', 'LDG')
acc = 0
for i, r in enumerate(data):
    s = int(r[iS]); acc += s
    src = r[iSrc].strip()
    if s >= n or any(m in src for m in marks):
        print(f'{i:5d} {s:6d} cum={acc:7d}  {src[:90]}')
