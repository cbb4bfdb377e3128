#!/usr/bin/env python
"""Per-kernel histogram of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA / TMEM / TMA / mbarrier), from
`cuobjdump -sass libb2jax.so`, plus the full SASS of one kernel.  usage: sass_summary.py libb2jax.so out.md [kernel-substring out.sass]"""
import collections, re, subprocess, sys
sass = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMAPF', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS.ARRIVE', 'SYNCS.PHASECHK',
        'LDGSTS', 'UBLKCP', 'MEMBAR.ALL.GPU', 'ERRBAR', 'STL', 'LDL', 'FMNMX3', 'REDUX', 'SHFL']
fn, per, order, body = None, {}, [], {}
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r'\(.*', '', fn).replace('void ', '').replace('b2j::', '')
        per[fn] = collections.Counter(); order.append(fn); body[fn] = []
        continue
    if fn is None:
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        per[fn]['total'] += 1
        body[fn].append(line.rstrip())
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k == 'UTMALDG' and op.startswith('UTMALDG')):
                per[fn][k] += 1
        if op.startswith('UTMALDG'):
            per[fn][op] += 1
with open(sys.argv[2], 'w') as f:
    f.write('# SASS evidence: instruction counts per kernel of libb2jax.so (cuobjdump -sass, sm_100a)\n\n')
    f.write('Blackwell paths: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2; operand list starting with `tmem[..]` twice = A from TMEM), `LDTM` / `STTM` = '
            'tcgen05.ld / tcgen05.st, `UTMALDG` = cp.async.bulk.tensor (TMA load; `.IM2COL`, `.2CTA` variants), `UTMAPF` = TMA L2 prefetch, `UTCBAR` = '
            'tcgen05.commit, `SYNCS.*` = mbarrier, `LDGSTS` = cp.async (residual staging of the 3xTF32 and RES2 epilogues), `UTMASTG` = cp.async.bulk.tensor store (TMA-store epilogue of the programs without a residual).  `MEMBAR.ALL.GPU` remains only in the cluster barrier at kernel start / end of the CTA-pair kernels.\n\n')
    cols = ['total'] + KEYS
    f.write('| kernel | ' + ' | '.join(cols) + ' | TMA load forms |\n|---|' + '---|' * (len(cols) + 1) + '\n')
    for fn in order:
        c = per[fn]
        if not any(c[k] for k in KEYS):
            continue
        forms = ', '.join(f'{k}×{v}' for k, v in sorted(c.items()) if k.startswith('UTMALDG.') )
        f.write(f'| `{fn}` | ' + ' | '.join(str(c[k]) if c[k] else '' for k in cols) + f' | {forms} |\n')
if len(sys.argv) > 4:
    for fn in order:
        if sys.argv[3] in fn:
            open(sys.argv[4], 'w').write(f'// {fn}\n' + '\n'.join(body[fn]) + '\n')
            break
