#!/usr/bin/env python
"""Markdown section(s) for profiles/rNN_ncu_full.md from `ncu --set full` reports.  Runs where the .ncu-rep files are (the GPU box:
the reports are ~20 MB each and gpurun_out only travels back under 64 MiB), writes for each report a section of the key raw metrics
plus the gzip-ed source page (per-instruction samples, for scripts/ncu_src_regions.py), then the caller can delete the reports.
usage: ncu_full_md.py out.md "title 1" rep1.ncu-rep ["title 2" rep2.ncu-rep ...]"""
import csv, gzip, io, subprocess, sys

ROWS = [('time', 'gpu__time_duration.sum'),
        ('tensor pipe active %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('IPC', 'sm__inst_executed.avg.per_cycle_active'),
        ('warp instructions', 'smsp__inst_executed.sum'),
        ('DRAM read', 'dram__bytes_read.sum'),
        ('DRAM write', 'dram__bytes_write.sum'),
        ('DRAM %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('L2->SM bytes', 'l1tex__m_xbar2l1tex_read_bytes.sum'),
        ('L2 hit %', 'lts__t_sector_hit_rate.pct'),
        ('smem wavefronts', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
        ('regs/thread', 'launch__registers_per_thread'),
        ('achieved warps %', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('stall membar', 'smsp__pcsamp_warps_issue_stalled_membar'),
        ('SM clock', 'sm__cycles_elapsed.avg.per_second')]


def section(title, rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    md = []
    for k, r in enumerate(data):
        get = lambda name: (r[hdr.index(name)], units[hdr.index(name)]) if name in hdr else None
        name = get('Kernel Name')[0]
        md.append(f'\n## {title}' + (f' (launch {k})' if len(data) > 1 else '') + '\n')
        md.append(f'`{name[:70]}` grid {get("Grid Size")[0]} block {get("Block Size")[0]}\n')
        md.append('| metric | value |\n|---|---|')
        for label, metric in ROWS:
            v = get(metric)
            if v is not None:
                md.append(f'| {label} | {v[0]} {v[1]} |')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    with gzip.open(rep.replace('.ncu-rep', '_src.csv.gz'), 'wt') as f:
        f.write(src)
    return '\n'.join(md) + '\n'


if __name__ == '__main__':
    out, args = sys.argv[1], sys.argv[2:]
    with open(out, 'a') as f:
        for title, rep in zip(args[0::2], args[1::2]):
            try:
                f.write(section(title, rep))
            except Exception as e:                       # a missing report must not lose the others
                f.write(f'\n## {title}\n\n(no report: {e})\n')
