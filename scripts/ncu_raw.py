#!/usr/bin/env python
"""Key per-launch metrics from an .ncu-rep: ncu_raw.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f'{w[:60]:60s} {units[i]:12s}', ' | '.join(r[i][:22] for r in data))
