#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'pipe_fma', 'pipe_alu', 'pipe_lsu', 'pipe_xu',
        'issue_active', 'hit_rate', 'sm__throughput.avg.pct', 'launch__grid_size', 'launch__block_size',
        'shared_mem_per_block', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'pipe_tensor',
        'warp_issue_stalled', 'l1tex__data_bank_conflicts', 'shared_ld_bank', 'shared_st_bank', 'local_load', 'local_store',
        'smsp__thread_inst_executed_per_inst_executed', 'derived__smsp__inst_executed_op_local']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
lines = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    lines.append(f'## {name}\n\n| metric | unit | value |\n|---|---|---|')
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS):
            lines.append(f'| {h} | {u} | {v} |')
text = '\n'.join(lines)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(f'# ncu --set full summary of {sys.argv[1]}\n\n' + text + '\n')
print(text)
