#!/usr/bin/env python
"""Hot source lines (warp stall samples) from an .ncu-rep captured with --import-source on and -lineinfo.
usage: python profiles/ncu_hot_lines.py prof.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
agg = collections.OrderedDict()
stall_cols = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        si = hdr.index('# Samples')
        ii = hdr.index('Instructions Executed')
        stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if hdr is None or len(r) <= si or r[2] != '-':
        continue   # keep only the per-source-line aggregate rows (Address == '-')
    try:
        s = int(r[si]); n = int(r[ii])
    except ValueError:
        continue
    key = (cur_file, r[0])
    st = {h: int(r[i]) for h, i in stall_cols.items() if r[i].isdigit() and int(r[i]) > 0}
    if key in agg:
        agg[key][0] += s; agg[key][1] += n
        for h, v in st.items():
            agg[key][3][h] = agg[key][3].get(h, 0) + v
    else:
        agg[key] = [s, n, r[1].strip()[:100], st]
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print(f'total samples {tot}, warp instructions {toti}')
for (f, l), (s, n, txt, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    main = ','.join(f'{h[6:]}={v}' for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{100*s/tot:5.1f}% samp {100*n/toti:5.1f}% inst  {f}:{l:>4}  {txt}   [{main}]')
