#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py list.csv [skip_first_n]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[hdr + 1 + skip:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else v if u in ('us', 'usecond') else v * 1e3 if u in ('ms', 'msecond') else v * 1e6
    a = agg.setdefault(r[ki][:80], [0, 0.])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f'{t / tot:6.3f} {t:10.1f}us n={n:5d} avg={t / n:7.1f}us  {k}')
