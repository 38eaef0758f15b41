#!/usr/bin/env python3
"""Summarise an ncu launch list (gpu__time_duration.sum CSV): python tools_launches.py file.csv [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[h]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0]
    v = float(r[vi].replace(',', ''))
    v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(r[ui], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print('%-78s %5s %10s %9s %6s' % ('kernel', 'n', 'total ms', 'avg ms', 'share'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-78s %5d %10.3f %9.4f %5.1f%%' % (k[:78], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print('total %.3f ms over %d launches' % (tot, sum(v[0] for v in agg.values())))
