#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
import collections


def load(f):
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith('==')]
    return list(csv.DictReader(lines))


rows = load(sys.argv[1])
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    idx = [i for i, r in enumerate(rows) if marker in r['Kernel Name']]
    rows = rows[idx[-1]:]
tot = 0.0
agg = collections.OrderedDict()
for r in rows:
    t = float(r['Metric Value']) / 1000
    tot += t
    name = r['Kernel Name'].split('(')[0][-44:]
    if len(sys.argv) > 3:
        print(f"{name:46s} grid {r['Grid Size']:14s} {t:8.1f} us")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
for k, (n, t) in agg.items():
    print(f"{k:46s} x{n:3d} {t:9.1f} us  {100*t/tot:5.1f}%")
print('total us', round(tot, 1))
