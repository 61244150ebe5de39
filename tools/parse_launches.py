#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list.   python tools/parse_launches.py <csv> [first-kernel-marker] [v]"""
import collections
import csv
import sys

UNIT_US = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 's': 1e6}
UNIT_B = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def load(f):
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith('==')]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(r['ID'], {'name': r['Kernel Name'], 'grid': r['Grid Size'], 'us': 0.0, 'bytes': 0.0})
        v = float(r['Metric Value'].replace(',', ''))
        if r['Metric Name'].startswith('gpu__time'):
            d['us'] = v * UNIT_US.get(r['Metric Unit'], 1e-3)
        elif r['Metric Name'].startswith('dram__bytes'):
            d['bytes'] += v * UNIT_B.get(r['Metric Unit'], 1)
    return list(by.values())


rows = load(sys.argv[1])
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    idx = [i for i, r in enumerate(rows) if marker in r['name']]
    rows = rows[idx[-1]:]
tot = sum(r['us'] for r in rows)
agg = collections.OrderedDict()
for r in rows:
    name = r['name'].split('(')[0][-44:]
    if len(sys.argv) > 3:
        print(f"{name:46s} grid {r['grid']:14s} {r['us']:8.1f} us {r['bytes']/1e6:9.1f} MB dram")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += r['us']
    a[2] += r['bytes']
for k, (n, t, b) in agg.items():
    print(f"{k:46s} x{n:3d} {t:9.1f} us  {100*t/tot:5.1f}%  {b/1e6:9.1f} MB dram")
print('total us', round(tot, 1))
