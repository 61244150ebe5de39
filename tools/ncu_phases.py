#!/usr/bin/env python
"""Split a kernel of an .ncu-rep (--set full --import-source on) into the phases between its block barriers and print
warp instructions executed / stall samples per phase.   python tools/ncu_phases.py report.ncu-rep kernel_regex"""
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) > 10 and r[0].startswith('0x')]
ie, ss = h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
tot_i = sum(int(r[ie]) for r in body)
tot_s = sum(int(r[ss]) for r in body)
print(f"total warp instructions {tot_i}  samples {tot_s}  SASS lines {len(body)}")
ph_i = ph_s = 0
start = 0
mix = {}
for k, r in enumerate(body):
    ph_i += int(r[ie])
    ph_s += int(r[ss])
    op = r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
    op = op.split('.')[0]
    mix[op] = mix.get(op, 0) + int(r[ie])
    if 'BAR.' in r[1] or k == len(body) - 1:
        top = sorted(mix.items(), key=lambda x: -x[1])[:6]
        print(f"lines {start:5d}-{k:5d}  inst {ph_i:10d} ({100*ph_i/tot_i:4.1f}%)  samples {ph_s:6d} ({100*ph_s/max(tot_s,1):4.1f}%)  "
              + " ".join(f"{a}:{b*100//max(ph_i,1)}%" for a, b in top))
        ph_i = ph_s = 0
        start = k + 1
        mix = {}
