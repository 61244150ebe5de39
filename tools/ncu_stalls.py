#!/usr/bin/env python
"""Top stalled instructions of a kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_stalls.py report.ncu-rep kernel_regex [n_top] [instance]"""
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
print("instances:", len(hdr_idx))
h = rows[hdr_idx[inst]]
end = hdr_idx[inst + 1] - 1 if len(hdr_idx) > inst + 1 else None
body = rows[hdr_idx[inst] + 1:end]
si = h.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[si]) for r in body if len(r) > si and r[si].isdigit())
print('total samples', tot)
top = sorted([r for r in body if len(r) > si and r[si].isdigit()], key=lambda r: -int(r[si]))[:ntop]
for r in top:
    reasons = [(h[j], int(r[j])) for j in range(len(h)) if j > si + 2 and h[j].startswith('stall_')
               and 'Not' not in h[j] and r[j].isdigit() and int(r[j]) > 0]
    reasons.sort(key=lambda x: -x[1])
    print(f"{int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  {r[1].strip()[:58]:58s} {reasons[:2]}")
