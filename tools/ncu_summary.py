#!/usr/bin/env python
"""Key metrics of every kernel instance in an .ncu-rep (ncu --set full), as a markdown table.
    python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma subpipe active %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor mem active % (elapsed)"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]
print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
print("|---|" + "---|" * len(WANT))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        cells = []
        for key, _ in WANT:
            if key in h:
                i = h.index(key)
                cells.append(f"{r[i]} {units[i]}".strip())
            else:
                cells.append("-")
        name = r[ki].split("(")[0].replace("mmbs::", "")[:60]
        print(f"| `{name}` | " + " | ".join(cells) + " |")
