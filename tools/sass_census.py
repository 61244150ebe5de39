#!/usr/bin/env python
"""Per-kernel SASS instruction census of libmmbs.so: the Blackwell-native mnemonics (B200_PROFILING.md "What proves a
Blackwell-native kernel") counted per kernel with `cuobjdump -sass`.  Runs without a GPU.

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodalbrainsurvival_b200", "libmmbs.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "REDG|RED\\.", "ATOMS", "VOTE", "MATCH"]


def demangle(names):
    try:
        r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True)
        return r.stdout.splitlines()
    except Exception:
        return names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for c in COLS:
            if re.match(c, op):
                counts[cur][c] += 1
    names = demangle(list(counts))
    print("# SASS census of libmmbs.so (sm_100a), `cuobjdump -sass`, one row per kernel\n")
    print("`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG`/`UTMASTG` = TMA tensor load/store, "
          "`SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (must be 0).\n")
    print("| kernel | SASS instrs | " + " | ".join(c.replace("|", "/").replace("\\", "") for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    tot = collections.Counter()
    for raw, nice in zip(counts, names):
        c = counts[raw]
        short = re.sub(r"\(.*", "", nice).replace("void ", "").replace("mmbs::", "")
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) for k in COLS) + " |")
        tot.update(c)
    print(f"| **total ({len(counts)} kernels)** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
