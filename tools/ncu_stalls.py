#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep, aggregated per opcode (source page, needs -lineinfo + --import-source on).

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
by_op = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows:
    src = r["Source"].strip()
    op = re.sub(r"^@!?U?P\w+\s+", "", src).split()[0] if src else "?"
    op = re.sub(r"\.(reuse|F32x2|HI_LO)", "", op)
    n = int(r["# Samples"] or 0)
    ex = int(r["Instructions Executed"] or 0)
    by_op[op]["samples"] += n
    by_op[op]["executed"] += ex
    tot["samples"] += n
    tot["executed"] += ex
    for c in stall_cols:
        v = int(r[c] or 0)
        by_op[op][c] += v
        tot[c] += v
print(f"total samples {tot['samples']}, warp instructions {tot['executed']}")
print("stall mix: " + ", ".join(f"{c[6:]} {100 * tot[c] / max(tot['samples'], 1):.1f}%" for c in sorted(stall_cols, key=lambda c: -tot[c])[:9]))
print(f"{'opcode':22s} {'exec %':>7s} {'samples %':>9s}  top stalls")
for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((c[k], k[6:]) for k in stall_cols if c[k]), reverse=True)[:3]
    print(f"{op:22s} {100 * c['executed'] / max(tot['executed'], 1):7.1f} {100 * c['samples'] / max(tot['samples'], 1):9.1f}  "
          + ", ".join(f"{k} {100 * v / max(c['samples'], 1):.0f}%" for v, k in st))
