#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep attributed to CUDA source lines: the SASS page of the report (per-instruction
samples) joined with `nvdisasm -g` line markers of the same kernel in libstg.so (built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <mangled kernel substring> [top] [--lib path]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, pat = args[0], args[1]
top = int(args[2]) if len(args) > 2 else 30
lib = "spin_torque_rl_gym_b200/libstg.so"
if "--lib" in sys.argv:
    lib = sys.argv[sys.argv.index("--lib") + 1]

# ---- line table of the kernel: byte offset -> (file, line); inlined frames report the innermost location
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of, fn_name = {}, None
for cub in sorted(os.listdir(tmp)):
    if cub.count("-") > 2:
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, loc = None, None
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur = m.group(1) if pat in m.group(1) else None
            fn_name = cur or fn_name
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m and loc:
            line_of[int(m.group(1), 16)] = loc
    if line_of:
        break
if not line_of:
    sys.exit(f"kernel matching {pat!r} not found in {lib}")

# ---- samples per SASS instruction from the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = min(int(r["Address"], 16) for r in rows)
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
by_line = collections.defaultdict(collections.Counter)
tot = 0
for r in rows:
    off = int(r["Address"], 16) - base
    loc = line_of.get(off, ("?", 0))
    n = int(r["# Samples"] or 0)
    by_line[loc]["samples"] += n
    by_line[loc]["executed"] += int(r["Instructions Executed"] or 0)
    tot += n
    for c in stall_cols:
        by_line[loc][c] += int(r[c] or 0)
src_cache = {}
def src(loc):
    f, ln = loc
    for root in ("spin_torque_rl_gym_b200/csrc", "include"):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[p]) else ""
    return ""
print(f"{fn_name}: {tot} samples")
for loc, c in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((c[k], k[6:]) for k in stall_cols if c[k]), reverse=True)[:2]
    print(f"{100 * c['samples'] / max(tot, 1):5.1f}%  {loc[0]}:{loc[1]:<4d} ex={c['executed']:>10d}  "
          + ", ".join(f"{k} {100 * v / max(c['samples'], 1):.0f}%" for v, k in st) + f"   | {src(loc)}")
