#!/usr/bin/env python
"""Registers / spills per kernel of one .cu file (nvcc -Xptxas -v for sm_100a), demangled.

    python tools/ptxas_report.py spin_torque_rl_gym_b200/csrc/stt_kernels.cu [-DNAME=VALUE ...] [--filter substring]
"""
import re
import subprocess
import sys

args = sys.argv[1:]
flt = None
if "--filter" in args:
    k = args.index("--filter")
    flt = args[k + 1]
    del args[k:k + 2]
src, defs = args[0], args[1:]
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v", "-c", src,
       "-o", "/dev/null"] + defs
err = subprocess.run(cmd, capture_output=True, text=True).stderr
names = re.findall(r"Compiling entry function '(\S+)'", err)
if not names:
    sys.exit(err)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"Compiling entry function '\S+' for 'sm_100a'", err)[1:]
for name, b in zip(dem, blocks):
    if flt and flt not in name:
        continue
    regs = re.search(r"Used (\d+) registers", b)
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    short = re.sub(r"\(StgSttStepArgs\)|\(StgSttSolveArgs\)|stg::|void ", "", name)
    print(f"{short:70s} regs {regs.group(1) if regs else '?':>4s}  stack {spill.group(1):>4s}  spill st/ld {spill.group(2)}/{spill.group(3)}")
if "error" in err:
    print(err)
