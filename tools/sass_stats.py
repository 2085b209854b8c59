#!/usr/bin/env python
"""Per-kernel SASS instruction mix of libstg.so (cuobjdump -sass): whole function and the hottest loop body.

    python tools/sass_stats.py <substring of mangled kernel name> [--dump]
"""
import collections
import re
import subprocess
import sys

LIB = "spin_torque_rl_gym_b200/libstg.so"


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, name = [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
        elif name:
            cur.append(line)
    if name:
        yield name, cur


def parse(lines):
    ins = []
    for l in lines:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            addr = int(m.group(1), 16)
            txt = m.group(2).strip()
            op = re.sub(r"^@!?U?P\w+\s+", "", txt).split()[0]
            ins.append((addr, op, txt))
    return ins


def main():
    pat = sys.argv[1]
    for name, lines in functions():
        if pat not in name:
            continue
        ins = parse(lines)
        print(f"== {name}: {len(ins)} instructions")
        # backward branches = loops; pick the one enclosing the most instructions that is innermost-largest
        loops = []
        for addr, op, txt in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", txt)
                if m and int(m.group(1), 16) < addr:
                    loops.append((int(m.group(1), 16), addr))
        for lo, hi in sorted(loops, key=lambda t: t[0] - t[1])[:4]:
            body = [i for i in ins if lo <= i[0] <= hi]
            c = collections.Counter(re.sub(r"\..*", "", op) for _, op, _ in body)
            print(f"-- loop 0x{lo:x}..0x{hi:x}: {len(body)} instr: " + ", ".join(f"{k}:{v}" for k, v in c.most_common(24)))
        if "--dump" in sys.argv:
            for addr, op, txt in ins:
                print(f"{addr:06x}  {txt}")


if __name__ == "__main__":
    main()
