#!/usr/bin/env python
"""Static instruction mix of one kernel from `cuobjdump -sass` (no GPU needed): python tools/sass_mix.py <object or .so> <regex on the mangled kernel name>"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2])
cur, keep = None, False
ops = collections.Counter()
n = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        if keep:
            break
        keep = bool(pat.search(cur))
        if keep:
            print("kernel:", cur[:200])
        continue
    if not keep:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        full = m.group(2)
        op = full.split(".")[0]
        if op in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "LDL", "STL"):
            op = ".".join(full.split(".")[:3])
        ops[op] += 1
        n += 1
print("static instructions:", n)
cls = collections.Counter()
FP = {"DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "HFMA2", "FFMA2", "FADD2", "FMUL2"}
MEM = lambda o: o.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL")
for op, c in ops.items():
    cls["fp" if op in FP else "mem" if MEM(op) else "other"] += c
print({k: f"{v} ({100 * v / n:.0f}%)" for k, v in cls.items()})
for op, c in ops.most_common(28):
    print(f"  {op:18s} {c:6d} {100 * c / n:5.1f}%")
