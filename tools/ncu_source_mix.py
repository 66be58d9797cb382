#!/usr/bin/env python
"""Instruction mix and stall samples per opcode / per program region from an `ncu --page source --csv --print-source sass` export.
    python tools/ncu_source_mix.py gpurun_out/x.source.csv [region boundaries as instruction indices...]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
ci, si = idx['Instructions Executed'], idx['# Samples']
ops, samp = collections.Counter(), collections.Counter()
tot = stot = 0
seq = []
for r in rows[2:]:
    try:
        n = int(r[ci]); s = int(r[si])
    except (ValueError, IndexError):
        continue
    src = r[idx['Source']]
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    full = m.group(2) if m else src[:12]
    op = full.split('.')[0]
    if op in ('LDS', 'STS', 'LDG', 'STG', 'ST', 'LD', 'LDC', 'BAR'):
        op = '.'.join(full.split('.')[:3]) if op in ('LDG', 'STG', 'LD', 'ST') else full.split('.')[0] + ('.' + full.split('.')[1] if '.' in full else '')
    ops[op] += n; samp[op] += s; tot += n; stot += s
    seq.append((src.strip(), n, s))
print(f"warp instructions executed {tot}, stall samples {stot}")
for op, n in ops.most_common(45):
    print(f"{op:22s} {n:10d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(stot, 1):5.1f}%")
if '--top' in sys.argv:
    print("--- instructions with the most stall samples")
    for i, (src, n, s) in sorted(enumerate(seq), key=lambda t: -t[1][2])[:40]:
        print(f"{i:5d} {s:6d} {100 * s / stot:5.1f}%  x{n:8d}  {src[:100]}")
