#!/usr/bin/env python
"""Kernel-variant sweep (PHASTFT_VARIANT) on plans built from R=512 / R=256 f64 passes."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan  # noqa: E402

for name, n_log, nbuf, reps, fac in (("f64 2^26", 26, 1, 3, "26:9,9,8"), ("f64 2^24", 24, 1, 5, "24:8,8,8"), ("f64 2^18", 18, 16, 50, "18:9,9")):
    rows = []
    for var in range(0, 12):
        os.environ["PHASTFT_FACTORS"] = fac
        os.environ["PHASTFT_VARIANT"] = str(var)
        os.environ["PHASTFT_TILE_C"] = "8"
        try:
            ms, desc = time_plan("f64", 1 << n_log, 1, nbuf, reps)
        except Exception as e:  # noqa: BLE001
            print(name, var, "FAILED", e, flush=True); continue
        rows.append((ms, var))
        print(f"{name} variant {var:2d}: {ms*1e3:9.2f} us | {desc[:160]}", flush=True)
    rows.sort()
    print(f"== best {name}: variant {rows[0][1]} {rows[0][0]*1e3:.2f} us", flush=True)
