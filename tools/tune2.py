#!/usr/bin/env python
"""Second tuning sweep: async vs register-staged loads, L2 group size, factorizations."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan  # noqa: E402


def run(name, sfx, n_log, batch, nbuf, reps, configs):
    n = 1 << n_log
    rows = []
    for env in configs:
        for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_ASYNC", "PHASTFT_L2_GROUP_MB"):
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in env.items()})
        try:
            ms, desc = time_plan(sfx, n, batch, nbuf, reps)
        except Exception as e:  # noqa: BLE001
            print(f"{name} {env}: FAILED {e}", flush=True)
            continue
        us = ms * 1e3
        rows.append((us, env))
        print(f"{name} {us:9.2f} us {batch * n / us / 1e3:7.1f} Gpt/s  {env} | {desc[:200]}", flush=True)
    rows.sort(key=lambda r: r[0])
    print(f"== best {name}: {rows[0][0]:.2f} us {rows[0][1]}", flush=True)


what = sys.argv[1] if len(sys.argv) > 1 else "all"
F = "PHASTFT_FACTORS"; A = "PHASTFT_ASYNC"; C = "PHASTFT_TILE_C"; G = "PHASTFT_L2_GROUP_MB"
if what in ("all", "f64_20"):
    cfgs = []
    for a in (0, 1):
        for f in ("20:10,10", "20:7,7,6", "20:6,6,8", "20:8,6,6"):
            for c in (8, 16) if f != "20:10,10" else (4, 8):
                cfgs.append({F: f, A: a, C: c, G: 0})
    run("f64 2^20", "f64", 20, 1, 16, 50, cfgs)
if what in ("all", "f64_26"):
    cfgs = []
    for a in (0, 1):
        for f in ("26:9,9,8", "26:10,8,8", "26:9,8,9", "26:8,9,9", "26:8,8,10"):
            for g in (0, 32, 64):
                for c in (8, 16):
                    if c == 16 and g == 0:
                        continue
                    cfgs.append({F: f, A: a, C: c, G: g})
    run("f64 2^26", "f64", 26, 1, 1, 3, cfgs)
if what in ("all", "f32_16b"):
    cfgs = []
    for a in (0, 1):
        for f in ("16:8,8", "16:7,9", "16:9,7", "16:6,10"):
            for c in (16, 32):
                cfgs.append({F: f, A: a, C: c})
    run("f32 4096x2^16", "f32", 16, 4096, 1, 3, cfgs)
if what in ("all", "f64_23"):
    cfgs = []
    for a in (0, 1):
        for f in ("23:8,8,7", "23:7,8,8", "23:8,7,8"):
            for g in (0, 16, 32, 64):
                cfgs.append({F: f, A: a, C: 8, G: g})
    run("f64 2^23", "f64", 23, 1, 2, 5, cfgs)
