#!/usr/bin/env python
"""Batches of small transforms (2^24 points per call): one-CTA kernel candidates (PHASTFT_ROW_VARIANT) vs defaults."""
import os, sys
sys.path.insert(0, "tools")
from check_variant import check
from tune import time_plan

CANDS = {4: [80], 8: [80], 16: [80, 81], 512: [80, 81, 82, 83, 84], 1024: [70, 80, 81, 82, 83], 2048: [70, 80, 81, 82], 4096: [70, 80, 81]}
ok = True
for sfx in ("f64", "f32"):
    for n, ids in CANDS.items():
        for v in ids:
            if n == 4096 and sfx == "f64" and v == 81:
                continue
            try:
                ok &= check(sfx, n.bit_length() - 1, {"PHASTFT_ROW_VARIANT": str(v)}, batch=max(2, (1 << 21) // n + 3))
            except Exception as e:  # noqa: BLE001
                print("FAILED", sfx, n, v, e, flush=True); ok = False
print("ALL OK" if ok else "FAILURES", flush=True)
for sfx, esz in (("f64", 8), ("f32", 4)):
    for n, ids in CANDS.items():
        batch = (1 << 24) // n
        for v in [None] + ids:
            os.environ.pop("PHASTFT_ROW_VARIANT", None)
            if v is not None:
                os.environ["PHASTFT_ROW_VARIANT"] = str(v)
            try:
                ms, d = time_plan(sfx, n, batch, 3, 5)
            except Exception as e:  # noqa: BLE001
                print(sfx, n, v, "FAILED", e, flush=True); continue
            tb = (1 << 24) * esz * 4 / (ms * 1e-3) / 1e12
            print(f"{sfx} n={n:5d} variant={v}: {ms*1e3:8.1f} us {tb:5.2f} TB/s | {d[-70:]}", flush=True)
