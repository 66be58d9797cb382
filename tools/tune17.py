import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
    os.environ.pop(k, None)
for sfx in ("f64", "f32"):
    for n in (10, 11, 12):
        for fac in (None, f"{n}:{n//2},{n-n//2}"):
            if fac: os.environ["PHASTFT_FACTORS"] = fac
            else: os.environ.pop("PHASTFT_FACTORS", None)
            try:
                ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
                print(f"{sfx} 2^{n} {fac}: {ms*1e3:8.2f} us | {d[12:150]}", flush=True)
            except Exception as e:
                print(sfx, n, fac, "FAILED", str(e)[:100])
    os.environ.pop("PHASTFT_FACTORS", None)
    for n in range(13, 23):
        ms, d = time_plan(sfx, 1 << n, 1, 16 if n <= 20 else 4, 40 if n <= 20 else 10)
        print(f"{sfx} 2^{n}: {ms*1e3:8.2f} us {(1<<n)/ms/1e6:6.1f} Gpt/s | {d[12:170]}", flush=True)
