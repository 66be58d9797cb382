import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
    os.environ.pop(k, None)
for sfx, n, batch in (("f32", 16, 4096), ("f64", 20, 64), ("f64", 16, 1024), ("f64", 13, 8192), ("f64", 12, 16384), ("f64", 11, 32768), ("f32", 12, 32768), ("f32", 20, 128), ("f64", 18, 256), ("f32", 8, 1 << 19), ("f64", 6, 1 << 20)):
    ms, d = time_plan(sfx, 1 << n, batch, 1, 3)
    esz = 8 if sfx == "f64" else 4
    print(f"{sfx} {batch}x2^{n}: {ms*1e3:9.1f} us {batch*(1<<n)/ms/1e6:7.1f} Gpt/s  alg {batch*(1<<n)*esz*4/ms/1e9:6.2f} TB/s | {d[12:220]}", flush=True)
