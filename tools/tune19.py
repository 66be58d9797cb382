import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
for sfx in ("f64", "f32"):
    for n, batch in ((12, 16384), (11, 32768), (10, 65536), (8, 1 << 18)):
        for v in (0, 70, 71, 72):
            os.environ["PHASTFT_VARIANT"] = str(v)
            try:
                ms, d = time_plan(sfx, 1 << n, batch, 1, 3)
            except Exception as e:
                print(sfx, n, v, "FAILED", str(e)[:80]); continue
            esz = 8 if sfx == "f64" else 4
            print(f"{sfx} {batch}x2^{n} v{v}: {ms*1e3:9.1f} us alg {batch*(1<<n)*esz*4/ms/1e9:6.2f} TB/s | {d[-60:]}", flush=True)
        for v in (0, 70, 71, 72):
            os.environ["PHASTFT_VARIANT"] = str(v)
            if n > (10 if sfx == "f64" else 12): continue
            ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
            print(f"{sfx} single 2^{n} v{v}: {ms*1e3:9.2f} us | {d[-60:]}", flush=True)
