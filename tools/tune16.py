import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
    os.environ.pop(k, None)
for sfx in ("f64", "f32"):
    for n in range(13, 23):
        ms, d = time_plan(sfx, 1 << n, 1, 16 if n <= 20 else 4, 40 if n <= 20 else 10)
        print(f"{sfx} 2^{n}: {ms*1e3:8.2f} us {(1<<n)/ms/1e6:6.1f} Gpt/s | {d[12:170]}", flush=True)
for pv, pc in (("0,0", "8,8"), ("60,60", "4,4"), ("0,0", "4,4"), ("60,60", "8,8")):
    os.environ["PHASTFT_PASS_VARIANT"] = pv; os.environ["PHASTFT_PASS_C"] = pc
    for sfx in ("f64", "f32"):
        if sfx == "f32": os.environ["PHASTFT_PASS_C"] = ",".join(str(2*int(x)) for x in pc.split(","))
        ms, d = time_plan(sfx, 1 << 20, 1, 16, 40)
        print(f"{sfx} 2^20 v={pv} C={os.environ['PHASTFT_PASS_C']}: {ms*1e3:8.2f} us | {d[12:170]}", flush=True)
