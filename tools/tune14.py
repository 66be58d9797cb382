import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
def setenv(env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
for n, facs in ((20, ("20:10,10", "20:7,6,7", "20:8,4,8", "20:6,8,6")), (19, ("19:10,9", "19:9,10", "19:6,7,6")), (18, ("18:9,9", "18:8,10", "18:10,8", "18:6,6,6"))):
    for fac in facs:
        for pc in ("16,16", "8,8", "16,8", "8,16", "16,16,16", "32,16,32"):
            if len(pc.split(",")) != len(fac.split(":")[1].split(",")): continue
            setenv({"PHASTFT_FACTORS": fac, "PHASTFT_PASS_C": pc})
            try:
                ms, d = time_plan("f32", 1 << n, 1, 16, 50)
            except Exception as e:
                print("f32", n, fac, pc, "FAILED", str(e)[:80]); continue
            print(f"f32 2^{n} {fac} C={pc}: {ms*1e3:7.2f} us | {d[12:150]}", flush=True)
