import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
def setenv(env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
rows = []
for fac in ("20:10,10", "20:7,6,7", "20:6,8,6", "20:8,4,8", "20:8,3,9", "20:7,5,8", "20:8,5,7", "20:6,7,7", "20:7,7,6", "20:8,6,6", "20:6,6,8"):
    np_ = len(fac.split(":")[1].split(","))
    for pc in (("8,8",) if np_ == 2 else ("4,4,4", "8,8,8", "4,8,4", "8,4,8", "16,8,16", "8,64,8", "4,64,4", "16,64,16")):
        setenv({"PHASTFT_FACTORS": fac, "PHASTFT_PASS_C": pc})
        try:
            ms, d = time_plan("f64", 1 << 20, 1, 16, 40)
        except Exception as e:
            continue
        rows.append((ms * 1e3, fac, pc, d[12:200]))
rows.sort()
for r in rows[:14]:
    print(f"{r[0]:7.2f} us {r[1]} C={r[2]} | {r[3]}", flush=True)
