import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
def setenv(env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
for sfx, cs in (("f32", (8, 16)), ("f64", (4, 8))):
    for n in range(13, 21):
        rows = []
        for a in sorted({n // 2, (n + 1) // 2}):
            for c1 in cs:
                for c2 in cs:
                    setenv({"PHASTFT_FACTORS": f"{n}:{a},{n-a}", "PHASTFT_PASS_C": f"{c1},{c2}"})
                    try:
                        ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
                    except Exception as e:
                        continue
                    rows.append((ms * 1e3, a, n - a, c1, c2))
        rows.sort()
        print(f"{sfx} 2^{n}: " + "  ".join(f"({r[1]},{r[2]}) C={r[3]},{r[4]}: {r[0]:.2f}" for r in rows[:4]) + f"  worst {rows[-1][0]:.2f}", flush=True)
