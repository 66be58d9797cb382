#!/usr/bin/env python
"""2-pass size sweep (n = 13..20, both precisions) + batched f32 2^16: factor split and tile widths."""
import os, sys, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan  # noqa: E402

def setenv(env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "single"):
    for sfx, cn, cw in (("f64", 8, 16), ("f32", 16, 32)):
        for n in range(13, 21):
            rows = []
            for a in range(max(5, n - 10), min(10, n - 5) + 1):
                for pc in ((cw, cw), (cw, cn), (cn, cn), (cn, cw)):
                    setenv({"PHASTFT_FACTORS": f"{n}:{a},{n-a}", "PHASTFT_PASS_C": f"{pc[0]},{pc[1]}"})
                    nbuf = max(2, min(16, (256 << 20) // ((1 << n) * (16 if sfx == "f64" else 8))))
                    try:
                        ms, desc = time_plan(sfx, 1 << n, 1, nbuf, 30)
                    except Exception as e:
                        continue
                    rows.append((ms * 1e3, a, n - a, desc.split(":", 1)[1][:150]))
            rows.sort()
            for r in rows[:3]:
                print(f"{sfx} 2^{n}: {r[0]:8.2f} us ({r[1]},{r[2]}) {r[3]}", flush=True)
            print(f"== {sfx} 2^{n} best ({rows[0][1]},{rows[0][2]}) {rows[0][0]:.2f} us; worst {rows[-1][0]:.2f}", flush=True)
if what in ("all", "batch"):
    # batch config runs in a subprocess per chunk size (the chunk size is read once per process)
    for mb in (24, 48, 96):
        for fac, pc in (("16:7,9", "32,16"), ("16:7,9", "16,16"), ("16:8,8", "32,32"), ("16:8,8", "16,16"), ("16:6,10", "32,16"), ("16:9,7", "16,32")):
            env = dict(os.environ, PHASTFT_L2_CHUNK_MB=str(mb), PHASTFT_FACTORS=fac, PHASTFT_PASS_C=pc)
            code = ("import sys; sys.path.insert(0,'tools'); from tune import time_plan; ms,d=time_plan('f32',1<<16,4096,1,3);"
                    f"print('f32 4096x2^16 chunk {mb} MiB {fac} C={pc}: %.1f us %.1f Gpt/s | %s' % (ms*1e3, 4096*65536/ms/1e6, d[:120]))")
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            print(out.stdout.strip() or out.stderr[-300:], flush=True)
