#!/usr/bin/env python
"""After hoisting the layout tests: half-width 1024-row tiles with interleaved intermediates, f32 2^20 regression."""
import sys, os
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
from tune import time_plan

IL = {"PHASTFT_WS_IL": "1"}
ok = True
for v in (61, 62, 63, 64, 65, 66):
    ok &= check("f64", 24, dict(IL, PHASTFT_FACTORS="24:7,10,7", PHASTFT_PASS_C="16,4,16", PHASTFT_PASS_VARIANT=f"0,{v},0"))
    ok &= check("f32", 24, dict(IL, PHASTFT_FACTORS="24:7,10,7", PHASTFT_PASS_C="32,8,32", PHASTFT_PASS_VARIANT=f"0,{v},0"))
print("ALL OK" if ok else "FAILURES", flush=True)

prof("f64", 26, {})
for v in (60, 61, 62, 63, 64, 65, 66, 32):
    prof("f64", 26, dict(IL, PHASTFT_PASS_C="16,4,16", PHASTFT_PASS_VARIANT=f"0,{v},0"))
prof("f64", 25, {})
prof("f64", 25, dict(IL))
prof("f64", 25, dict(IL, PHASTFT_PASS_C="16,4,16"))
prof("f64", 23, {})
prof("f64", 23, dict(IL))
prof("f64", 21, {})
prof("f64", 21, dict(IL))
prof("f32", 26, {})
for v in (0, 60, 61, 62, 63, 64, 65, 66, 32):
    prof("f32", 26, dict(IL, PHASTFT_PASS_C="32,8,32", PHASTFT_PASS_VARIANT=f"0,{v},0"))
prof("f32", 25, {})
prof("f32", 25, dict(IL))
prof("f32", 25, dict(IL, PHASTFT_PASS_C="32,8,32"))
prof("f32", 22, {})
prof("f32", 22, dict(IL))


def tp(sfx, n, env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_VARIANT", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_WS_IL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
    print(f"{sfx} 2^{n} {env}: {ms*1e3:.2f} us | {d[:200]}", flush=True)


for v in ("0", "60", "61", "62", "63", "32"):
    tp("f32", 20, {"PHASTFT_VARIANT": v, "PHASTFT_PASS_C": "8,8"})
    tp("f32", 19, {"PHASTFT_VARIANT": v, "PHASTFT_PASS_C": "8,8"})
for v in ("0", "32"):
    tp("f64", 20, {"PHASTFT_VARIANT": v, "PHASTFT_PASS_C": "8,8"})
    tp("f64", 19, {"PHASTFT_VARIANT": v})
