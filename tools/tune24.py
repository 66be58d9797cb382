#!/usr/bin/env python
"""Interleaved-complex intermediates (PHASTFT_WS_IL=1): 64-byte runs from half as many columns."""
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
from tune import time_plan
import os

IL = {"PHASTFT_WS_IL": "1"}
ok = True
for sfx, c in (("f64", "16,4,16"), ("f32", "32,8,32")):
    ok &= check(sfx, 24, dict(IL))
    ok &= check(sfx, 24, dict(IL, PHASTFT_PASS_C=c))
    ok &= check(sfx, 24, dict(IL, PHASTFT_FACTORS="24:7,10,7", PHASTFT_PASS_C=c, PHASTFT_PASS_VARIANT="0,32,0"))
    ok &= check(sfx, 24, dict(IL, PHASTFT_FACTORS="24:7,10,7", PHASTFT_PASS_C=c))
    ok &= check(sfx, 20, dict(IL))
    ok &= check(sfx, 16, dict(IL), batch=8)
    ok &= check(sfx, 22, dict(IL), batch=3)
print("ALL OK" if ok else "FAILURES", flush=True)

prof("f64", 26, {})
prof("f64", 26, dict(IL))
prof("f64", 26, dict(IL, PHASTFT_PASS_C="16,4,16"))
prof("f64", 26, dict(IL, PHASTFT_PASS_C="16,4,16", PHASTFT_PASS_VARIANT="0,60,0"))
prof("f64", 26, dict(IL, PHASTFT_PASS_C="16,4,16", PHASTFT_PASS_VARIANT="0,32,0"))
prof("f64", 26, dict(IL, PHASTFT_PASS_C="16,8,16", PHASTFT_PASS_VARIANT="0,32,0"))
prof("f64", 26, dict(IL, PHASTFT_FACTORS="26:8,9,9", PHASTFT_PASS_C="16,4,8"))
prof("f64", 26, dict(IL, PHASTFT_FACTORS="26:9,9,8", PHASTFT_PASS_C="16,4,16"))
prof("f64", 24, {})
prof("f64", 24, dict(IL))
prof("f64", 24, dict(IL, PHASTFT_PASS_C="16,4,16"))
prof("f64", 22, {})
prof("f64", 22, dict(IL))
prof("f64", 22, dict(IL, PHASTFT_PASS_C="16,4,16"))
prof("f32", 26, {})
prof("f32", 26, dict(IL))
prof("f32", 26, dict(IL, PHASTFT_PASS_C="32,8,32"))
prof("f32", 26, dict(IL, PHASTFT_PASS_C="32,8,32", PHASTFT_PASS_VARIANT="0,32,0"))
prof("f32", 26, dict(IL, PHASTFT_PASS_C="32,16,32", PHASTFT_PASS_VARIANT="0,32,0"))
prof("f32", 24, {})
prof("f32", 24, dict(IL))
prof("f32", 24, dict(IL, PHASTFT_PASS_C="32,8,32"))
# batch f32 4096 x 2^16
prof("f32", 16, {}, batch=4096)
prof("f32", 16, dict(IL), batch=4096)


def tp(sfx, n, env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_VARIANT", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_WS_IL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
    print(f"{sfx} 2^{n} {env}: {ms*1e3:.2f} us | {d[:200]}", flush=True)


for sfx in ("f64", "f32"):
    for n in (20, 18, 16):
        tp(sfx, n, {})
        tp(sfx, n, dict(IL))
    tp(sfx, 20, dict(IL, PHASTFT_VARIANT="32", PHASTFT_PASS_C="8,8"))
