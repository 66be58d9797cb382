#!/usr/bin/env python
"""Radix-32 two-stage tiles (variant id 32) vs the default three-stage 1024/512-row tiles."""
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from tune import time_plan
import os

# 2^26 f64: middle pass
prof("f64", 26, {})
prof("f64", 26, {"PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "16,4,16"})
prof("f64", 26, {"PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "16,8,16"})
# 2^24 f64: {7,10,7} with the radix-32 middle vs default {8,8,8}
prof("f64", 24, {})
prof("f64", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "16,4,16"})
prof("f64", 24, {"PHASTFT_FACTORS": "24:8,9,7", "PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "16,8,16"})
prof("f64", 24, {"PHASTFT_FACTORS": "24:8,9,7", "PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "16,4,16"})
prof("f64", 24, {"PHASTFT_FACTORS": "24:8,9,7", "PHASTFT_PASS_VARIANT": "0,34,0", "PHASTFT_PASS_C": "16,8,16"})
# 2^27 / 2^28 f64 with 9-bit ends (128-byte runs need C=16)
prof("f64", 28, {})
prof("f64", 28, {"PHASTFT_FACTORS": "28:9,10,9", "PHASTFT_PASS_VARIANT": "32,32,32", "PHASTFT_PASS_C": "16,4,16"})
# f32 2^26 / 2^28
prof("f32", 26, {})
prof("f32", 26, {"PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "32,8,32"})
prof("f32", 26, {"PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": "32,16,32"})
prof("f32", 28, {})
prof("f32", 28, {"PHASTFT_FACTORS": "28:9,10,9", "PHASTFT_PASS_VARIANT": "32,32,32", "PHASTFT_PASS_C": "32,8,32"})

# L2-resident sizes, graph replay
def tp(sfx, n, env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_VARIANT", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_WS_IL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ms, d = time_plan(sfx, 1 << n, 1, 16, 40)
    print(f"{sfx} 2^{n} {env}: {ms*1e3:.2f} us | {d[:200]}", flush=True)

for sfx, cs in (("f64", ("4,4", "8,8")), ("f32", ("8,8", "16,16"))):
    for n in (20, 19, 18):
        tp(sfx, n, {})
        for c in cs:
            tp(sfx, n, {"PHASTFT_VARIANT": "32", "PHASTFT_PASS_C": c})
