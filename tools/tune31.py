#!/usr/bin/env python
"""Persistent double-buffered middle pass (ids 90-93) vs the two-CTA/SM half-width tile (id 62)."""
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
ok = True
for v in (90, 91, 92, 93):
    ok &= check("f64", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
    ok &= check("f32", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
ok &= check("f64", 26, {"PHASTFT_PASS_VARIANT": "0,90,0"})
ok &= check("f64", 22, {"PHASTFT_FACTORS": "22:6,10,6", "PHASTFT_PASS_VARIANT": "0,91,0"}, batch=3)
print("ALL OK" if ok else "FAILURES", flush=True)
prof("f64", 26, {})
for v in (90, 91, 92, 93):
    prof("f64", 26, {"PHASTFT_PASS_VARIANT": f"0,{v},0"})
prof("f32", 26, {})
for v in (90, 91):
    prof("f32", 26, {"PHASTFT_PASS_VARIANT": f"0,{v},0"})
