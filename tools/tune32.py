#!/usr/bin/env python
"""Quarter-width 1024-row middle tiles (32 KB, 4-6 CTAs/SM) vs the half-width two-CTA tile (id 62)."""
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
ok = True
for v in (94, 95, 96, 97):
    ok &= check("f64", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_C": "16,2,16", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
    ok &= check("f32", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_C": "32,4,32", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
print("ALL OK" if ok else "FAILURES", flush=True)
prof("f64", 26, {})
for v in (94, 95, 96, 97):
    prof("f64", 26, {"PHASTFT_PASS_C": "16,2,16", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
prof("f32", 26, {})
for v in (94, 95, 96, 97):
    prof("f32", 26, {"PHASTFT_PASS_C": "32,4,32", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
