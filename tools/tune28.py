#!/usr/bin/env python
"""Four-stage 64-register builds of the half-width 1024-row middle tile (32 warps/SM) vs the default (16 warps/SM)."""
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
ok = True
for v in (68, 69, 71, 72):
    ok &= check("f64", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": f"0,{v},0"})
print("ALL OK" if ok else "FAILURES", flush=True)
prof("f64", 26, {})
for v in (68, 69, 71, 72):
    prof("f64", 26, {"PHASTFT_PASS_VARIANT": f"0,{v},0"})
