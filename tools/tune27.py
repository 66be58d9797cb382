#!/usr/bin/env python
import sys
sys.path.insert(0, "tools")
from tune4_lib import prof
from check_variant import check
ok = check("f64", 24, {"PHASTFT_FACTORS": "24:7,10,7"}) & check("f32", 24, {"PHASTFT_FACTORS": "24:7,10,7"})
ok &= check("f32", 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": "0,67,0"})
print("ALL OK" if ok else "FAILURES", flush=True)
prof("f64", 26, {})
prof("f64", 26, {"PHASTFT_PASS_VARIANT": "0,67,0"})
prof("f64", 26, {"PHASTFT_PASS_VARIANT": "0,65,0"})
prof("f64", 26, {"PHASTFT_WS_IL": "0"})
prof("f32", 26, {})
prof("f32", 26, {"PHASTFT_PASS_VARIANT": "0,67,0"})
prof("f32", 26, {"PHASTFT_PASS_VARIANT": "0,32,0"})
prof("f32", 26, {"PHASTFT_WS_IL": "0"})
prof("f64", 25, {})
prof("f64", 24, {})
prof("f64", 22, {})
prof("f32", 24, {})
