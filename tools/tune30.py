#!/usr/bin/env python
"""Lone transforms of 2^11 / 2^12 points: two launches vs one CTA with the newer three-stage builds."""
import os, sys
sys.path.insert(0, "tools")
from tune import time_plan
KEYS = ("PHASTFT_FACTORS", "PHASTFT_VARIANT", "PHASTFT_PASS_C")
for sfx in ("f64", "f32"):
    for ln in (10, 11, 12, 13):
        envs = [{}]
        for v in ("0", "70", "81"):
            envs.append({"PHASTFT_FACTORS": f"{ln}:{ln}", "PHASTFT_VARIANT": v})
        if ln >= 11:
            envs.append({"PHASTFT_FACTORS": f"{ln}:{ln//2},{ln-ln//2}"})
            envs.append({"PHASTFT_FACTORS": f"{ln}:{ln-ln//2},{ln//2}"})
        for env in envs:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            try:
                ms, d = time_plan(sfx, 1 << ln, 1, 16, 200)
                print(f"{sfx} 2^{ln} {env}: {ms*1e3:.2f} us | {d[:110]}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{sfx} 2^{ln} {env}: FAILED {str(e)[:80]}", flush=True)
