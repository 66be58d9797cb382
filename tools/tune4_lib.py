#!/usr/bin/env python
"""Per-pass timing of selected plans (uses the *_dev_profile entry point)."""
import ctypes as C, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf
from phastft_b200 import _lib
dev = torch.device("cuda", 0)


def prof(sfx, n_log, env, batch=1, reps=10):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_ASYNC", "PHASTFT_L2_GROUP_MB", "PHASTFT_VARIANT", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_WS_IL", "PHASTFT_ROW_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    n = 1 << n_log
    P = pf.PlannerDit64 if sfx == "f64" else pf.PlannerDit32
    dt = torch.float64 if sfx == "f64" else torch.float32
    pl = P(n, 0)
    re = torch.rand(batch * n, dtype=dt, device=dev); im = torch.rand(batch * n, dtype=dt, device=dev)
    f = _lib.fn("phastft_fft_dit_{s}_dev_profile", sfx)
    ms = (C.c_float * 3)(); npass = C.c_int(0)
    acc = [0.0, 0.0, 0.0]
    for r in range(reps + 2):
        re.uniform_(-1, 1); im.uniform_(-1, 1)
        torch.cuda.synchronize()
        _lib.check(f(pl._h, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1, batch, n,
                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), ms, C.byref(npass)))
        if r >= 2:
            for i in range(npass.value):
                acc[i] += ms[i] / reps
    esz = 8 if sfx == "f64" else 4
    chunk = batch if batch == 1 else max(1, min(batch, (4 << 30) // (n * 2 * esz)))
    gb = 2 * n * esz * 2 * chunk / 1e9
    parts = "  ".join(f"p{i+1} {acc[i]*1e3:8.1f} us {gb/(acc[i]*1e-3)/1e3:5.2f} TB/s" for i in range(npass.value))
    print(f"{sfx} 2^{n_log} {env}: total {sum(acc)*1e3:8.1f} us | {parts} | {pl.describe()[:230]}", flush=True)


