#!/usr/bin/env python
"""Run a few transforms of one plan (for ncu): python tools/one_fft.py f64 24 3   (env vars select the plan)"""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf
from phastft_b200 import _lib
sfx, n_log, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
n = 1 << n_log
dev = torch.device("cuda", 0)
pl = (pf.PlannerDit64 if sfx == "f64" else pf.PlannerDit32)(n, 0)
dt = torch.float64 if sfx == "f64" else torch.float32
re = torch.rand(batch * n, dtype=dt, device=dev); im = torch.rand(batch * n, dtype=dt, device=dev)
f = _lib.fn("phastft_fft_dit_{s}_dev", sfx)
for r in range(reps):
    re.uniform_(-1, 1); im.uniform_(-1, 1)
    _lib.check(f(pl._h, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1, batch, n, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
torch.cuda.synchronize()
print(pl.describe())
