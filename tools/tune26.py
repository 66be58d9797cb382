#!/usr/bin/env python
"""Why does the 2^26 middle pass take 715 us in bench.py and 513 us in tune25?  Data values vs addresses."""
import ctypes as C, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf
from phastft_b200 import _lib
dev = torch.device("cuda", 0)
n = 1 << 26
f = _lib.fn("phastft_fft_dit_{s}_dev_profile", "f64")


def run(tag, env, reset, fill=None):
    for k in ("PHASTFT_WS_IL", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pl = pf.PlannerDit64(n, 0)
    re = torch.rand(n, dtype=torch.float64, device=dev); im = torch.rand(n, dtype=torch.float64, device=dev)
    ms = (C.c_float * 3)(); npass = C.c_int(0)
    out = []
    for r in range(12):
        if reset:
            re.uniform_(-1, 1); im.uniform_(-1, 1)
        if fill is not None:
            re.fill_(fill); im.fill_(fill)
        torch.cuda.synchronize()
        _lib.check(f(pl._h, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1, 1, n,
                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), ms, C.byref(npass)))
        out.append([round(ms[i] * 1e3) for i in range(3)])
    fin = torch.isfinite(re).float().mean().item()
    print(tag, env, "reset" if reset else "in-place", "fill", fill, "| finite frac %.2f |" % fin, out[2:], flush=True)


run("A", {}, True)
run("B", {}, False)
run("C", {}, False, fill=float("nan"))
run("D", {}, False, fill=0.0)
run("E", {"PHASTFT_WS_IL": "0"}, True)
run("F", {"PHASTFT_WS_IL": "0"}, False)
run("G", {"PHASTFT_WS_IL": "0"}, False, fill=float("nan"))
