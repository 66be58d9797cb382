#!/usr/bin/env python
"""Timing helper for the tuning / sweep scripts: device time of one plan on rotating device-resident buffers, CUDA events
on the launching stream, best of 3, optionally replayed from a CUDA graph (launch-bound sizes).

    python tools/timing.py f32 16 4096          # one size: precision, log2 N, batch  (PHASTFT_* env vars select the plan)
"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf  # noqa: E402
from phastft_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)


def time_plan(sfx, n, batch, nbuf, reps, graph=True, mode=None):
    P = pf.PlannerDit64 if sfx == "f64" else pf.PlannerDit32
    dt = torch.float64 if sfx == "f64" else torch.float32
    planner = P(n, 0) if mode is None else P.with_mode(n, mode, 0)
    planner.reserve(batch)
    bufs = [(torch.rand(batch * n, dtype=dt, device=dev) * 2 - 1, torch.rand(batch * n, dtype=dt, device=dev) * 2 - 1) for _ in range(nbuf)]
    f = _lib.fn("phastft_fft_dit_{s}_dev", sfx)

    def step(i):
        a, b = bufs[i % nbuf]
        _lib.check(f(planner._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), 1, batch, n,
                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    for i in range(nbuf):
        step(i)
    torch.cuda.synchronize()
    g = None
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nbuf):
                step(i)
        torch.cuda.synchronize()
    best = 1e9
    for trial in range(3):
        for a, b in bufs:
            a.uniform_(-1, 1); b.uniform_(-1, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            if g is not None:
                g.replay()
            else:
                for i in range(nbuf):
                    step(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (reps * nbuf))
    desc = planner.describe()
    del planner, bufs
    return best, desc


def report(sfx, ln, batch, label=""):
    n = 1 << ln
    esz = 8 if sfx == "f64" else 4
    total = batch * n * esz * 2
    nbuf = max(2, min(16, (512 << 20) // total))
    reps = max(3, min(50, int(2e9 // (total * nbuf))))
    ms, desc = time_plan(sfx, n, batch, nbuf, reps)
    tb = total * 2 / (ms * 1e-3) / 1e12
    print(f"{label:28s} {sfx} 2^{ln:2d} x {batch:6d}: {ms * 1e3:9.2f} us  {batch * n / ms / 1e6:7.1f} Gpoint/s  whole-transform {tb:5.2f} TB/s = "
          f"{tb / 6.5696:4.2f} of measured peak | {desc[:260]}", flush=True)
    return ms


if __name__ == "__main__":
    report(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "")
