#!/usr/bin/env python
"""Planner construction cost: wall time of plan create / destroy per size, heuristic and tune
mode, and of the one-shot API (plan per call) on pinned host slices.  (The reference's planner computes one cos/sin
table per stage, 2(N-64) values, planner.rs:68-100; ours builds 2*sqrt(N) + sum(R_p) table entries.)"""
import sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf

torch.zeros(1, device="cuda")
pf.PlannerDit64(1 << 10)          # module load, context
for ln in (4, 10, 16, 20, 24, 26):
    n = 1 << ln
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); p = pf.PlannerDit64(n); torch.cuda.synchronize(); t1 = time.perf_counter()
        del p
        ts.append(t1 - t0)
    t0 = time.perf_counter(); p = pf.PlannerDit64.with_mode(n, pf.PlannerMode.Tune); t_tune = time.perf_counter() - t0
    del p
    line = f"2^{ln:2d} f64: plan create min {min(ts)*1e3:8.3f} ms  median {sorted(ts)[2]*1e3:8.3f} ms | tune mode {t_tune*1e3:9.2f} ms"
    if ln <= 24:
        re = torch.empty(n, dtype=torch.float64).pin_memory().numpy(); im = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        re[:] = 1; im[:] = 0
        pf.fft_64_dit(re, im, pf.Direction.Forward)
        t0 = time.perf_counter()
        for _ in range(3):
            pf.fft_64_dit(re, im, pf.Direction.Forward)
        line += f" | one-shot fft_64_dit (plan + copies + transform) {(time.perf_counter()-t0)/3*1e3:8.3f} ms"
    print(line, flush=True)
