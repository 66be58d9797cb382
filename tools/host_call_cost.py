#!/usr/bin/env python
"""One synchronous host-slice call (phastft_fft_dit_f64_host): page-locked vs ordinary pageable host memory."""
import sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf

for ln in (16, 20, 24):
    n = 1 << ln
    pl = pf.PlannerDit64(n)
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            re = torch.empty(n, dtype=torch.float64).pin_memory().numpy(); im = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        else:
            re = np.empty(n); im = np.empty(n)
        re[:] = 0.5; im[:] = 0.25
        for _ in range(3):
            pf.fft_64_dit_with_planner(re, im, pf.Direction.Forward, pl)
        re[:] = 0.5; im[:] = 0.25
        reps = 20 if ln <= 20 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            pf.fft_64_dit_with_planner(re, im, pf.Direction.Forward, pl)
        dt = (time.perf_counter() - t0) / reps
        print(f"2^{ln} f64 {kind:8s}: {dt*1e3:8.3f} ms per call, {2*2*n*8/dt/1e9:6.1f} GB/s over PCIe (both ways), {n/dt/1e9:6.2f} Gpoint/s", flush=True)
