#!/usr/bin/env python
"""GPU tuning sweep: time alternative pass decompositions / tile widths per transform size.
Usage (on the GPU box): python tools/tune.py [f64_20|f64_26|f32_16b|all] > gpurun_out/tune.txt"""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf  # noqa: E402
from phastft_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)


def time_plan(sfx, n, batch, nbuf, reps, graph=True):
    P = pf.PlannerDit64 if sfx == "f64" else pf.PlannerDit32
    dt = torch.float64 if sfx == "f64" else torch.float32
    planner = P(n, 0)
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


def sweep(name, sfx, n_log, batch, nbuf, reps, factor_sets, cs):
    n = 1 << n_log
    esz = 8 if sfx == "f64" else 4
    rows = []
    for fs in factor_sets:
        for c in cs:
            os.environ["PHASTFT_FACTORS"] = f"{n_log}:{','.join(map(str, fs))}"
            os.environ["PHASTFT_TILE_C"] = str(c)
            try:
                ms, desc = time_plan(sfx, n, batch, nbuf, reps)
            except Exception as e:  # noqa: BLE001
                print(f"{name} factors={fs} C={c}: FAILED {e}", flush=True)
                continue
            us = ms * 1e3
            gpts = batch * n / (ms * 1e-3) / 1e9
            gbs = batch * n * esz * 2 * 2 / (ms * 1e-3) / 1e9
            rows.append((us, fs, c, gpts, gbs, desc))
            print(f"{name} factors={fs} C={c}: {us:9.2f} us  {gpts:7.1f} Gpt/s  alg {gbs:7.0f} GB/s | {desc}", flush=True)
    rows.sort()
    print(f"== best {name}: {rows[0][1]} C={rows[0][2]} {rows[0][0]:.2f} us {rows[0][3]:.1f} Gpt/s", flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("f64_20", "all"):
        sweep("f64 2^20", "f64", 20, 1, 16, 50,
              [(10, 10), (7, 7, 6), (6, 7, 7), (7, 6, 7), (8, 6, 6), (6, 6, 8), (6, 8, 6), (5, 7, 8), (8, 7, 5), (9, 6, 5), (5, 6, 9), (5, 5, 10), (10, 5, 5)],
              [4, 8, 16])
    if what in ("f64_26", "all"):
        sweep("f64 2^26", "f64", 26, 1, 1, 3,
              [(9, 9, 8), (8, 9, 9), (9, 8, 9), (10, 8, 8), (8, 8, 10), (8, 10, 8), (10, 10, 6), (6, 10, 10), (10, 9, 7), (7, 9, 10)],
              [4, 8, 16])
    if what in ("f32_16b", "all"):
        sweep("f32 4096x2^16", "f32", 16, 4096, 1, 3,
              [(8, 8), (9, 7), (7, 9), (10, 6), (6, 10), (6, 5, 5), (5, 5, 6), (5, 6, 5)],
              [8, 16, 32])
    if what in ("f64_23", "all"):
        sweep("f64 2^23", "f64", 23, 1, 2, 5, [(8, 8, 7), (7, 8, 8), (8, 7, 8), (9, 7, 7), (7, 7, 9), (10, 7, 6), (6, 7, 10)], [4, 8, 16])


if __name__ == "__main__":
    main()
