#!/usr/bin/env python
"""Size sweep (the reference's benches sweep 2^4..2^26, benches/bench.rs): our forward c2c time per size, single
transforms (CUDA-graph replay over rotating buffers) and batches of 2^24 points in total, f64 and f32, with
cuFFT (torch.fft on interleaved complex) beside it as a yardstick.   python tools/sweep.py > gpurun_out/sweep.txt"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan, dev  # noqa: E402


def cufft_time(cdt, n, batch, nbuf, reps):
    xs = [torch.randn(batch, n, dtype=cdt, device=dev) for _ in range(nbuf)]
    for x in xs:
        torch.fft.fft(x, dim=-1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for x in xs:
                torch.fft.fft(x, dim=-1)
    except Exception:  # noqa: BLE001
        g = None
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            if g is not None:
                g.replay()
            else:
                for x in xs:
                    torch.fft.fft(x, dim=-1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (reps * nbuf))
    return best


for sfx, cdt, esz in (("f64", torch.complex128, 8), ("f32", torch.complex64, 4)):
    print(f"# {sfx}: log2N | single: ours us, cuFFT us, ratio | batch of 2^24 points: ours us (TB/s algorithmic per 2 passes), cuFFT us, ratio | plan")
    for ln in range(4, 28):
        n = 1 << ln
        bytes_sig = n * esz * 2
        nbuf = max(2, min(16, (256 << 20) // bytes_sig))
        reps = 40 if ln <= 20 else (8 if ln <= 24 else 3)
        ours, desc = time_plan(sfx, n, 1, nbuf, reps)
        cu = cufft_time(cdt, n, 1, nbuf, reps)
        line = f"{sfx} 2^{ln:2d} | {ours*1e3:9.2f} {cu*1e3:9.2f} {cu/ours:5.2f}x"
        if ln <= 22:
            batch = (1 << 24) >> ln
            ob, _ = time_plan(sfx, n, batch, 3, 5)
            cb = cufft_time(cdt, n, batch, 3, 5)
            tb = (1 << 24) * esz * 2 * 2 / (ob * 1e-3) / 1e12
            line += f" | {ob*1e3:9.1f} ({tb:4.2f} TB/s/pass-pair) {cb*1e3:9.1f} {cb/ob:5.2f}x"
        print(line + " | " + desc[:150], flush=True)
