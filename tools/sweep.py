#!/usr/bin/env python
"""Size sweep, the GPU side of the reference's criterion benches (benches/bench.rs, benches/realfft.rs, sizes benches/common/mod.rs:33-35).

    python tools/sweep.py > gpurun_out/sweep.txt                      # table: ours vs cuFFT (torch.fft), lone and batched, 2^4..2^27
    python tools/sweep.py --criterion gpurun_out/criterion             # criterion-shaped output for benches/plot_criterion_overlay.py

--criterion writes, for the groups c2c_forward_{f64,f32}, c2c_inverse_{f64,f32}, r2c_{f64,f32}, c2r_{f64,f32} (the names of
benches/common/mod.rs `groups`) and the sizes 2^6..2^24, the files criterion keeps per benchmark id and the reference's overlay
script reads:  <dir>/<group>/<series>/<n>/new/{sample.json (iters, times in ns), benchmark.json (throughput ElementsAndBytes as
benches/common/mod.rs:91-105: N elements, 2 N sizeof(T) bytes complex / N sizeof(T) real), estimates.json}.
Two series: "PhastFT-B200 device" (data resident in HBM, CUDA-event time) and "PhastFT-B200 host" (the reference-shaped call on host
slices, wall clock, PCIe copies included).  Drop the directory into target/criterion next to the CPU series to overlay them.
"""
import argparse
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from timing import time_plan, dev  # noqa: E402
import phastft_b200 as pf  # noqa: E402


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


def table(lo, hi):
    for sfx, cdt, esz in (("f64", torch.complex128, 8), ("f32", torch.complex64, 4)):
        print(f"# {sfx}: log2N | single: ours us, cuFFT us, ratio | batch of 2^24 points: ours us (whole-transform TB/s), cuFFT us, ratio | plan")
        for ln in range(lo, hi + 1):
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
                line += f" | {ob*1e3:9.1f} ({tb:4.2f} TB/s) {cb*1e3:9.1f} {cb/ob:5.2f}x"
            print(line + " | " + desc[:150], flush=True)


# ---- criterion-shaped output ----------------------------------------------------------------------------------------
def write_id(root: Path, group: str, series: str, n: int, elements: int, nbytes: int, per_iter_ns, iters):
    d = root / group / series / str(n) / "new"
    d.mkdir(parents=True, exist_ok=True)
    times = [p * it for p, it in zip(per_iter_ns, iters)]
    (d / "sample.json").write_text(json.dumps({"sampling_mode": "Flat", "iters": [float(i) for i in iters], "times": times}))
    (d / "benchmark.json").write_text(json.dumps({
        "group_id": group, "function_id": series, "value_str": str(n), "throughput": {"ElementsAndBytes": {"elements": elements, "bytes": nbytes}},
        "full_id": f"{group}/{series}/{n}", "directory_name": f"{group}/{series}/{n}", "title": f"{group}/{series}/{n}"}))
    med, mean = statistics.median(per_iter_ns), statistics.fmean(per_iter_ns)
    sd = statistics.pstdev(per_iter_ns) if len(per_iter_ns) > 1 else 0.0

    def est(v):
        return {"confidence_interval": {"confidence_level": 0.95, "lower_bound": v, "upper_bound": v}, "point_estimate": v, "standard_error": 0.0}
    (d / "estimates.json").write_text(json.dumps({"mean": est(mean), "median": est(med), "std_dev": est(sd), "median_abs_dev": est(0.0), "slope": None}))


def device_samples(run, samples=15, target_ms=3.0):
    """per-call device time in ns: `samples` samples, each `iters` back-to-back calls between two CUDA events."""
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    iters = int(max(1, min(2000, target_ms / one)))
    out = []
    for _ in range(samples):
        e0.record()
        for _ in range(iters):
            run()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e6 / iters)
    return out, [iters] * samples


def host_samples(run, samples=10, target_ms=20.0):
    run()
    t0 = time.perf_counter(); run(); one = max((time.perf_counter() - t0) * 1e3, 1e-3)
    iters = int(max(1, min(200, target_ms / one)))
    out = []
    for _ in range(samples):
        t0 = time.perf_counter()
        for _ in range(iters):
            run()
        out.append((time.perf_counter() - t0) * 1e9 / iters)
    return out, [iters] * samples


def criterion(root: Path, lo: int, hi: int, host: bool):
    root.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(0)
    for sfx, tdt, ndt, P, R, fft, r2c, c2r in (
            ("f64", torch.float64, np.float64, pf.PlannerDit64, pf.PlannerR2c64, pf.fft_64_dit_with_planner, pf.r2c_fft_f64_with_planner, pf.c2r_fft_f64_with_planner),
            ("f32", torch.float32, np.float32, pf.PlannerDit32, pf.PlannerR2c32, pf.fft_32_dit_with_planner, pf.r2c_fft_f32_with_planner, pf.c2r_fft_f32_with_planner)):
        esz = np.dtype(ndt).itemsize
        for ln in range(lo, hi + 1):
            n = 1 << ln
            planner = P(n, 0)
            d_re = torch.rand(n, dtype=tdt, device=dev); d_im = torch.rand(n, dtype=tdt, device=dev)
            h_re = torch.rand(n, dtype=tdt).pin_memory(); h_im = torch.rand(n, dtype=tdt).pin_memory()
            a_re, a_im = h_re.numpy(), h_im.numpy()
            for group, direction in ((f"c2c_forward_{sfx}", pf.Direction.Forward), (f"c2c_inverse_{sfx}", pf.Direction.Reverse)):
                p_, it_ = device_samples(lambda: fft(d_re, d_im, direction, planner))
                write_id(root, group, "PhastFT-B200 device", n, n, 2 * n * esz, p_, it_)
                if host:
                    p_, it_ = host_samples(lambda: fft(a_re, a_im, direction, planner))
                    write_id(root, group, "PhastFT-B200 host", n, n, 2 * n * esz, p_, it_)
            if n >= 4:
                rp = R(n, 0)
                x = torch.rand(n, dtype=tdt, device=dev); y = torch.empty_like(x)
                sre = torch.empty(n // 2 + 1, dtype=tdt, device=dev); sim = torch.empty_like(sre)
                p_, it_ = device_samples(lambda: r2c(x, sre, sim, rp))
                write_id(root, f"r2c_{sfx}", "PhastFT-B200 device", n, n, n * esz, p_, it_)
                p_, it_ = device_samples(lambda: c2r(sre, sim, y, rp))
                write_id(root, f"c2r_{sfx}", "PhastFT-B200 device", n, n, n * esz, p_, it_)
                if host:
                    hx = rng.uniform(0, 1, n).astype(ndt); hre = np.zeros(n // 2 + 1, ndt); him = np.zeros(n // 2 + 1, ndt); hy = np.zeros(n, ndt)
                    p_, it_ = host_samples(lambda: r2c(hx, hre, him, rp))
                    write_id(root, f"r2c_{sfx}", "PhastFT-B200 host", n, n, n * esz, p_, it_)
                    p_, it_ = host_samples(lambda: c2r(hre, him, hy, rp))
                    write_id(root, f"c2r_{sfx}", "PhastFT-B200 host", n, n, n * esz, p_, it_)
            print(f"criterion {sfx} 2^{ln}: written", file=sys.stderr, flush=True)
            del planner


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--criterion", type=Path, default=None, help="write criterion-shaped results under this directory instead of the table")
    ap.add_argument("--lo", type=int, default=None)
    ap.add_argument("--hi", type=int, default=None)
    ap.add_argument("--no-host", action="store_true", help="--criterion: only the device-resident series")
    a = ap.parse_args()
    if a.criterion is not None:
        criterion(a.criterion, a.lo if a.lo is not None else 6, a.hi if a.hi is not None else 24, not a.no_host)
    else:
        table(a.lo if a.lo is not None else 4, a.hi if a.hi is not None else 27)
