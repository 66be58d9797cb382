#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the B200 FFT hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPUs

Metric (BASELINE.json): Gpoint/s (complex) for a 2^20-point f64 forward FFT, planar re/im.
A "step" is one in-place forward FFT of one 2^20-point signal through
phastft_fft_dit_f64_dev (device-resident, planner built outside the timed region).  Steps
rotate over NBUF distinct signals whose total size exceeds the 126 MB L2, so every step
reads its input from HBM.  With N GPUs every rank transforms its own stream of signals
(the batch of independent transforms is sharded over the ranks, no data-path collective;
the only collective is the init-time broadcast of the planner tables): weak scaling, and
`value` is the whole-job aggregate = N * 2^20 * K / max-over-ranks time.

`--workload batch_f32` measures BASELINE.json's configs[3] instead (4096 x 2^16 f32 forward,
batch sharded over the ranks, strong scaling); `--workload c2c_f64_2p26` and `r2c_f64_2p24`
measure configs[2] and configs[4] on one GPU.

Only this file's cpu_baseline / --impl reference legs touch oracle/ (the CPU restatement of the
reference): there it is the thing timed as the CPU baseline, never part of our arm.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gpoint/s (complex) for 2^20 f64 forward FFT"
L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------------------
def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline = the oracle (C++ restatement of the reference; the Rust reference cannot be built
# in this image: no cargo/rustc, dependencies not vendored -- see DESIGN.md)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(workload: str, steps: int, warmup: int, budget_s: float):
    """Time the reference algorithm on the host cores with the reference's big-N protocol
    (examples/benchmark.rs:42-63: planner reused, fresh random unit-norm signal per iteration,
    wall clock around exactly one in-place FFT, median).  Returns (value, unit, info dict)."""
    from oracle import oracle as O
    # all host threads the process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores that)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    O.set_threads(threads)
    if workload == "batch_f32":
        n, dt = 1 << 16, np.float32
    elif workload == "c2c_f64_2p26":
        n, dt = 1 << 26, np.float64
    elif workload == "r2c_f64_2p24":
        n, dt = 1 << 24, np.float64
    else:
        n, dt = 1 << 20, np.float64
    rng = np.random.default_rng(1234)
    results = {}
    t_start = time.perf_counter()
    if workload == "r2c_f64_2p24":
        pl = O.PlannerR2c(n, dt)
        x = rng.uniform(-1, 1, n)
        ore = np.zeros(n // 2 + 1); oim = np.zeros(n // 2 + 1)
        for par in (True, False):
            ts = []
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                O.r2c_fft(x, ore, oim, pl, parallel=par)
                dtm = time.perf_counter() - t0
                if it >= warmup:
                    ts.append(dtm)
                if time.perf_counter() - t_start > budget_s and len(ts) >= 3:
                    break
            results[par] = ts
    else:
        planner = O.PlannerDit(n, dt)
        # the reference's fork-join threading (rayon::join) does not scale to every core count -- the
        # upper stages are single-threaded and spinning idle workers cost on a shared host -- so give it
        # its best shot: time it with several thread counts and with one thread, keep the fastest median
        counts = sorted({c for c in (8, 16, 32, threads) if c <= threads}, reverse=True)
        modes = [(True, c) for c in counts] + [(False, 1)]
        for par, cnt in modes:
            O.set_threads(cnt)
            ts = []
            t_mode = time.perf_counter()
            for it in range(warmup + steps):
                re, im = O.gen_random_signal(n, dt, seed=1234 + it)
                t0 = time.perf_counter()
                O.fft_dit(re, im, O.FORWARD, planner, parallel=par)
                dtm = time.perf_counter() - t0
                if it >= warmup:
                    ts.append(dtm)
                if time.perf_counter() - t_mode > budget_s / len(modes) and len(ts) >= 3:
                    break
            results[(par, cnt)] = ts
        O.set_threads(threads)
        med = {k: statistics.median(ts) for k, ts in results.items() if ts}
        best = min(med, key=med.get)
        t = med[best]
        value = n / t / 1e9
        info = {
            "value": value, "unit": "Gpoint/s", "cores": best[1], "kind": "port",
            "sample": f"{len(results[best])} x one 2^{n.bit_length() - 1}-point {np.dtype(dt).name} forward c2c FFT, planner reused, "
                      f"fresh unit-norm signal per iteration, median; PhastFT-restatement (C++), "
                      f"{'fork-join on ' + str(best[1]) + ' threads' if best[0] else 'single thread'} (fastest of "
                      f"{[('fork-join x' + str(c)) if p_ else 'single' for p_, c in modes]})",
            "ms_per_transform": t * 1e3,
            "ms_by_mode": {(("fork_join_x" + str(c)) if p_ else "single"): round(m * 1e3, 3) for (p_, c), m in med.items()},
            "host_threads": threads,
        }
        return value, "Gpoint/s", info
    med = {par: statistics.median(ts) for par, ts in results.items() if ts}
    best_par = min(med, key=med.get)
    t = med[best_par]
    value = n / t / 1e9
    info = {
        "value": value, "unit": "Gpoint/s", "cores": threads if best_par else 1, "kind": "port",
        "sample": f"{len(results[best_par])} x one 2^{n.bit_length() - 1}-point {np.dtype(dt).name} "
                  f"{'r2c' if workload == 'r2c_f64_2p24' else 'forward c2c'} FFT, planner reused, median; "
                  f"PhastFT-restatement (C++), {'fork-join on ' + str(threads) + ' threads' if best_par else 'single thread (faster than fork-join here)'}",
        "ms_per_transform": t * 1e3,
        "ms_single_thread": med.get(False, float("nan")) * 1e3,
        "ms_all_threads": med.get(True, float("nan")) * 1e3,
        "host_threads": threads,
    }
    return value, "Gpoint/s", info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    value, unit, info = cpu_reference_run(args.workload, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC if args.workload == "c2c_f64_2p20" else f"Gpoint/s ({args.workload})",
        "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": info["ms_per_transform"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.workload == "batch_f32" else "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


def planner_passes(planner) -> int:
    d = planner.describe().split(" || ")[0]
    return 1 if "[fused launch" in d else d.count(" | ") + 1


def workload_config(workload: str, gpus: int):
    if workload == "batch_f32":
        return {"workload": "batch of 4096 x 2^16-point f32 forward FFTs, planar re/im, sharded over ranks (BASELINE.json configs[3])",
                "l2": "input 2 GiB per pass > L2", "parallelism": f"batch sharded x{gpus}, no data-path collective"}
    if workload == "c2c_f64_2p26":
        return {"workload": "single 2^26-point f64 forward FFT, planar re/im (BASELINE.json configs[2])", "l2": "input 1 GiB > L2"}
    if workload == "r2c_f64_2p24":
        return {"workload": "r2c_fft_f64 2^24 real input + c2r round trip (BASELINE.json configs[4])", "l2": "input 128 MiB > L2"}
    return {"workload": "single 2^20-point f64 forward FFT, planar re/im, 1 stream per GPU (BASELINE.json configs[1])",
            "l2": "steps rotate over 16 distinct 16 MiB signals (256 MiB > 126 MB L2): every step reads HBM",
            "parallelism": f"independent transforms sharded x{gpus} (one stream of signals per rank), planner tables broadcast once at init"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import __graft_entry__ as ge
    with contextlib.redirect_stdout(sys.stderr):      # stdout carries exactly one JSON line
        ge.build()
    import phastft_b200 as pf
    from phastft_b200 import _lib
    from phastft_b200.sharding import max_over_ranks, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        return max_over_ranks(x, device=dev) if dist is not None else x

    hbm_peak, peak_src = peaks()
    stream = torch.cuda.current_stream(dev)

    def cur_stream():
        # re-read every call: inside a CUDA-graph capture torch's current stream is the capture stream
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    K, W = args.steps, max(args.warmup, 3)
    wl = args.workload
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    extra = {}

    if wl in ("c2c_f64_2p20", "c2c_f64_2p26"):
        n = 1 << (20 if wl == "c2c_f64_2p20" else 26)
        planner = pf.PlannerDit64(n, local)
        if dist is not None:
            planner.broadcast_tables(src=0)          # the one collective of the whole job
        nbuf = 16 if n == (1 << 20) else 2
        nbuf = max(nbuf, -(-(K + W) // 90)) if n == (1 << 20) else nbuf
        bufs_re = [(torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1) for _ in range(nbuf)]
        bufs_im = [(torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1) for _ in range(nbuf)]
        keep = [(b.clone(), c.clone()) for b, c in zip(bufs_re[:2], bufs_im[:2])] if n == (1 << 20) else None
        f = _lib.fn("phastft_fft_dit_{s}_dev", "f64")

        def step(i):
            b = i % nbuf
            _lib.check(f(planner._h, C.c_void_p(bufs_re[b].data_ptr()), C.c_void_p(bufs_im[b].data_ptr()), 1, 1, n, cur_stream()))

        def reset():
            for b in range(nbuf):
                bufs_re[b].uniform_(-1, 1, generator=gen); bufs_im[b].uniform_(-1, 1, generator=gen)
        points_per_step = n
        alg_bytes_per_pass = 2 * n * 8 * 2            # each planar array read once + written once
        dtype = "f64"
        prof = ("f64", planner, lambda b: (bufs_re[b % nbuf], bufs_im[b % nbuf]), 1, n)
    elif wl == "batch_f32":
        n, total = 1 << 16, 4096
        lo, hi = shard_range(total, rank, world)
        nb = hi - lo
        planner = pf.PlannerDit32(n, local)
        if dist is not None:
            planner.broadcast_tables(src=0)
        re = torch.rand(nb * n, dtype=torch.float32, device=dev, generator=gen) * 2 - 1
        im = torch.rand(nb * n, dtype=torch.float32, device=dev, generator=gen) * 2 - 1
        f = _lib.fn("phastft_fft_dit_{s}_dev", "f32")

        def step(i):
            _lib.check(f(planner._h, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1, nb, n, cur_stream()))

        def reset():
            re.uniform_(-1, 1, generator=gen); im.uniform_(-1, 1, generator=gen)
        points_per_step = nb * n
        dtype = "f32"
        prof = ("f32", planner, lambda b: (re, im), nb, n)
        alg_bytes_per_pass = None
    elif wl == "r2c_f64_2p24":
        n = 1 << 24
        planner = pf.PlannerR2c64(n, local)
        x = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        y = torch.empty_like(x)
        sre = torch.empty(n // 2 + 1, dtype=torch.float64, device=dev); sim = torch.empty_like(sre)
        scr_re = torch.empty(n // 2, dtype=torch.float64, device=dev); scr_im = torch.empty_like(scr_re)

        def step(i):
            pf.r2c_fft_f64_with_planner(x, sre, sim, planner)
            pf.c2r_fft_f64_with_planner_and_scratch(sre, sim, y, planner, scr_re, scr_im)

        def reset():
            pass
        points_per_step = n
        dtype = "f64"
        prof = None
        alg_bytes_per_pass = None
    else:
        raise SystemExit(f"unknown workload {wl}")

    # ---- warm-up, then EXACTLY K timed steps between barrier+synchronize -------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    # The 2^20 step is ~10 us of GPU work issued by one C-ABI call (2 kernel launches): capture the
    # rotation over the NBUF signals into a CUDA graph so the host launch path cannot be the bottleneck.
    graph, group = None, 1
    if wl == "c2c_f64_2p20" and not args.no_graph:
        group = nbuf
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(group):
                step(i)
        torch.cuda.synchronize()
    # extra untimed load (~0.3 s) so the clocks are at their loaded level when timing starts
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.3:
        if graph is not None:
            graph.replay()
        else:
            step(0)
        torch.cuda.synchronize()
    reset()
    barrier()
    launches0 = pf.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    if graph is not None:
        for _ in range(K // group):
            graph.replay()
        for i in range(K % group):
            step(i)
    else:
        for i in range(K):
            step(i)
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = pf.launch_count() - launches0
    if graph is not None:
        launches += (K // group) * group * planner_passes(planner)   # graph replays launch kernels without entering the library
    ms_total = maxr(ev0.elapsed_time(ev1))
    barrier()
    ms_per_step = ms_total / K
    value = world * points_per_step / (ms_per_step * 1e-3) / 1e9 if wl != "batch_f32" else 4096 * (1 << 16) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: per-pass CUDA-event times, live, on the launching stream ----
    roofline = None
    if prof is not None:
        sfx, pl_, get, batch_, n_ = prof
        fprof = _lib.fn("phastft_fft_dit_{s}_dev_profile", sfx)
        pass_ms = (C.c_float * 3)()
        npass = C.c_int(0)
        acc = None
        reps = 40 if wl != "c2c_f64_2p26" else 10
        reset()
        # extra load so the clocks sampler sees a loaded GPU for >= ~1 s in total
        for i in range(reps):
            a, b = get(i)
            _lib.check(fprof(pl_._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), 1, batch_, n_, cur_stream(), pass_ms, C.byref(npass)))
            cur = [pass_ms[j] for j in range(npass.value)]
            acc = cur if acc is None else [x + y for x, y in zip(acc, cur)]
        avg_alone = [x / reps for x in acc]
        # Per-launch duration over the TIMED REGION: the step time apportioned by each pass's share of the
        # per-pass CUDA-event times (events around a lone launch also see its launch latency, which the
        # back-to-back launches of the timed region overlap).
        tot_alone = sum(avg_alone) or 1.0
        avg = [ms_per_step * x / tot_alone for x in avg_alone]
        dom = max(range(len(avg)), key=lambda j: avg[j])
        esz = 8 if sfx == "f64" else 4
        # algorithmic bytes of ONE launch: it reads each planar array of its chunk once and writes it once
        chunk = batch_
        if batch_ > 1:
            chunk = max(1, min(batch_, (4 << 30) // (n_ * 2 * esz)))
        bytes_launch = 2 * n_ * esz * 2 * chunk
        ach = bytes_launch / (avg[dom] * 1e-3) / 1e9
        traffic = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text()).get(wl)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel": f"pass {dom + 1}/{len(avg)} of {pl_.describe()}",
                    "algorithmic_bytes_per_launch": bytes_launch, "pass_ms": avg, "pass_ms_timed_alone": avg_alone,
                    "note": "achieved = algorithmic bytes of the dominant pass / its duration in the timed region (step time x the pass's share of the per-pass CUDA-event times); "
                            "a k-pass plan moves k x the compulsory 32N bytes, so the whole-transform fraction is value-based: "
                            f"{(world and 1) * 2 * n_ * esz * 2 * (batch_ if batch_ > 1 else 1) / (ms_per_step * 1e-3) / 1e9 / hbm_peak:.3f}"}
    clocks = sampler.stop()

    # ---- e2e: the reference-facing host-slice call, pinned host buffers, H2D + D2H inside the timed region ----
    e2e = None
    if wl in ("c2c_f64_2p20", "c2c_f64_2p26", "batch_f32"):
        if wl == "batch_f32":
            lo, hi = shard_range(4096, rank, world)
            nb = hi - lo
            h_re = torch.empty(nb * n, dtype=torch.float32).pin_memory(); h_im = torch.empty_like(h_re).pin_memory()
            h_re.uniform_(-1, 1); h_im.uniform_(-1, 1)
            a_re, a_im = h_re.numpy(), h_im.numpy()
            fh = _lib.fn("phastft_fft_dit_{s}_batch_sharded_host", "f32")
            arr = (C.c_void_p * 1)(planner._h)

            def host_step():
                _lib.check(fh(arr, 1, a_re.ctypes.data_as(C.c_void_p), a_im.ctypes.data_as(C.c_void_p), nb, n, 1))
            bytes_one_way = 2 * nb * n * 4
            e_points = 4096 * n
        else:
            h_re = torch.empty(n, dtype=torch.float64).pin_memory(); h_im = torch.empty_like(h_re).pin_memory()
            h_re.uniform_(-1, 1); h_im.uniform_(-1, 1)
            a_re, a_im = h_re.numpy(), h_im.numpy()
            pristine = (a_re.copy(), a_im.copy())

            def host_step():
                pf.fft_64_dit_with_planner(a_re, a_im, pf.Direction.Forward, planner)
            bytes_one_way = 2 * n * 8
            e_points = world * n
        ke = max(5, min(K, 30))
        for _ in range(3):
            host_step()
        if wl != "batch_f32":
            a_re[:] = pristine[0]; a_im[:] = pristine[1]
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            host_step()                      # synchronous: returns after the D2H copy has landed
        t_e = maxr(time.perf_counter() - t0)
        barrier()
        e2e = {"value": e_points / (t_e / ke) / 1e9, "unit": "Gpoint/s", "h2d_bytes_per_step": bytes_one_way,
               "d2h_bytes_per_step": bytes_one_way, "ms_per_step": t_e / ke * 1e3, "steps": ke,
               "api": ("phastft_fft_dit_f32_batch_sharded_host (one call per step: the whole shard, 3-slot H2D/FFT/D2H pipeline), pinned host memory"
                       if wl == "batch_f32" else
                       "phastft_fft_dit_f64_host (= fft_64_dit_with_planner on host slices; one synchronous call per step), pinned host memory")}
        if wl == "c2c_f64_2p20":
            # the same job -- a stream of independent 2^20 transforms in host memory -- handed to the library as ONE
            # batched call, so H2D of signal j+1, the FFT of j and D2H of j-1 overlap (both PCIe directions busy)
            nsig = 32
            s_re = torch.empty(nsig * n, dtype=torch.float64).pin_memory(); s_im = torch.empty_like(s_re).pin_memory()
            s_re.uniform_(-1, 1); s_im.uniform_(-1, 1)
            b_re, b_im = s_re.numpy(), s_im.numpy()
            fhb = _lib.fn("phastft_fft_dit_{s}_batch_sharded_host", "f64")
            arr1 = (C.c_void_p * 1)(planner._h)

            def stream_step():
                _lib.check(fhb(arr1, 1, b_re.ctypes.data_as(C.c_void_p), b_im.ctypes.data_as(C.c_void_p), nsig, n, 1))
            stream_step()
            s_re.uniform_(-1, 1); s_im.uniform_(-1, 1)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                stream_step()
            t_s = maxr(time.perf_counter() - t0)
            barrier()
            e2e["pipelined"] = {"value": world * 3 * nsig * n / t_s / 1e9, "unit": "Gpoint/s", "transforms_per_call": nsig,
                                "ms_per_transform": t_s / (3 * nsig) * 1e3,
                                "api": "phastft_fft_dit_f64_batch_sharded_host: 32 host-resident 2^20 signals per call, 3-slot H2D/FFT/D2H pipeline"}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu = cpu_reference_run(wl, steps=30, warmup=2, budget_s=25.0)

    if rank == 0:
        line = {
            "metric": METRIC if wl == "c2c_f64_2p20" else f"Gpoint/s ({wl})", "value": value, "unit": "Gpoint/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if wl == "batch_f32" else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(wl, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "plan": planner.describe() if hasattr(planner, "describe") else None,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="c2c_f64_2p20", choices=["c2c_f64_2p20", "c2c_f64_2p26", "batch_f32", "r2c_f64_2p24"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every step through the C ABI instead of replaying a captured CUDA graph")
    args = ap.parse_args()
    # defaults sized so the timed region lasts ~0.2-0.5 s (clock sampling needs that) and a run takes < 2 min
    dflt = {"c2c_f64_2p20": (20000, 200), "c2c_f64_2p26": (100, 5), "batch_f32": (100, 5), "r2c_f64_2p24": (200, 10)}[args.workload]
    if args.impl == "reference":
        dflt = {"c2c_f64_2p20": (100, 5), "c2c_f64_2p26": (5, 1), "batch_f32": (500, 20), "r2c_f64_2p24": (10, 1)}[args.workload]
    if args.steps is None:
        args.steps = dflt[0]
    if args.warmup is None:
        args.warmup = dflt[1]
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
