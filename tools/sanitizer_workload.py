#!/usr/bin/env python
"""Workload for `compute-sanitizer --tool memcheck|racecheck python tools/sanitizer_workload.py`:
every kernel kind, both precisions, lone transforms and batches, r2c/c2r, checked against numpy."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf
import torch
rng = np.random.default_rng(0)
for dt, P, f in ((np.float64, pf.PlannerDit64, pf.fft_64_dit_with_planner), (np.float32, pf.PlannerDit32, pf.fft_32_dit_with_planner)):
    for n in (1, 2, 8, 64, 256, 1024, 2048, 4096, 1 << 13, 1 << 15, 1 << 16, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
        re = rng.uniform(-1, 1, n).astype(dt); im = rng.uniform(-1, 1, n).astype(dt)
        pl = P(n)
        ref = np.fft.fft(re.astype(np.float64) + 1j * im)
        f(re, im, pf.Direction.Forward, pl)
        err = np.max(np.abs(re + 1j * im - ref)) / max(np.max(np.abs(ref)), 1e-30)
        assert err < (1e-13 if dt == np.float64 else 1e-5), (dt, n, err)
    # small batches, then batches with batch * N >= 2^21 (their own kernels, interleaved intermediates, ragged last CTA)
    for n, b in ((256, 40), (2048, 9), (4096, 5), (1 << 16, 40), (4, (1 << 19) + 3), (8, (1 << 18) + 5), (16, (1 << 17) + 7),
                 (512, 4099), (1024, 2051), (2048, 1027), (4096, 515), (1 << 13, 300), (1 << 14, 131), (1 << 21, 2),
                 (1 << 16, 150), (1 << 18, 36), (1 << 20, 9), (1 << 17, 67)):        # >= 32 MiB per array: the TMA / bulk pair
        pl = P(n)
        tdt = torch.float64 if dt == np.float64 else torch.float32
        d_re = torch.rand(n * b, dtype=tdt, device="cuda"); d_im = torch.rand(n * b, dtype=tdt, device="cuda")
        x = (d_re.cpu().numpy().astype(np.float64) + 1j * d_im.cpu().numpy()).reshape(b, n)
        pf.fft_dit_batch(d_re, d_im, pf.Direction.Forward, pl, b)
        got = (d_re.cpu().numpy().astype(np.float64) + 1j * d_im.cpu().numpy()).reshape(b, n)
        ref = np.fft.fft(x, axis=1)
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-13 if dt == np.float64 else 1e-5), (dt, n, b)
for n in (4, 64, 4096, 1 << 15, 1 << 19, 1 << 22):      # from 2^13 (f64) c2r builds its first pass's input while loading
    x = rng.uniform(-1, 1, n)
    ore = np.zeros(n // 2 + 1); oim = np.zeros(n // 2 + 1)
    pf.r2c_fft_f64(x, ore, oim)
    ref = np.fft.rfft(x)
    assert np.max(np.abs(ore + 1j * oim - ref)) / np.max(np.abs(ref)) < 1e-13
    y = np.zeros(n); pf.c2r_fft_f64(ore, oim, y)
    assert np.max(np.abs(y - x)) < 1e-12
# host-resident batch through the 3-slot copy / compute pipeline (5 chunks, ragged tail, strided), both layouts forced
import os
os.environ["PHASTFT_HOST_CHUNK_MB"] = "1"
n, b, stride = 1 << 14, 37, (1 << 14) + 24
tot = (b - 1) * stride + n
h_re = torch.from_numpy(rng.uniform(-1, 1, tot)).pin_memory(); h_im = torch.from_numpy(rng.uniform(-1, 1, tot)).pin_memory()
a_re, a_im = h_re.numpy(), h_im.numpy()
x = np.stack([a_re[i * stride:i * stride + n] + 1j * a_im[i * stride:i * stride + n] for i in range(b)])
pf.fft_dit_batch_sharded(a_re, a_im, pf.Direction.Forward, [pf.PlannerDit64(n)], b, batch_stride=stride)
got = np.stack([a_re[i * stride:i * stride + n] + 1j * a_im[i * stride:i * stride + n] for i in range(b)])
assert np.max(np.abs(got - np.fft.fft(x, axis=1))) / np.max(np.abs(x)) < 1e-11
for il in ("0", "1"):
    os.environ["PHASTFT_WS_IL"] = il
    for n in (1 << 13, 1 << 21):
        re = rng.uniform(-1, 1, n); im = rng.uniform(-1, 1, n)
        ref = np.fft.fft(re + 1j * im)
        pf.fft_64_dit_with_planner(re, im, pf.Direction.Forward, pf.PlannerDit64(n))
        assert np.max(np.abs(re + 1j * im - ref)) / np.max(np.abs(ref)) < 1e-13
os.environ.pop("PHASTFT_WS_IL")
z = (rng.uniform(-1, 1, 1 << 15) + 1j * rng.uniform(-1, 1, 1 << 15)).astype(np.complex128)
ref = np.fft.ifft(z)
pf.fft_64_interleaved(z, pf.Direction.Reverse)
assert np.max(np.abs(z - ref)) / np.max(np.abs(ref)) < 1e-13
# ---- round 2: thread-block-cluster launch (DSMEM exchange), TMA / bulk tile input, pipelined two-pass launch, 128 KB one-CTA kernels ----
def batch_check(dt, n, b, tolv):
    P = pf.PlannerDit64 if dt == np.float64 else pf.PlannerDit32
    tdt = torch.float64 if dt == np.float64 else torch.float32
    pl = P(n)
    d_re = torch.rand(n * b, dtype=tdt, device="cuda"); d_im = torch.rand(n * b, dtype=tdt, device="cuda")
    x = (d_re.cpu().numpy().astype(np.float64) + 1j * d_im.cpu().numpy()).reshape(b, n)
    pf.fft_dit_batch(d_re, d_im, pf.Direction.Forward, pl, b)
    got = (d_re.cpu().numpy().astype(np.float64) + 1j * d_im.cpu().numpy()).reshape(b, n)
    ref = np.fft.fft(x, axis=1)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < tolv, (dt, n, b, pl.describe())
    return pl.describe()


os.environ["PHASTFT_CLUSTER"] = "1"
os.environ["PHASTFT_ONE_CTA_MAX"] = "14"
for dt, sizes in ((np.float64, (13, 14, 15, 16)), (np.float32, (14, 15, 16))):
    for ln in sizes:
        d = batch_check(dt, 1 << ln, 19, 1e-13 if dt == np.float64 else 1e-5)
        assert "CLUSTER" in d or "batches: ROW" in d, d
os.environ.pop("PHASTFT_CLUSTER"); os.environ.pop("PHASTFT_ONE_CTA_MAX")
os.environ["PHASTFT_PIPE"] = "1"; os.environ["PHASTFT_PIPE_RING_MB"] = "1"
for tma in ("0", "1"):
    os.environ["PHASTFT_PIPE_TMA"] = tma
    for dt, n, b in ((np.float32, 1 << 16, 70), (np.float64, 1 << 16, 40), (np.float64, 1 << 14, 300)):
        assert "pipelined" in batch_check(dt, n, b, 1e-13 if dt == np.float64 else 1e-5)
for k in ("PHASTFT_PIPE", "PHASTFT_PIPE_RING_MB", "PHASTFT_PIPE_TMA"):
    os.environ.pop(k)
os.environ["PHASTFT_TMA"] = "1"
for dt, f, P in ((np.float64, pf.fft_64_dit_with_planner, pf.PlannerDit64), (np.float32, pf.fft_32_dit_with_planner, pf.PlannerDit32)):
    for n in (1 << 18, 1 << 20):
        re = rng.uniform(-1, 1, n).astype(dt); im = rng.uniform(-1, 1, n).astype(dt)
        ref = np.fft.fft(re.astype(np.float64) + 1j * im)
        pl = P(n)
        assert ",tma" in pl.describe()
        f(re, im, pf.Direction.Forward, pl)
        assert np.max(np.abs(re + 1j * im - ref)) / np.max(np.abs(ref)) < (1e-13 if dt == np.float64 else 1e-5)
os.environ.pop("PHASTFT_TMA")
n = 1 << 25          # middle pass by TMA (the default for 2^25+)
re = rng.uniform(-1, 1, n); im = rng.uniform(-1, 1, n)
ref0 = np.sum(re) + 1j * np.sum(im)
pl = pf.PlannerDit64(n)
assert "middle pass by TMA" in pl.describe()
pf.fft_64_dit_with_planner(re, im, pf.Direction.Forward, pl)
assert abs(re[0] + 1j * im[0] - ref0) < 1e-6
print("sanitizer workload ok")
