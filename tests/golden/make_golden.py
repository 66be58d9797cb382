#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/.

The reference (Rust) cannot be built or imported here, and its own tests hold no stored
vectors (unseeded RNG; expectations are closed-form or computed by RustFFT at test time).
These fixtures therefore pin the *mathematical* answer: seeded inputs (the reference's
fixture shapes: ramp 1..=n `lib.rs:298-338`, unit-norm uniform[-1,1) `utilities/src/lib.rs:26-75`)
and their DFTs computed by a direct O(n^2) extended-precision (np.longdouble) DFT that shares
no code with the oracle or the CUDA kernels.

Run: python tests/golden/make_golden.py   (rewrites tests/golden/*.npz)
"""
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def dft_longdouble(re, im):
    n = re.size
    k = np.arange(n, dtype=np.int64)
    xr = re.astype(np.longdouble)
    xi = im.astype(np.longdouble)
    out_r = np.empty(n, np.longdouble)
    out_i = np.empty(n, np.longdouble)
    two_pi = 2 * np.longdouble("3.14159265358979323846264338327950288")
    for kk in range(n):
        idx = (k * kk) % n           # exact integer phase reduction
        ang = two_pi * idx.astype(np.longdouble) / n
        c, s = np.cos(ang), np.sin(ang)
        out_r[kk] = np.sum(xr * c + xi * s)
        out_i[kk] = np.sum(xi * c - xr * s)
    return out_r, out_i


def unit_norm_signal(n, seed):
    rng = np.random.default_rng(seed)
    re = rng.uniform(-1.0, 1.0, n)
    im = rng.uniform(-1.0, 1.0, n)
    s = 1.0 / np.sqrt(np.sum(re * re + im * im))
    return re * s, im * s


def main():
    out = {}
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096):
        re, im = unit_norm_signal(n, 1234)
        fr, fi = dft_longdouble(re, im)
        out[f"rand_{n}_in_re"] = re
        out[f"rand_{n}_in_im"] = im
        out[f"rand_{n}_out_re"] = fr.astype(np.float64)
        out[f"rand_{n}_out_im"] = fi.astype(np.float64)
    for n in (16, 64, 256, 1024):
        ramp = np.arange(1, n + 1, dtype=np.float64)
        fr, fi = dft_longdouble(ramp, ramp)
        out[f"ramp_{n}_out_re"] = fr.astype(np.float64)
        out[f"ramp_{n}_out_im"] = fi.astype(np.float64)
    np.savez_compressed(HERE / "c2c_golden.npz", **out)

    out = {}
    for n in (4, 8, 16, 64, 256, 2048):
        rng = np.random.default_rng(1234)
        x = rng.uniform(-1.0, 1.0, n)
        fr, fi = dft_longdouble(x, np.zeros(n))
        out[f"real_{n}_in"] = x
        out[f"real_{n}_out_re"] = fr.astype(np.float64)[: n // 2 + 1]
        out[f"real_{n}_out_im"] = fi.astype(np.float64)[: n // 2 + 1]
    np.savez_compressed(HERE / "r2c_golden.npz", **out)
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main()
