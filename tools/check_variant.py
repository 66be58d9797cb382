#!/usr/bin/env python
"""Correctness of an env-selected plan against torch.fft (cuFFT): prints relative L-inf error."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import phastft_b200 as pf

dev = torch.device("cuda", 0)
KEYS = ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_VARIANT", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_WS_IL", "PHASTFT_ROW_VARIANT")


def check(sfx, n_log, env, batch=1):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    n = 1 << n_log
    dt = torch.float64 if sfx == "f64" else torch.float32
    P = pf.PlannerDit64 if sfx == "f64" else pf.PlannerDit32
    pl = P(n, 0)
    g = torch.Generator(device=dev).manual_seed(n_log)
    re = torch.rand(batch, n, dtype=dt, device=dev, generator=g) - 0.5
    im = torch.rand(batch, n, dtype=dt, device=dev, generator=g) - 0.5
    want = torch.fft.fft(torch.complex(re.double(), im.double()), dim=-1)
    fn = pf.fft_64_dit_with_planner if sfx == "f64" else pf.fft_32_dit_with_planner
    if batch == 1:
        r, i = re[0].clone(), im[0].clone()
        fn(r, i, pf.Direction.Forward, pl)
        got = torch.complex(r.double(), i.double())[None]
    else:
        r, i = re.clone().reshape(-1), im.clone().reshape(-1)
        pf.fft_dit_batch(r, i, pf.Direction.Forward, pl, batch)
        got = torch.complex(r.double(), i.double()).reshape(batch, n)
    err = ((got - want).abs().max() / want.abs().max()).item()
    eps = 2.0 ** -52 if sfx == "f64" else 2.0 ** -23
    ok = err <= 4 * eps * n_log
    print(f"{'OK ' if ok else 'BAD'} {sfx} 2^{n_log} b={batch} {env}: rel {err:.2e} (tol {4*eps*n_log:.1e}) | {pl.describe()[:160]}", flush=True)
    return ok


if __name__ == "__main__":
    good = True
    for sfx, cs in (("f64", ("4,4", "8,8", "16,16")), ("f32", ("8,8", "16,16", "32,32"))):
        for n in (20, 19, 18):
            for c in cs:
                good &= check(sfx, n, {"PHASTFT_VARIANT": "32", "PHASTFT_PASS_C": c})
        good &= check(sfx, 18, {"PHASTFT_VARIANT": "34", "PHASTFT_PASS_C": cs[1]})
        good &= check(sfx, 24, {"PHASTFT_FACTORS": "24:7,10,7", "PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": f"{cs[2].split(',')[0]},{cs[0].split(',')[0]},{cs[2].split(',')[0]}"})
        good &= check(sfx, 24, {"PHASTFT_FACTORS": "24:8,9,7", "PHASTFT_PASS_VARIANT": "0,32,0", "PHASTFT_PASS_C": f"{cs[2].split(',')[0]},{cs[1].split(',')[0]},{cs[2].split(',')[0]}"})
        good &= check(sfx, 18, {"PHASTFT_VARIANT": "32", "PHASTFT_PASS_C": cs[0]}, batch=8)
    print("ALL OK" if good else "FAILURES")
