#!/usr/bin/env python
"""numpy model of the multi-pass GPU algorithm (design validation only; not product code).

Validates, against numpy.fft, the index algebra the CUDA kernels implement:
  * N = R_1 * ... * R_P passes with the "in-place layout" between passes,
  * twiddle-on-load W_L^{k_{p-1} * (r*B + b)} and its per-CTA separable factorisation,
  * digit-reversed placement + in-place mixed-radix DIT inside a tile,
  * the transposed (digit-reversed) store of the last pass.
"""
import itertools
import numpy as np


def digits_rev(r, radices):
    """r = n_1*(R/r_1) + ... + n_S  ->  pos = n_1 + r_1*n_2 + r_1*r_2*n_3 + ..."""
    R = int(np.prod(radices))
    pos = np.zeros_like(r)
    w_in = R
    w_out = 1
    for rad in radices:
        w_in //= rad
        d = (r // w_in) % rad
        pos += d * w_out
        w_out *= rad
    return pos


def tile_fft(tile, radices):
    """tile: [R, C] complex, rows already multiplied by inter-pass twiddles, natural row order.
    Returns [R, C] natural-order DFT along axis 0 using digit-reversed placement + in-place DIT."""
    R, C = tile.shape
    pos = digits_rev(np.arange(R), radices)
    s = np.empty_like(tile)
    s[pos] = tile                      # copy-in places row r at rev(r)
    Ns = 1
    for rad in radices:
        out = np.empty_like(s)
        for g in range(R // (Ns * rad)):
            for m in range(Ns):
                idx = g * Ns * rad + np.arange(rad) * Ns + m
                v = s[idx] * np.exp(-2j * np.pi * m * np.arange(rad) / (Ns * rad))[:, None]
                # rad-point DFT
                k = np.arange(rad)
                D = np.exp(-2j * np.pi * np.outer(k, k) / rad)
                out[idx] = D @ v
        s = out
        Ns *= rad
    return s


def W(L, e):
    return np.exp(-2j * np.pi * (np.asarray(e) % L) / L)


def multipass_fft(x, Rs, C, radices_of):
    N = x.size
    P = len(Rs)
    cur = x.astype(np.complex128).copy()
    for p in range(P - 1):                       # COL passes
        R = Rs[p]
        A = int(np.prod(Rs[:p])) if p else 1
        B = int(np.prod(Rs[p + 1:]))
        nxt = np.empty_like(cur)
        L = (Rs[p - 1] * R * B) if p else None
        for a in range(A):
            kp = a % Rs[p - 1] if p else 0
            for b0 in range(0, B, C):
                r = np.arange(R)
                c = np.arange(C)
                addr = (a * R + r[:, None]) * B + b0 + c[None, :]
                tile = cur[addr]
                if p:
                    # separable: Urow[r] * V[c]
                    Urow = W(L, kp * B * r)
                    V = W(L, kp * (b0 + c))
                    full = W(L, kp * (r[:, None] * B + b0 + c[None, :]))
                    assert np.allclose(Urow[:, None] * V[None, :], full)
                    tile = tile * Urow[:, None] * V[None, :]
                y = tile_fft(tile, radices_of(R))
                nxt[addr] = y                   # same layout: k_p replaces n_p
        cur = nxt
    # last pass: ROW load, TRANS store
    R = Rs[-1]
    A = N // R
    out = np.empty_like(cur)
    if P == 1:
        for a in range(A):
            out[a * R:(a + 1) * R] = tile_fft(cur[a * R:(a + 1) * R, None], radices_of(R))[:, 0]
        return out
    R1 = Rs[0]
    rest_n = A // R1
    L = Rs[-2] * R
    Cc = min(C, R1)
    for rest in range(rest_n):
        # rev(rest): rest = (k_2..k_{P-1}) most-significant-first -> k_2 + R_2*k_3 ...
        rr, wrev, rem = 0, 1, rest
        mids = Rs[1:-1]
        wgt = rest_n
        for Rm in mids:
            wgt //= Rm
            d = (rem // wgt) % Rm
            rr += d * wrev
            wrev *= Rm
        for k0 in range(0, R1, Cc):
            c = np.arange(Cc)
            r = np.arange(R)
            a = (k0 + c) * rest_n + rest           # row ids
            tile = cur[(a[None, :] * R + r[:, None])]
            kp = a % Rs[-2]                        # [C]
            if P == 2:
                # kp = k0 + c : Urow0[r] * X[c][r]
                tw = W(L, k0 * r)[:, None] * W(L, np.outer(r, c))
                assert np.allclose(tw, W(L, kp[None, :] * r[:, None]))
            else:
                assert np.all(kp == kp[0])
                tw = W(L, kp[0] * r)[:, None] * np.ones((1, Cc))
            y = tile_fft(tile * tw, radices_of(R))
            kP = np.arange(R)
            oaddr = (k0 + c[None, :] + R1 * rr) + A * kP[:, None]
            out[oaddr] = y
    return out


def radices_of(R):
    out = []
    while R > 1:
        for rad in (16, 8, 4, 2):
            if R % rad == 0 and (R // rad == 1 or R // rad >= 2):
                # avoid leaving a lone factor of 2 when a better split exists
                out.append(rad)
                R //= rad
                break
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for Rs, C in (([64], 1), ([16, 16], 4), ([8, 32], 8), ([32, 4], 4), ([8, 4, 16], 4), ([16, 8, 8], 8), ([4, 16, 2], 2)):
        N = int(np.prod(Rs))
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        y = multipass_fft(x, Rs, C, radices_of)
        err = np.max(np.abs(y - np.fft.fft(x))) / np.max(np.abs(y))
        print(Rs, C, "rel err", err)
        assert err < 1e-12
    print("ok")
