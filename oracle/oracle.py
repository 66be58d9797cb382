"""ctypes loader for the CPU oracle (oracle/phastft_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (phastft_b200/) never imports it.

The functions mirror PhastFT's public API names (lib.rs:143-226, r2c.rs:521-895) and raise
`OraclePanic` with the reference's panic message where the reference would panic.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libphastft_oracle.so"

FORWARD = 1
REVERSE = -1

_MESSAGES = {
    1: "assertion `left == right` failed: reals.len() == imags.len()",
    2: "assertion failed: length must be a non-zero power of two",
    3: "assertion `left == right` failed: log_n == planner.log_n",
    4: "n must be a power of 2 >= 4",
    5: "input length must match planner size",
    6: "output_re must have length N/2 + 1",
    7: "output_im must have length N/2 + 1",
    8: "output length must match planner size",
    9: "input_re must have length N/2 + 1",
    10: "input_im must have length N/2 + 1",
    11: "scratch_re must have length N/2",
    12: "scratch_im must have length N/2",
    13: "invalid argument",
}


class OraclePanic(AssertionError):
    def __init__(self, code: int):
        self.code = code
        super().__init__(_MESSAGES.get(code, f"status {code}"))


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (g++ is in the image)."""
    src = _HERE / "phastft_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(L):
    sz, vp, i32, ci, u = C.c_size_t, C.c_void_p, C.c_int32, C.c_int, C.c_uint
    L.oracle_max_threads.restype = ci
    L.oracle_set_threads.argtypes = [ci]
    for sfx in ("f64", "f32"):
        g = lambda name: getattr(L, name.replace("SFX", sfx))
        g("oracle_plan_dit_SFX_new").argtypes = [sz, C.POINTER(vp)]
        g("oracle_plan_dit_SFX_new").restype = i32
        g("oracle_plan_dit_SFX_free").argtypes = [vp]
        g("oracle_plan_dit_SFX_num_tables").argtypes = [vp]
        g("oracle_plan_dit_SFX_num_tables").restype = sz
        g("oracle_plan_dit_SFX_table").argtypes = [vp, sz, vp, vp]
        g("oracle_fft_dit_SFX_with_planner").argtypes = [vp, sz, vp, sz, ci, vp, ci]
        g("oracle_fft_dit_SFX_with_planner").restype = i32
        g("oracle_fft_dit_SFX").argtypes = [vp, sz, vp, sz, ci, ci]
        g("oracle_fft_dit_SFX").restype = i32
        g("oracle_plan_r2c_SFX_new").argtypes = [sz, C.POINTER(vp)]
        g("oracle_plan_r2c_SFX_new").restype = i32
        g("oracle_plan_r2c_SFX_free").argtypes = [vp]
        g("oracle_plan_r2c_SFX_twiddles").argtypes = [vp, vp, vp]
        g("oracle_r2c_SFX_with_planner").argtypes = [vp, sz, vp, sz, vp, sz, vp, ci]
        g("oracle_r2c_SFX_with_planner").restype = i32
        g("oracle_r2c_SFX").argtypes = [vp, sz, vp, sz, vp, sz, ci]
        g("oracle_r2c_SFX").restype = i32
        g("oracle_c2r_SFX_with_planner_and_scratch").argtypes = [vp, sz, vp, sz, vp, sz, vp, vp, sz, vp, sz, ci]
        g("oracle_c2r_SFX_with_planner_and_scratch").restype = i32
        g("oracle_c2r_SFX_with_planner").argtypes = [vp, sz, vp, sz, vp, sz, vp, ci]
        g("oracle_c2r_SFX_with_planner").restype = i32
        g("oracle_c2r_SFX").argtypes = [vp, sz, vp, sz, vp, sz, ci]
        g("oracle_c2r_SFX").restype = i32
        g("oracle_bit_reverse_SFX").argtypes = [vp, u, ci]
        g("oracle_codelet_SFX").argtypes = [vp, vp, sz]
        g("oracle_codelet_stages_SFX").restype = u
        g("oracle_stage_SFX").argtypes = [vp, vp, sz, u]
        g("oracle_literal_twiddles_SFX").argtypes = [u, vp, vp]


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"unsupported dtype {dtype}")


def _ptr(a: np.ndarray):
    assert a.flags.c_contiguous
    return a.ctypes.data_as(C.c_void_p)


def _check(code: int):
    if code != 0:
        raise OraclePanic(code)


def max_threads() -> int:
    return lib().oracle_max_threads()


def set_threads(t: int) -> None:
    lib().oracle_set_threads(int(t))


class _Plan:
    def __init__(self, kind: str, n: int, dtype):
        self.sfx = _sfx(dtype)
        self.dtype = np.dtype(dtype)
        self.kind = kind
        self.n = int(n)
        h = C.c_void_p()
        _check(getattr(lib(), f"oracle_plan_{kind}_{self.sfx}_new")(self.n, C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            getattr(lib(), f"oracle_plan_{self.kind}_{self.sfx}_free")(h)
            self._h = None


class PlannerDit(_Plan):
    """planner.rs:34-100 (PlannerDit64 / PlannerDit32)."""

    def __init__(self, n: int, dtype=np.float64):
        super().__init__("dit", n, dtype)

    def stage_twiddles(self):
        L = lib()
        out = []
        nt = getattr(L, f"oracle_plan_dit_{self.sfx}_num_tables")(self._h)
        for i in range(nt):
            dist = 64 << i
            re = np.empty(dist, self.dtype)
            im = np.empty(dist, self.dtype)
            getattr(L, f"oracle_plan_dit_{self.sfx}_table")(self._h, i, _ptr(re), _ptr(im))
            out.append((re, im))
        return out


class PlannerR2c(_Plan):
    """planner.rs:164-212 (PlannerR2c64 / PlannerR2c32)."""

    def __init__(self, n: int, dtype=np.float64):
        super().__init__("r2c", n, dtype)

    def twiddles(self):
        re = np.empty(self.n // 2, self.dtype)
        im = np.empty(self.n // 2, self.dtype)
        getattr(lib(), f"oracle_plan_r2c_{self.sfx}_twiddles")(self._h, _ptr(re), _ptr(im))
        return re, im


def fft_dit(re: np.ndarray, im: np.ndarray, direction: int, planner: PlannerDit | None = None,
            parallel: bool = False) -> None:
    """fft_64_dit / fft_32_dit [_with_planner] (lib.rs:143-226): in place on planar arrays."""
    sfx = _sfx(re.dtype)
    assert im.dtype == re.dtype
    L = lib()
    if planner is None:
        _check(getattr(L, f"oracle_fft_dit_{sfx}")(_ptr(re), re.size, _ptr(im), im.size, direction, int(parallel)))
    else:
        assert planner.sfx == sfx
        _check(getattr(L, f"oracle_fft_dit_{sfx}_with_planner")(_ptr(re), re.size, _ptr(im), im.size, direction,
                                                                 planner._h, int(parallel)))


def r2c_fft(x: np.ndarray, out_re: np.ndarray, out_im: np.ndarray, planner: PlannerR2c | None = None,
            parallel: bool = False) -> None:
    """r2c_fft_f64 / r2c_fft_f32 [_with_planner] (r2c.rs:521-662)."""
    sfx = _sfx(x.dtype)
    L = lib()
    if planner is None:
        _check(getattr(L, f"oracle_r2c_{sfx}")(_ptr(x), x.size, _ptr(out_re), out_re.size, _ptr(out_im), out_im.size,
                                                int(parallel)))
    else:
        _check(getattr(L, f"oracle_r2c_{sfx}_with_planner")(_ptr(x), x.size, _ptr(out_re), out_re.size, _ptr(out_im),
                                                             out_im.size, planner._h, int(parallel)))


def c2r_fft(in_re: np.ndarray, in_im: np.ndarray, out: np.ndarray, planner: PlannerR2c | None = None,
            scratch_re: np.ndarray | None = None, scratch_im: np.ndarray | None = None,
            parallel: bool = False) -> None:
    """c2r_fft_f64 / f32 [_with_planner[_and_scratch]] (r2c.rs:695-895)."""
    sfx = _sfx(out.dtype)
    L = lib()
    if planner is None:
        _check(getattr(L, f"oracle_c2r_{sfx}")(_ptr(in_re), in_re.size, _ptr(in_im), in_im.size, _ptr(out), out.size,
                                                int(parallel)))
    elif scratch_re is None:
        _check(getattr(L, f"oracle_c2r_{sfx}_with_planner")(_ptr(in_re), in_re.size, _ptr(in_im), in_im.size,
                                                             _ptr(out), out.size, planner._h, int(parallel)))
    else:
        _check(getattr(L, f"oracle_c2r_{sfx}_with_planner_and_scratch")(
            _ptr(in_re), in_re.size, _ptr(in_im), in_im.size, _ptr(out), out.size, planner._h,
            _ptr(scratch_re), scratch_re.size, _ptr(scratch_im), scratch_im.size, int(parallel)))


def bit_reverse(data: np.ndarray, tiled: bool = True) -> None:
    """bit_rev_bravo_f32 / f64 (bravo.rs:303-325): in-place bit-reversal permutation."""
    n = data.size
    assert n & (n - 1) == 0 and n > 0
    getattr(lib(), f"oracle_bit_reverse_{_sfx(data.dtype)}")(_ptr(data), n.bit_length() - 1, int(tiled))


def codelet(re: np.ndarray, im: np.ndarray) -> None:
    """fft_dit_codelet_16_f64 / fft_dit_codelet_32_f32 (codelets.rs:34,218)."""
    getattr(lib(), f"oracle_codelet_{_sfx(re.dtype)}")(_ptr(re), _ptr(im), re.size)


def codelet_stages(dtype) -> int:
    return getattr(lib(), f"oracle_codelet_stages_{_sfx(dtype)}")()


def stage(re: np.ndarray, im: np.ndarray, stage_idx: int) -> None:
    """One staged kernel fft_dit_chunk_{2..64} (kernels/dit.rs:13-967); stage_idx <= 5."""
    assert stage_idx <= 5
    getattr(lib(), f"oracle_stage_{_sfx(re.dtype)}")(_ptr(re), _ptr(im), re.size, stage_idx)


def literal_twiddles(chunk: int, dtype):
    re = np.empty(chunk // 2, dtype)
    im = np.empty(chunk // 2, dtype)
    getattr(lib(), f"oracle_literal_twiddles_{_sfx(dtype)}")(chunk, _ptr(re), _ptr(im))
    return re, im


def gen_random_signal(n: int, dtype, seed: int = 1234):
    """utilities/src/lib.rs:26-75 restated with a *seeded* generator: uniform[-1,1) re/im,
    scaled to unit L2 norm (the reference seeds from the OS, so there are no stored vectors)."""
    rng = np.random.default_rng(seed)
    re = rng.uniform(-1.0, 1.0, n).astype(dtype)
    im = rng.uniform(-1.0, 1.0, n).astype(dtype)
    mag = np.sqrt(np.sum(re.astype(dtype) ** 2 + im.astype(dtype) ** 2, dtype=dtype))
    scale = dtype(1.0) / mag if callable(dtype) else np.dtype(dtype).type(1.0) / mag
    return (re * scale).astype(dtype), (im * scale).astype(dtype)
