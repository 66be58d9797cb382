#!/usr/bin/env python
"""Build another copy of libphastft_cuda.so with extra nvcc defines, for A/B experiments on the GPU box:

    python tools/experiments/build_variant.py notw -DPHAST_EXP_STAGE_TW=1
    PHASTFT_LIB=build/variants/notw/libphastft_cuda.so python tools/timing.py f64 20 1

Objects and the library go to build/variants/<tag>/ (git-ignored, shipped by gpurun)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import __graft_entry__ as g  # noqa: E402

tag, defs = sys.argv[1], sys.argv[2:]
out = g.ROOT / "build" / "variants" / tag
out.mkdir(parents=True, exist_ok=True)
nvcc = g._nvcc()


def compile_unit(unit):
    name, src, udefs = unit
    obj = out / f"{name}.o"
    subprocess.run([nvcc, *g.NVCC_FLAGS, *udefs, *defs, "-c", "-o", str(obj), str(g.CSRC / src)], check=True)
    return obj


with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    objs = list(ex.map(compile_unit, g.UNITS))
lib = out / "libphastft_cuda.so"
subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), *map(str, objs), "-ldl"], check=True)
print(lib)
