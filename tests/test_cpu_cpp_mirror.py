"""cpp/phastft.hpp (the C++ host-side mirror of the reference API) compiles against the C ABI and
its validation / error mapping behaves like the reference's panics.  On the GPU box the same
binary also runs a transform."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_cpp_mirror_builds_and_maps_panics(tmp_path):
    import __graft_entry__ as g
    g.build()
    gxx = shutil.which("g++") or "/usr/bin/g++"
    exe = tmp_path / "test_phastft_hpp"
    libdir = ROOT / "phastft_b200"
    subprocess.run([gxx, "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "cpp" / "test_phastft_hpp.cpp"),
                    f"-L{libdir}", "-lphastft_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
