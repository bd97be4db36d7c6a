"""CPU: the two reference-side seams compile against the C ABI (shim with stand-in OpenCV types, NIF with stand-in
erl_nif declarations -- the real headers are absent from this image)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_shim_compiles(tmp_path):
    subprocess.check_call(["g++", "-std=c++17", "-DORBX_SHIM_CV_MIN", "-Wall", "-Werror", "-c",
                           os.path.join(ROOT, "shim", "ORBextractor.cc"), "-o", str(tmp_path / "shim.o")])


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_nif_compiles(tmp_path):
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-DORBX_NIF_MIN", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "nif"), "-c", os.path.join(ROOT, "nif", "orbx_nif.c"), "-o", str(tmp_path / "nif.o")])


def test_shim_distance_matches_oracle(oracle, tmp_path):
    src = tmp_path / "d.cc"
    src.write_text('#include "%s"\nextern "C" int dd(const unsigned char*a,const unsigned char*b){return orbx_descriptor_distance(a,b);}\n'
                   % os.path.join(ROOT, "shim", "ORBmatcherDistance.h"))
    so = tmp_path / "d.so"
    subprocess.check_call(["g++", "-shared", "-fPIC", "-O2", str(src), "-o", str(so)])
    import ctypes
    import numpy as np
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        assert lib.dd(a[i].ctypes.data_as(ctypes.c_void_p), b[i].ctypes.data_as(ctypes.c_void_p)) == oracle.distance(a[i], b[i])
