"""GPU: seam b1 -- the ORB_SLAM3::ORBextractor shim (shim/ORBextractor.cc) built against the stand-in OpenCV types and CALLED the way
UPSTREAM Frame::ExtractORB calls it (constructor of orbslam3_mono_networked.cc:193-206's parameters, operator(), the six getters)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from send_slam_b200 import orbx, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("shim") / "shim_driver.so")
    libdir = os.path.join(ROOT, "send_slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-DORBX_SHIM_CV_MIN", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "shim_driver.cc"), os.path.join(ROOT, "shim", "ORBextractor.cc"),
                           "-L", libdir, "-lorbx", "-Wl,-rpath," + libdir, "-o", so])
    L = C.CDLL(so)
    L.shim_new.restype = C.c_void_p
    L.shim_new.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.shim_delete.argtypes = [C.c_void_p]
    L.shim_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                               C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.shim_getters.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    return L


def run(L, ex, frame, lap=(0, 1000), cap=8000):
    kps = np.zeros(cap, orbx.KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n, rows = C.c_int(), C.c_int()
    mono = L.shim_extract(ex, frame.ctypes.data, frame.shape[1], frame.shape[0], frame.strides[0], lap[0], lap[1], kps.ctypes.data,
                          desc.ctypes.data, cap, C.byref(n), C.byref(rows))
    assert rows.value == n.value
    return mono, kps[:n.value], desc[:n.value]


def test_shim_operator_and_getters(shim, oracle):
    L = shim
    ex = L.shim_new(1250, 1.2, 8, 20, 7)               # the reference's YAML values
    o = oracle.Oracle(1250)
    for (w, h, kind, lap) in [(1280, 800, "textured", (0, 1000)), (640, 480, "mixed", (0, 1000)), (1920, 1080, "textured", (0, 1000)),
                              (640, 480, "textured", (100, 300))]:
        f = synth.textured_frame(900 + w, w, h, kind)
        mono, kps, desc = run(L, ex, f, lap)
        k_o, d_o, m_o = o.extract(f, lap=lap)
        assert mono == m_o and len(kps) == len(k_o), (w, h)
        assert np.array_equal(kps.view(np.int32), k_o.view(np.int32)) and np.array_equal(desc, d_o), (w, h)
    # constant image: no keypoints, descriptors released, monoIndex 0; strided sub-matrix view
    mono, kps, desc = run(L, ex, np.full((480, 640), 77, np.uint8))
    assert (mono, len(kps)) == (0, 0)
    big = synth.textured_frame(901, 700, 500)
    view = big[10:490, 20:660]
    mono, kps, desc = run(L, ex, view)
    k_o, d_o, m_o = o.extract(np.ascontiguousarray(view))
    assert mono == m_o and np.array_equal(desc, d_o)
    # getters == the oracle's constructor tables (Frame's constructor reads them)
    bufs = [np.zeros(8, np.float32) for _ in range(4)]
    sf = C.c_float()
    assert L.shim_getters(ex, *[b.ctypes.data for b in bufs], C.byref(sf)) == 8
    t = o.tables()
    assert abs(sf.value - 1.2) < 1e-7
    for got, name in zip(bufs, ["scale", "inv_scale", "sigma2", "inv_sigma2"]):
        assert np.array_equal(got, np.asarray(t[name], np.float32)[:8]), name
    L.shim_delete(ex)
