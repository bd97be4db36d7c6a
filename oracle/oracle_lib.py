"""ctypes binding of oracle/_build/liborb_oracle.so (ORACLE = test infrastructure; see orb_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "orb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, f32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orb_oracle_create.restype = C.c_void_p
        L.orb_oracle_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orb_oracle_destroy.argtypes = [C.c_void_p]
        L.orb_oracle_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
        L.orb_oracle_level_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.orb_oracle_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orb_oracle_blur7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_fast_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_fast_atan2.restype = C.c_float
        L.orb_oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orb_oracle_ic_moments.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]
        L.orb_oracle_ic_angle.restype = C.c_float
        L.orb_oracle_ic_angle.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.orb_oracle_set_trig_mode.argtypes = [C.c_int]
        L.orb_oracle_brief.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orb_oracle_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, i32p, i32p]
        L.orb_oracle_stage_level.argtypes = [C.c_void_p, C.c_int, i32p, i32p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orb_oracle_stage_keys.argtypes = [C.c_void_p, C.c_int, i32p, C.POINTER(C.c_void_p), i32p, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orb_oracle_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orb_oracle_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orb_oracle_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_int]
        L.orb_oracle_match_windowed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_oracle_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orb_oracle_image_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orb_oracle_frame_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_oracle_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_oracle_pnm_header.argtypes = [C.c_void_p, C.c_size_t, i32p, i32p, i32p, C.POINTER(C.c_size_t)]
        L.orb_oracle_pnm_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One ORBextractor restatement instance."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.orb_oracle_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("orb_oracle_create rejected the parameters")
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        self.scale_factor = scale_factor

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orb_oracle_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        quota, umax = np.zeros(n, np.int32), np.zeros(16, np.int32)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        self.L.orb_oracle_tables(self.h, sc.ctypes.data_as(f32p), inv.ctypes.data_as(f32p), s2.ctypes.data_as(f32p),
                                 is2.ctypes.data_as(f32p), quota.ctypes.data_as(i32p), umax.ctypes.data_as(i32p))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, quota=quota, umax=umax)

    def level_size(self, w, h, l):
        a, b = C.c_int32(), C.c_int32()
        self.L.orb_oracle_level_size(self.h, w, h, l, C.byref(a), C.byref(b))
        return a.value, b.value

    def extract(self, img: np.ndarray, lap=(0, 1000), cap=None):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        cap = cap or (self.nfeatures + 4 * self.nlevels + 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int32(), C.c_int32()
        rc = self.L.orb_oracle_extract(self.h, _p(img), w, h, w, lap[0], lap[1], _p(kps), _p(desc), cap,
                                       C.byref(n), C.byref(mono))
        if rc != 0:
            raise RuntimeError(f"orb_oracle_extract rc={rc}")
        assert n.value <= cap
        return kps[:n.value].copy(), desc[:n.value].copy(), mono.value

    def stage_level(self, l):
        w, h = C.c_int32(), C.c_int32()
        pix, bl = C.c_void_p(), C.c_void_p()
        self.L.orb_oracle_stage_level(self.h, l, C.byref(w), C.byref(h), C.byref(pix), C.byref(bl))
        shape = (h.value, w.value)
        a = np.ctypeslib.as_array(C.cast(pix, C.POINTER(C.c_uint8)), shape=shape).copy()
        b = np.ctypeslib.as_array(C.cast(bl, C.POINTER(C.c_uint8)), shape=shape).copy() if bl.value else None
        return a, b

    def stage_keys(self, l):
        nc, ns = C.c_int32(), C.c_int32()
        cand, sel, ang, desc = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.L.orb_oracle_stage_keys(self.h, l, C.byref(nc), C.byref(cand), C.byref(ns), C.byref(sel), C.byref(ang), C.byref(desc))
        c = np.ctypeslib.as_array(C.cast(cand, C.POINTER(C.c_float)), shape=(nc.value, 3)).copy() if nc.value else np.zeros((0, 3), np.float32)
        if ns.value:
            s = np.ctypeslib.as_array(C.cast(sel, C.POINTER(C.c_int32)), shape=(ns.value,)).copy()
            a = np.ctypeslib.as_array(C.cast(ang, C.POINTER(C.c_float)), shape=(ns.value,)).copy()
            d = np.ctypeslib.as_array(C.cast(desc, C.POINTER(C.c_uint8)), shape=(ns.value, 32)).copy()
        else:
            s, a, d = np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros((0, 32), np.uint8)
        return c, s, a, d


def set_trig_mode(mode: int):
    """1 (default) = correctly rounded cos/sin, 0 = this host's libm cosf/sinf (see orb_oracle.c 'Trig rule')."""
    lib().orb_oracle_set_trig_mode(int(mode))


def resize(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orb_oracle_resize(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def gray(src: np.ndarray, fmt: int, shift: int = 15) -> np.ndarray:
    """cv::cvtColor(*2GRAY) restatement; fmt 1 RGB, 2 BGR, 3 RGBA, 4 BGRA."""
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros(src.shape[:2], np.uint8)
    rc = lib().orb_oracle_gray(_p(src), src.shape[1], src.shape[0], src.strides[0], int(fmt), int(shift), _p(dst), dst.shape[1])
    assert rc == 0
    return dst


def pnm_decode(data: bytes):
    """cv::imdecode(data, IMREAD_UNCHANGED) for binary PNM: gray [H,W] (P5) or BGR [H,W,3] (P6); None where imdecode returns an
    empty Mat; raises ValueError for variants the extractor cannot take (ASCII / bitmap / 16-bit)."""
    buf = np.frombuffer(data, np.uint8)
    w, h, ch, off = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t()
    rc = lib().orb_oracle_pnm_header(_p(buf), len(buf), C.byref(w), C.byref(h), C.byref(ch), C.byref(off)) if len(buf) else -1
    if rc == -1:
        return None
    if rc == -2:
        raise ValueError("PNM variant outside the CV_8U binary forms")
    dst = np.zeros((h.value, w.value) if ch.value == 1 else (h.value, w.value, 3), np.uint8)
    assert lib().orb_oracle_pnm_decode(_p(buf), len(buf), _p(dst)) == 0
    return dst


def blur7(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orb_oracle_blur7(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast(img: np.ndarray, t: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.zeros((cap, 3), np.float32)
    n = lib().orb_oracle_fast(_p(img), img.shape[1], img.shape[0], img.shape[1], t, _p(out), cap)
    return out[:n].copy()


def fast_cells(img: np.ndarray, ini_th=20, min_th=7) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.zeros((cap, 3), np.float32)
    n = lib().orb_oracle_fast_cells(_p(img), img.shape[1], img.shape[0], img.shape[1], ini_th, min_th, _p(out), cap)
    return out[:n].copy()


def octree(keys: np.ndarray, minX, maxX, minY, maxY, N) -> np.ndarray:
    keys = np.ascontiguousarray(keys, np.float32).reshape(-1, 3)
    out = np.zeros(max(len(keys), 1), np.int32)
    m = lib().orb_oracle_octree(_p(keys), len(keys), minX, maxX, minY, maxY, N, _p(out), len(out))
    if m < 0:
        raise ValueError("octree: bad aspect ratio")
    return out[:m].copy()


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orb_oracle_fast_atan2(float(y), float(x)))


def ic_moments(img: np.ndarray, cx: int, cy: int, umax: np.ndarray):
    img = np.ascontiguousarray(img, np.uint8)
    a, b = C.c_int32(), C.c_int32()
    um = np.ascontiguousarray(umax, np.int32)
    lib().orb_oracle_ic_moments(_p(img), img.shape[1], int(cx), int(cy), um.ctypes.data_as(C.POINTER(C.c_int32)),
                                C.byref(a), C.byref(b))
    return a.value, b.value  # m01, m10


def brief(blurred: np.ndarray, x: float, y: float, angle: float) -> np.ndarray:
    blurred = np.ascontiguousarray(blurred, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orb_oracle_brief(_p(blurred), blurred.shape[1], float(x), float(y), float(angle), _p(d))
    return d


def extract_batch(frames: np.ndarray, nfeatures, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lap=(0, 1000),
                  nthreads=1):
    """frames: [B,H,W] uint8. Returns (kps [B,cap], desc [B,cap,32], n [B], mono [B])."""
    frames = np.ascontiguousarray(frames, np.uint8)
    B, H, W = frames.shape
    cap = nfeatures + 4 * nlevels + 64
    kps = np.zeros((B, cap), KP_DTYPE)
    desc = np.zeros((B, cap, 32), np.uint8)
    n = np.zeros(B, np.int32)
    mono = np.zeros(B, np.int32)
    lib().orb_oracle_extract_batch(nfeatures, scale_factor, nlevels, ini_th, min_th, _p(frames), W, H, W, H * W, B,
                                   lap[0], lap[1], _p(kps), _p(desc), cap, _p(n), _p(mono), nthreads)
    return kps, desc, n, mono


def distance(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orb_oracle_distance(_p(a), _p(b)))


def knn2(q: np.ndarray, db: np.ndarray, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    idx = np.zeros((len(q), 2), np.int32)
    dist = np.zeros((len(q), 2), np.int32)
    lib().orb_oracle_knn2(_p(q), len(q), _p(db), len(db), _p(idx), _p(dist), nthreads)
    return idx, dist


def match_windowed(qdesc, quvr, qlevels, tkp, tdesc, bounds):
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    quvr = np.ascontiguousarray(quvr, np.float32)
    qlevels = np.ascontiguousarray(qlevels, np.int32)
    tkp = np.ascontiguousarray(tkp, KP_DTYPE)
    tdesc = np.ascontiguousarray(tdesc, np.uint8)
    bounds = np.ascontiguousarray(bounds, np.float32)
    nq = len(qdesc)
    out = [np.zeros(nq, np.int32) for _ in range(4)]
    lib().orb_oracle_match_windowed(_p(qdesc), _p(quvr), _p(qlevels), nq, _p(tkp), _p(tdesc), len(tkp), _p(bounds),
                                    _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]))
    return tuple(out)  # best_idx, best_dist, second_idx, second_dist


def _cam(cam) -> np.ndarray:
    c = np.zeros(9, np.float32)
    c[:len(cam)] = np.asarray(cam, np.float32)
    return c


def undistort_points(xy: np.ndarray, cam) -> np.ndarray:
    """cv::undistortPoints(xy, K, D, P=K) restatement; cam = (fx, fy, cx, cy, k1, k2, p1, p2[, k3])."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros_like(xy)
    c = _cam(cam)
    lib().orb_oracle_undistort_points(_p(xy), len(xy), _p(c), _p(out))
    return out


def image_bounds(cam, w: int, h: int) -> np.ndarray:
    b = np.zeros(4, np.float32)
    c = _cam(cam)
    lib().orb_oracle_image_bounds(_p(c), int(w), int(h), _p(b))
    return b


def frame_grid(kps: np.ndarray, cam, bounds):
    """Frame::UndistortKeyPoints + AssignFeaturesToGrid.  Returns (kps_un, cell_start[3073], cell_items[inside])."""
    kps = np.ascontiguousarray(kps)
    un = np.zeros_like(kps)
    start = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(len(kps), 1), np.int32)
    c, b = _cam(cam), np.ascontiguousarray(bounds, np.float32)
    lib().orb_oracle_frame_grid(_p(kps), len(kps), _p(c), _p(b), _p(un), _p(start), _p(items))
    return un, start, items[:start[-1]]


def bow_transform(parent, ndesc, weight, feat, levelsup=4):
    """DBoW2 transform restatement.  Returns (word_id, word_weight, node_id, depth)."""
    parent = np.ascontiguousarray(parent, np.int32); ndesc = np.ascontiguousarray(ndesc, np.uint8)
    weight = np.ascontiguousarray(weight, np.float32); feat = np.ascontiguousarray(feat, np.uint8)
    n = len(feat)
    w, wt, nid = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    depth = lib().orb_oracle_bow_transform(_p(parent), _p(ndesc), _p(weight), len(parent), _p(feat), n, int(levelsup), _p(w), _p(wt), _p(nid))
    return w, wt, nid, depth
