"""GPU: the erl_nif wrapper (nif/orbx_nif.c) loaded and CALLED through the mock BEAM host (nif/mock_host.c) -- SURVEY.md §8b
seam b3 and the shape of BASELINE config 5: 8 concurrent 1280x720 camera streams, nFeatures 1250 (the reference's YAML value,
orbslam3_mono_networked.cc:193), extraction + previous-frame windowed matching, each stream on its own NIF resource from its
own thread (a dirty-scheduler call per frame in the BEAM; here ctypes drops the GIL for the duration of the call)."""
import ctypes as C
import os
import shutil
import subprocess
import threading

import numpy as np
import pytest

from send_slam_b200 import orbx, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_INT, T_ATOM, T_BIN, T_TUPLE, T_RES, T_BADARG = 1, 3, 4, 5, 6, 7


@pytest.fixture(scope="module")
def nif(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    so = str(tmp_path_factory.mktemp("nif") / "orbx_nif_mock.so")
    libdir = os.path.join(ROOT, "send_slam_b200")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-DORBX_NIF_MIN", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "nif"), "-shared", "-fPIC", os.path.join(ROOT, "nif", "orbx_nif.c"),
                           os.path.join(ROOT, "nif", "mock_host.c"), "-L", libdir, "-lorbx", "-lpthread", "-Wl,-rpath," + libdir, "-o", so])
    L = C.CDLL(so)
    ul = C.c_ulong
    L.mock_int.restype = ul; L.mock_int.argtypes = [C.c_int]
    L.mock_long.restype = ul; L.mock_long.argtypes = [C.c_long]
    L.mock_double.restype = ul; L.mock_double.argtypes = [C.c_double]
    L.mock_binary.restype = ul; L.mock_binary.argtypes = [C.c_void_p, C.c_size_t]
    L.mock_tuple4.restype = ul; L.mock_tuple4.argtypes = [ul, ul, ul, ul]
    L.mock_resource_keep.restype = C.c_void_p; L.mock_resource_keep.argtypes = [ul]
    L.mock_resource_term.restype = ul; L.mock_resource_term.argtypes = [C.c_void_p]
    L.mock_resource_drop.argtypes = [C.c_void_p]
    L.mock_call.restype = ul; L.mock_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(ul)]
    L.mock_kind.argtypes = [ul]; L.mock_get_int.restype = C.c_long; L.mock_get_int.argtypes = [ul]
    L.mock_get_atom.restype = C.c_char_p; L.mock_get_atom.argtypes = [ul]
    L.mock_tuple_arity.argtypes = [ul]; L.mock_tuple_elem.restype = ul; L.mock_tuple_elem.argtypes = [ul, C.c_int]
    L.mock_bin_size.restype = C.c_size_t; L.mock_bin_size.argtypes = [ul]
    L.mock_bin_data.restype = C.c_void_p; L.mock_bin_data.argtypes = [ul]
    assert L.orbx_nif_mock_load() == 0
    return L


def call(L, name, *terms):
    arr = (C.c_ulong * len(terms))(*terms)
    return L.mock_call(name.encode(), len(terms), arr)


def decode(L, t):
    k = L.mock_kind(t)
    if k == T_INT:
        return int(L.mock_get_int(t))
    if k == T_ATOM:
        return L.mock_get_atom(t).decode()
    if k == T_BIN:
        return C.string_at(L.mock_bin_data(t), L.mock_bin_size(t))
    if k == T_TUPLE:
        return tuple(decode(L, L.mock_tuple_elem(t, i)) for i in range(L.mock_tuple_arity(t)))
    if k == T_BADARG:
        return "badarg"
    return ("term", k)


def nif_create(L, nfeatures, w, h, max_batch=None):
    args = [L.mock_int(nfeatures), L.mock_double(1.2), L.mock_int(8), L.mock_int(20), L.mock_int(7), L.mock_int(0), L.mock_int(w), L.mock_int(h)]
    if max_batch is not None:
        args.append(L.mock_int(max_batch))
    r = call(L, "create", *args)
    assert L.mock_kind(r) == T_TUPLE and decode(L, L.mock_tuple_elem(r, 0)) == "ok", decode(L, r)
    res = L.mock_resource_keep(L.mock_tuple_elem(r, 1))      # the "process" keeps the handle term alive across calls
    L.mock_reset()
    return res


def nif_extract(L, res, frame, fmt=None):
    h, w = frame.shape[:2]
    buf = np.ascontiguousarray(frame)
    args = [L.mock_resource_term(res), L.mock_binary(buf.ctypes.data, buf.nbytes), L.mock_int(w), L.mock_int(h)]
    r = decode(L, call(L, "extract", *args) if fmt is None else call(L, "extract_color", *args, L.mock_int(fmt)))
    L.mock_reset()
    assert r[0] == "ok", r
    _, n, mono, kp, desc = r
    return mono, np.frombuffer(kp, orbx.KP_DTYPE).copy(), np.frombuffer(desc, np.uint8).reshape(n, 32).copy()


def test_nif_errors_and_single_stream(nif, oracle):
    L = nif
    w, h, nf = 1280, 720, 1250
    res = nif_create(L, nf, w, h)
    f = synth.textured_frame(700, w, h)
    mono, kps, desc = nif_extract(L, res, f)
    k_o, d_o, m_o = oracle.Oracle(nf).extract(f)
    assert mono == m_o and np.array_equal(desc, d_o) and np.array_equal(kps.view(np.int32), k_o.view(np.int32))
    # colour binary (BGR, what Evision hands over): gray conversion on the device
    bgr = np.stack([f, f, f], -1)
    mono2, kps2, desc2 = nif_extract(L, res, bgr, fmt=orbx.FMT_BGR8)
    g = oracle.gray(bgr, orbx.FMT_BGR8, 15)
    k_g, d_g, m_g = oracle.Oracle(nf).extract(g)
    assert mono2 == m_g and np.array_equal(desc2, d_g)
    # the wire's PPM binary as SlamHandler builds it (Evision.imencode(".ppm"), slam_handler.ex:275-277): header parse + imdecode's
    # BGR order + cvtColor by Camera.RGB all inside the library
    col = np.stack([synth.textured_frame(701 + k, w, h) for k in range(3)], -1)             # the BGR Mat the camera produced
    ppm = b"P6\n%d %d\n255\n" % (w, h) + col[:, :, ::-1].tobytes()                          # PPM stores RGB
    for camera_rgb in (1, 0):
        r = decode(L, call(L, "extract_ppm", L.mock_resource_term(res), L.mock_binary(ppm, len(ppm)), L.mock_int(camera_rgb)))
        L.mock_reset()
        assert r[0] == "ok" and r[5:] == (w, h), r[:3]
        gp = oracle.gray(oracle.pnm_decode(ppm), 1 if camera_rgb else 2, 15)
        k_p, d_p, m_p = oracle.Oracle(nf).extract(gp)
        assert r[1] == len(k_p) and r[2] == m_p and np.array_equal(np.frombuffer(r[4], np.uint8).reshape(-1, 32), d_p)
        assert np.array_equal(np.frombuffer(r[3], np.int32), k_p.view(np.int32).ravel())
    r = decode(L, call(L, "extract_ppm", L.mock_resource_term(res), L.mock_binary(ppm, len(ppm) - 5), L.mock_int(1)))
    L.mock_reset()
    assert r == ("error", "empty_image")                        # truncated: the backend would skip this frame
    # error tuples, never exceptions: size mismatch, frame larger than the handle was created for, bad arguments
    r = decode(L, call(L, "extract", L.mock_resource_term(res), L.mock_binary(f.ctypes.data, 100), L.mock_int(w), L.mock_int(h)))
    L.mock_reset()
    assert r == ("error", "size_mismatch")
    big = np.zeros((800, 1400), np.uint8)
    r = decode(L, call(L, "extract", L.mock_resource_term(res), L.mock_binary(big.ctypes.data, big.nbytes), L.mock_int(1400), L.mock_int(800)))
    L.mock_reset()
    assert r == ("error", "capacity")
    assert decode(L, call(L, "extract", L.mock_int(3), L.mock_int(4), L.mock_int(5), L.mock_int(6))) == "badarg"
    L.mock_reset()
    L.mock_resource_drop(res)                                  # last reference: the resource destructor destroys the handle


def test_eight_concurrent_streams(nif, oracle):
    """BASELINE config 5 on one GPU: 8 streams x 1280x720, each frame extracted and matched against the stream's previous frame."""
    L = nif
    w, h, nf, nstreams, nframes = 1280, 720, 1250, 8, 3
    frames = [[synth.shifted_frame(synth.textured_frame(800 + s, w, h), 3 * t, -2 * t, seed=s) for t in range(nframes)] for s in range(nstreams)]
    out = [None] * nstreams
    errs = []

    def stream(s):
        try:
            res = nif_create(L, nf, w, h)
            got, prev = [], None
            for t in range(nframes):
                mono, kps, desc = nif_extract(L, res, frames[s][t])
                m = None
                if prev is not None:
                    pk, pd = prev
                    quvr = np.stack([pk["x"] + 3, pk["y"] - 2, 15.0 * np.float32(1.2) ** pk["octave"]], 1).astype(np.float32)
                    qlev = np.stack([pk["octave"] - 1, pk["octave"] + 1], 1).astype(np.int32)
                    b = L.mock_tuple4(L.mock_double(0.0), L.mock_double(0.0), L.mock_double(float(w)), L.mock_double(float(h)))
                    r = decode(L, call(L, "match_windowed", L.mock_resource_term(res), L.mock_binary(pd.ctypes.data, pd.nbytes),
                                       L.mock_binary(quvr.ctypes.data, quvr.nbytes), L.mock_binary(qlev.ctypes.data, qlev.nbytes),
                                       L.mock_binary(kps.ctypes.data, kps.nbytes), L.mock_binary(desc.ctypes.data, desc.nbytes), b))
                    L.mock_reset()
                    assert r[0] == "ok", r
                    m = (tuple(np.frombuffer(x, np.int32).copy() for x in r[1:]), quvr, qlev)
                got.append((mono, kps, desc, m))
                prev = (kps, desc)
            L.mock_resource_drop(res)
            out[s] = got
        except Exception as e:                                  # surface failures of worker threads
            errs.append((s, repr(e)))

    th = [threading.Thread(target=stream, args=(s,)) for s in range(nstreams)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    o = oracle.Oracle(nf)
    for s in range(nstreams):
        for t in range(nframes):
            mono, kps, desc, m = out[s][t]
            if s in (0, 5) or t == 0:                            # full oracle comparison on a subset keeps the test short
                k_o, d_o, m_o = o.extract(frames[s][t])
                assert mono == m_o and np.array_equal(desc, d_o) and np.array_equal(kps.view(np.int32), k_o.view(np.int32)), (s, t)
            if m is not None and s in (0, 5):
                (bi, bd, si, sd), quvr, qlev = m
                pk, pd = out[s][t - 1][1], out[s][t - 1][2]
                want = oracle.match_windowed(pd, quvr, qlev, kps, desc, np.array([0, 0, w, h], np.float32))
                assert np.array_equal(bi, want[0]) and np.array_equal(bd, want[1]) and np.array_equal(si, want[2]) and np.array_equal(sd, want[3])
                assert ((bi >= 0) & (bd <= 100)).mean() > 0.4


def test_one_handle_shared_by_several_processes(nif, oracle):
    """A resource term can reach several BEAM processes; the handle underneath is single-flight.  Every call either runs alone
    (result identical to the oracle's) or comes back as {:error, :busy} -- never a torn result, never a crash."""
    L = nif
    w, h, nf = 640, 480, 1000
    res = nif_create(L, nf, w, h)
    frames = [synth.textured_frame(900 + k, w, h) for k in range(4)]
    want = [oracle.Oracle(nf).extract(f) for f in frames]
    stats = {"ok": 0, "busy": 0}
    errs = []
    lock = threading.Lock()

    def proc(k):
        try:
            for it in range(12):
                f = frames[(k + it) % 4]
                buf = np.ascontiguousarray(f)
                r = decode(L, call(L, "extract", L.mock_resource_term(res), L.mock_binary(buf.ctypes.data, buf.nbytes), L.mock_int(w), L.mock_int(h)))
                L.mock_reset()
                if r == ("error", "busy"):
                    with lock:
                        stats["busy"] += 1
                    continue
                assert r[0] == "ok", r
                k_o, d_o, m_o = want[(k + it) % 4]
                assert r[1] == len(k_o) and r[2] == m_o and r[4] == d_o.tobytes() and r[3] == k_o.tobytes()
                with lock:
                    stats["ok"] += 1
        except Exception as e:
            errs.append(repr(e))

    th = [threading.Thread(target=proc, args=(k,)) for k in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert stats["ok"] >= 4 and stats["ok"] + stats["busy"] == 72, stats
    L.mock_resource_drop(res)


def test_nif_extract_batch_and_knn2(nif, oracle):
    L = nif
    w, h, nf, B = 640, 480, 1000, 5
    res = nif_create(L, nf, w, h, max_batch=8)
    frames = np.stack([synth.textured_frame(950 + k, w, h, "mixed" if k == 2 else "textured") for k in range(B)])
    r = decode(L, call(L, "extract_batch", L.mock_resource_term(res), L.mock_binary(frames.ctypes.data, frames.nbytes), L.mock_int(B),
                       L.mock_int(w), L.mock_int(h)))
    L.mock_reset()
    assert r[0] == "ok", r
    _, nb, mb, kpb, deb, cap = r
    n, mono = np.frombuffer(nb, np.int32), np.frombuffer(mb, np.int32)
    kps = np.frombuffer(kpb, orbx.KP_DTYPE).reshape(B, cap)
    desc = np.frombuffer(deb, np.uint8).reshape(B, cap, 32)
    o = oracle.Oracle(nf)
    for i in range(B):
        k_o, d_o, m_o = o.extract(frames[i])
        assert n[i] == len(k_o) and mono[i] == m_o and np.array_equal(desc[i, :n[i]], d_o) and kps[i, :n[i]].tobytes() == k_o.tobytes(), i
    r = decode(L, call(L, "extract_batch", L.mock_resource_term(res), L.mock_binary(frames.ctypes.data, frames.nbytes), L.mock_int(9),
                       L.mock_int(w), L.mock_int(h)))
    L.mock_reset()
    assert r == ("error", "capacity")
    L.mock_resource_drop(res)
    # brute-force kNN (k = 2) through the NIF: a database shard as one binary, both distance backends, global row indices
    db = synth.descriptor_db(30011, seed=21)
    q, _ = synth.queries_from_db(db, 77, seed=22)
    r = call(L, "knn2_create", L.mock_binary(db.ctypes.data, db.nbytes), L.mock_int(0), L.mock_long(4_000_000_000))
    assert decode(L, L.mock_tuple_elem(r, 0)) == "ok", decode(L, r)
    dbres = L.mock_resource_keep(L.mock_tuple_elem(r, 1))
    L.mock_reset()
    idx_o, dist_o = oracle.knn2(q, db)
    for backend in (orbx.Knn2Index.POPC, orbx.Knn2Index.TENSOR, orbx.Knn2Index.TENSOR_FP4):
        r = decode(L, call(L, "knn2", L.mock_resource_term(dbres), L.mock_binary(q.ctypes.data, q.nbytes), L.mock_int(backend)))
        L.mock_reset()
        assert r[0] == "ok", r
        idx, dist = np.frombuffer(r[1], np.int32).reshape(-1, 2), np.frombuffer(r[2], np.int32).reshape(-1, 2)
        # the 32-bit index field carries the low word of (row_offset + row): row_offset + row modulo 2^32 (the ABI keeps row_offset + rows below 2^32)
        assert np.array_equal(dist, dist_o) and np.array_equal(idx.astype(np.int64) & 0xFFFFFFFF, (idx_o.astype(np.int64) + 4_000_000_000) & 0xFFFFFFFF), backend
    r = decode(L, call(L, "knn2", L.mock_resource_term(dbres), L.mock_binary(q.ctypes.data, 33), L.mock_int(1)))
    L.mock_reset()
    assert r == ("error", "size_mismatch")
    L.mock_resource_drop(dbres)
