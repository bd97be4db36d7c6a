"""CPU: the C-ABI library loads, exports every symbol include/orbx.h declares, fails loudly without a GPU, and its
host-side geometry plan agrees with the oracle (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from send_slam_b200 import orbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = "".join(open(os.path.join(ROOT, "include", f)).read() for f in ("orbx.h", "orbx_wire.h"))     # include/*.h
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = orbx.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"liborbx.so does not export {n}"
    assert sorted(orbx.EXPORTS) == names, "orbx.py EXPORTS and include/orbx.h disagree"
    assert b"sm_100a" in L.orbx_version()


def test_keypoint_record_is_cv_keypoint():
    assert orbx.KP_DTYPE.itemsize == 28
    assert [orbx.KP_DTYPE.fields[n][1] for n in orbx.KP_DTYPE.names] == [0, 4, 8, 12, 16, 20, 24]


def test_plan_matches_oracle_geometry(oracle):
    for (nf, w, h) in [(1000, 640, 480), (1200, 752, 480), (2000, 1920, 1080), (1250, 1280, 720), (500, 320, 240)]:
        p = orbx.plan_probe(nf, 1.2, 8, 20, 7, w, h)
        o = oracle.Oracle(nf)
        assert [(int(a), int(b)) for a, b in zip(p["widths"], p["heights"])] == [o.level_size(w, h, l) for l in range(8)]
        assert p["quota"].tolist() == o.tables()["quota"].tolist()
    # cells per frame and algorithmic bytes per frame
    assert int(orbx.plan_probe(1000, 1.2, 8, 20, 7, 640, 480)["ncells"].sum()) == 577
    assert int(orbx.plan_probe(1200, 1.2, 8, 20, 7, 752, 480)["ncells"].sum()) == 700
    # 4620 FAST calls in the reference at 1080p; 20 of them get a ROI with no testable pixel (height < 7) and are not
    # scheduled here
    assert int(orbx.plan_probe(2000, 1.2, 8, 20, 7, 1920, 1080)["ncells"].sum()) == 4600
    assert orbx.plan_probe(1000, 1.2, 8, 20, 7, 640, 480)["algorithmic_bytes"] == 5742474
    assert orbx.plan_probe(1200, 1.2, 8, 20, 7, 752, 480)["algorithmic_bytes"] == 6782935
    assert orbx.plan_probe(2000, 1.2, 8, 20, 7, 1920, 1080)["algorithmic_bytes"] == 32503669
    assert orbx.plan_probe(1200, 1.2, 8, 20, 7, 752, 480)["n_ini"].tolist()[0] == 2


def test_plan_rejects_bad_parameters():
    with pytest.raises(orbx.OrbxError):
        orbx.plan_probe(1000, 2.0, 8, 20, 7, 640, 480)      # scale 2 would be cv::resize's INTER_AREA path
    with pytest.raises(orbx.OrbxError):
        orbx.plan_probe(1000, 1.2, 17, 20, 7, 640, 480)
    with pytest.raises(orbx.OrbxError):
        orbx.plan_probe(1000, 1.2, 8, 20, 7, 5000, 480)
    with pytest.raises(orbx.OrbxError):
        orbx.plan_probe(1000, 1.2, 8, 5, 7, 640, 480)       # minTh > iniTh


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the host."""
    with pytest.raises(orbx.OrbxError) as e:
        orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == orbx.ORBX_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(orbx.OrbxError) as e:
        orbx.Knn2Index(np.zeros((4, 32), np.uint8))
    assert e.value.code == orbx.ORBX_E_CUDA
    L = orbx.lib()
    assert L.orbx_extract(None, None, 0, 0, 0, 0, 0, None, None, 0, None, None) == orbx.ORBX_E_INVALID
    assert L.orbx_knn2_query(None, None, 0, None, None) == orbx.ORBX_E_INVALID


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under send_slam_b200/ may reference it."""
    pkg = os.path.join(ROOT, "send_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "orb_oracle" not in txt and "from oracle" not in txt, f


def test_pnm_header_host_logic(golden_dir):
    """orbx_pnm_header is host code (no GPU): accept / reject verdict, frame geometry and payload offset equal what cv2's imdecode
    did for the fixture's byte strings (tests/golden/make_golden_pnm.py) and what the oracle's restatement says."""
    import numpy as np
    from oracle import oracle_lib
    from send_slam_b200 import orbx
    g = np.load(os.path.join(golden_dir, "pnm_cases.npz"))
    blob, pos = g["blob"].tobytes(), 0
    for i, ln in enumerate(g["lengths"]):
        data, verdict, shape = blob[pos:pos + ln], int(g["verdict"][i]), tuple(int(v) for v in g["shape"][i])
        pos += ln
        if verdict == 2:
            with pytest.raises(orbx.OrbxError):
                orbx.pnm_header(data)
            continue
        got = orbx.pnm_header(data)
        if verdict == 0:
            assert got is None, data[:16]
            continue
        w, h, ch, off = got
        assert (h, w, ch) == shape, data[:16]
        m = oracle_lib.pnm_decode(data)
        pay = np.frombuffer(data, np.uint8)[off:off + w * h * ch].reshape(h, w, ch)
        assert np.array_equal(pay[:, :, ::-1] if ch == 3 else pay[:, :, 0], m), data[:16]
    assert orbx.pnm_header(b"") is None


def test_wire_messages_host_logic():
    """include/orbx_wire.h against the Python msgpack package as the independent MessagePack implementation: the 'frame' message the
    Elixir side builds (slam_handler.ex:140-156) parses to what ParseMessage extracts (orbslam3_mono_networked.cc:302-337), malformed
    payloads are refused the way msgpack-c's unpack / convert throw, and the 'features' message round-trips both ways."""
    msgpack = pytest.importorskip("msgpack")
    ppm = b"P6\n4 2\n255\n" + bytes(range(24))
    frame = {"type": "frame", "camera_id": 1, "encoding": "ppm", "timestamp": 12.5, "width": 4, "height": 2, "channels": 3, "frame": ppm}
    m = orbx.wire_parse_frame(msgpack.packb(frame, use_bin_type=True))
    assert m == {"type": "frame", "image": ppm, "timestamp": 12.5, "camera_id": 1}
    # key order, the "image" alias, integer / float32 timestamps, unknown fields of every MessagePack family, big frames (bin32)
    import struct
    big = bytes(70000)
    extra = {"nested": {"a": [1, -2, 3.5, None, True, {"b": b"x" * 300, "c": "y" * 40}], "d": [[]], "e": 2 ** 40, "f": -2 ** 33},
             "ext": msgpack.ExtType(5, b"12345"), "ext8": msgpack.ExtType(1, b"12345678"), "calibration": {"fx": 1.0, "dist": [0.1, 0.2]}}
    msg = dict(extra, image=big, camera_id=300, timestamp=7, type="frame")
    m = orbx.wire_parse_frame(msgpack.packb(msg, use_bin_type=True))
    assert m["image"] == big and m["camera_id"] == 300 and m["timestamp"] == 7.0 and m["type"] == "frame"
    f32 = b"\x83" + msgpack.packb("type") + msgpack.packb("frame") + msgpack.packb("timestamp") + b"\xca" + struct.pack(">f", 1.25) + \
        msgpack.packb("camera_id") + b"\xd1" + struct.pack(">h", -7)
    m = orbx.wire_parse_frame(f32)
    assert m["timestamp"] == 1.25 and m["camera_id"] == -7 and m["image"] is None
    dup = b"\x83" + msgpack.packb("type") + msgpack.packb("frame") + msgpack.packb("camera_id") + msgpack.packb(1) + msgpack.packb("camera_id") + msgpack.packb(2)
    assert orbx.wire_parse_frame(dup)["camera_id"] == 2                                   # later duplicate wins, as in the loop
    assert orbx.wire_parse_frame(msgpack.packb({"type": "calibration", "calibration": {"fx": 1}}))["type"] == "calibration"
    good = msgpack.packb(frame, use_bin_type=True)
    bad = [b"", b"\xc1", msgpack.packb([1, 2]), msgpack.packb("frame"), good[:-1], good[:10],               # empty, reserved tag, non-map, truncated
           msgpack.packb({"type": "frame", "frame": "not-bin"}, use_bin_type=True),                           # image must be bin
           msgpack.packb({"type": "frame", "timestamp": "soon"}), msgpack.packb({"type": "frame", "camera_id": 1.5}),
           msgpack.packb({"type": "frame", "camera_id": 2 ** 31}), msgpack.packb({"type": ""}), msgpack.packb({"camera_id": 1}),
           msgpack.packb({1: "frame"}), msgpack.packb({"type": "frame", "x": [1, 2, 3]})[:-1], b"\xdf\xff\xff\xff\xff", b"\x81\xa4type\xdd\xff\xff\xff\xff"]
    for b in bad:
        with pytest.raises(orbx.OrbxError):
            orbx.wire_parse_frame(b)
    # the 'features' message: library -> msgpack package, msgpack package -> library, library -> library
    rng = np.random.default_rng(3)
    n = 37
    kps = np.zeros(n, orbx.KP_DTYPE)
    for name in ("x", "y", "size", "angle", "response"):
        kps[name] = rng.uniform(0, 640, n).astype(np.float32)
    kps["octave"] = rng.integers(0, 8, n); kps["class_id"] = -1
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    wire = orbx.wire_pack_features(3.75, 2, 640, 480, 11, kps, desc, framed=True)
    assert struct.unpack(">I", wire[:4])[0] == len(wire) - 4
    d = msgpack.unpackb(wire[4:], raw=False)
    assert d == {"type": "features", "camera_id": 2, "timestamp": 3.75, "width": 640, "height": 480, "mono_index": 11, "n": n,
                 "keypoints": kps.tobytes(), "descriptors": desc.tobytes()}
    for payload in (wire[4:], msgpack.packb(dict(reversed(list(d.items())), note="extra"), use_bin_type=True)):
        f = orbx.wire_parse_features(payload)
        assert (f["timestamp"], f["camera_id"], f["width"], f["height"], f["mono_index"]) == (3.75, 2, 640, 480, 11)
        assert f["keypoints"].tobytes() == kps.tobytes() and np.array_equal(f["descriptors"], desc)
    empty = orbx.wire_parse_features(orbx.wire_pack_features(1.0, 1, 64, 48, 0, kps[:0], desc[:0], framed=False))
    assert len(empty["keypoints"]) == 0 and empty["descriptors"].shape == (0, 32)
    for b in (msgpack.packb(frame, use_bin_type=True), wire[4:-1], msgpack.packb(dict(d, n=n + 1), use_bin_type=True),
              msgpack.packb(dict(d, keypoints="str"), use_bin_type=True)):
        with pytest.raises(orbx.OrbxError):
            orbx.wire_parse_features(b)
