"""CPU: the backend half of the `features` message.  integration/orbslam3_mono_networked.features.patch adds a branch to
the receive loop of slam_backends/orb_slam_3/orbslam3_mono_networked.cc; tests/harness/features_loop_driver.cc is that branch's body
without ORB-SLAM3.  Messages are packed here with the independent `msgpack` package and with the library's own writer, framed as the
socket frames them (4-byte big-endian length), and must come out of the C++ side as the same keypoints and descriptor bytes."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from send_slam_b200 import orbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_features_messages_reach_the_tracking_containers(tmp_path):
    msgpack = pytest.importorskip("msgpack")
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "features_loop_driver")
    libdir = os.path.join(ROOT, "send_slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fsanitize=alignment,undefined", "-fno-sanitize-recover=all", "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "harness", "features_loop_driver.cc"), "-L", libdir, "-lorbx",
                           "-Wl,-rpath," + libdir, "-o", exe])
    rng = np.random.default_rng(3)
    msgs, want = [], []
    for k, n in enumerate((0, 1, 257, 1004)):
        kps = np.zeros(n, orbx.KP_DTYPE)
        kps["x"], kps["y"] = rng.uniform(19, 600, n).astype(np.float32), rng.uniform(19, 400, n).astype(np.float32)
        kps["size"], kps["angle"], kps["response"] = 31.0, rng.uniform(0, 360, n).astype(np.float32), rng.integers(8, 200, n).astype(np.float32)
        kps["octave"], kps["class_id"] = rng.integers(0, 8, n), -1
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if k % 2 == 0:      # the library's writer (what the NIF side sends), already framed
            framed = orbx.wire_pack_features(0.5 + k, 3 + k, 640, 480, n // 2, kps, desc, framed=True)
        else:               # an independent MessagePack implementation, odd string lengths push the binaries to odd offsets
            body = msgpack.packb({"type": "features", "pad": "x" * (k + 2), "camera_id": 3 + k, "timestamp": 0.5 + k, "width": 640, "height": 480,
                                  "mono_index": n // 2, "n": n, "keypoints": kps.tobytes(), "descriptors": desc.tobytes()}, use_bin_type=True)
            framed = struct.pack(">I", len(body)) + body
        msgs.append(framed)
        sd = 0
        for b in desc.tobytes():
            sd = (sd * 131 + b) & 0xFFFFFFFFFFFFFFFF
        want.append((3 + k, 0.5 + k, n // 2, n, float(kps["x"].astype(np.float64).sum()), float(kps["y"].astype(np.float64).sum()),
                     float(kps["angle"].astype(np.float64).sum()), int(kps["octave"].sum()), sd))
    # what the branch skips: a frame message, garbage, a features message without camera id
    bad = [msgpack.packb({"type": "frame", "camera_id": 1}, use_bin_type=True), b"\xc1\xc1\xc1",
           msgpack.packb({"type": "features", "camera_id": 0, "timestamp": 1.0, "width": 1, "height": 1, "mono_index": 0, "n": 0, "keypoints": b"",
                          "descriptors": b""}, use_bin_type=True)]
    path = tmp_path / "stream.bin"
    with open(path, "wb") as f:
        for m in msgs:
            f.write(m)
        for b in bad:
            f.write(struct.pack(">I", len(b)) + b)
    out = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-800:]
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "done handled=4 skipped=3", lines[-3:]
    got = [l for l in lines if l.startswith("features ")]
    for l, w in zip(got, want):
        f = dict(tok.split("=") for tok in l.split()[1:] if "=" in tok)
        assert int(f["camera"]) == w[0] and abs(float(f["t"]) - w[1]) < 1e-9 and int(f["mono"]) == w[2] and int(f["n"]) == w[3]
        assert abs(float(f["sx"]) - w[4]) < 1e-2 and abs(float(f["sy"]) - w[5]) < 1e-2 and abs(float(f["sa"]) - w[6]) < 1e-2
        assert int(f["so"]) == w[7] and int(f["sd"]) == w[8]
