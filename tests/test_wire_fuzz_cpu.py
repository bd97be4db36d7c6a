"""CPU: the parsers of untrusted bytes (MessagePack frame / features messages, binary-PNM header) built on their own with
AddressSanitizer + UBSan and fed mutated messages: no out-of-bounds read, no undefined behaviour, and every accepted input
yields pointers inside the input (tests/harness/wire_fuzz_driver.c)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mutations(rng, seeds, count):
    out = list(seeds)
    for _ in range(count):
        b = bytearray(seeds[int(rng.integers(0, len(seeds)))])
        for _ in range(int(rng.integers(1, 6))):
            op = int(rng.integers(0, 6))
            if not b:
                b = bytearray(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes())
            pos = int(rng.integers(0, len(b)))
            if op == 0:
                b[pos] ^= 1 << int(rng.integers(0, 8))                       # bit flip
            elif op == 1:
                b[pos] = int(rng.choice([0x00, 0x7F, 0x80, 0xC1, 0xC6, 0xDB, 0xDD, 0xDF, 0xFF, ord("#"), ord(" ")]))   # hostile tags
            elif op == 2:
                del b[pos:]                                                  # truncate
            elif op == 3:
                b[pos:pos] = rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8).tobytes()                 # insert
            elif op == 4:
                del b[pos:pos + int(rng.integers(1, 9))]                     # delete
            else:
                b[pos:pos + 4] = struct.pack(">I", int(rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, len(b), len(b) + 1])))   # hostile lengths
        out.append(bytes(b))
    return out


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_parsers_under_asan(tmp_path):
    msgpack = pytest.importorskip("msgpack")
    exe = str(tmp_path / "wire_fuzz")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", os.path.join(ROOT, "include"),
           "-x", "c++", os.path.join(ROOT, "send_slam_b200", "csrc", "orbx_wire.cpp"), "-x", "c", os.path.join(ROOT, "tests", "harness", "wire_fuzz_driver.c"),
           "-o", exe]
    try:
        subprocess.check_call(cmd)
    except subprocess.CalledProcessError:
        pytest.skip("sanitizer runtime not available to this g++")
    rng = np.random.default_rng(11)
    ppm = b"P6\n# c\n6 4\n255\n" + bytes(range(72))
    frame = {"type": "frame", "camera_id": 1, "encoding": "ppm", "timestamp": 1.5, "width": 6, "height": 4, "channels": 3, "frame": ppm,
             "calibration": {"fx": 1.0, "d": [1, 2, {"k": b"zz"}]}, "ext": msgpack.ExtType(3, b"abcd")}
    kp = rng.integers(0, 256, 5 * 28, dtype=np.uint8).tobytes()
    de = rng.integers(0, 256, 5 * 32, dtype=np.uint8).tobytes()
    feat = {"type": "features", "camera_id": 2, "timestamp": 2.5, "width": 64, "height": 48, "mono_index": 3, "n": 5, "keypoints": kp, "descriptors": de}
    seeds = [msgpack.packb(frame, use_bin_type=True), msgpack.packb(feat, use_bin_type=True), ppm, b"P5\n3 3\n255\n" + bytes(9), b"",
             msgpack.packb({"type": "frame", "frame": b"", "timestamp": 0, "camera_id": -3}, use_bin_type=True)]
    cases = _mutations(rng, seeds, 30000)
    path = tmp_path / "cases.bin"
    with open(path, "wb") as f:
        for c in cases:
            f.write(struct.pack("<I", len(c)))
            f.write(c)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    total, ok_frame, ok_feat, ok_pnm = (int(v) for v in r.stdout.split())
    assert total == len(cases)
    assert ok_frame >= 3 and ok_feat >= 1 and ok_pnm >= 2            # the unmutated seeds are accepted ...
    assert ok_frame < total // 2 and ok_pnm < total // 2               # ... and most mutants are not
