"""Generates tests/golden/pnm_cases.npz: what the REAL cv::imdecode (cv2) returns for binary-PNM byte strings as the reference
puts them on the wire (send_slam/lib/send_slam/slam_handler.ex:275-277 -> orbslam3_mono_networked.cc:546), and what
cv::cvtColor makes of the decoded Mat for both values of Camera.RGB.  Run here (needs cv2):
    python tests/golden/make_golden_pnm.py
"""
from __future__ import annotations

import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from send_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def header_cases():
    rng = np.random.default_rng(5)
    px = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8).tobytes()
    return [
        b"P6\n7 5\n255\n" + px, b"P6\n# made by a test\n7 5\n255\n" + px, b"P6 7 5 255 " + px, b"P6\n7\n#x\r5\n255\t" + px,
        b"P6\n7 5 #c\n#d\n\n 255\n" + px + b"trailing", b"P6\n7 5\n100\n" + px, b"P6\n7 5\n255\n\n" + px, b"P6\n7 5\n255\r\n" + px + b"x",
        b"P6\n7 5\n255\n#" + px, b"P5\n7 5\n255\n" + px[:35], b"P5\n7 5\n255\n" + px[:34], b"P6\n7 5\n255\n" + px[:-1],
        b"P6\n7 5\n255\n" + px[:50], b"P6\n7 5\n255" + px, b"P6\n0 5\n255\n" + px, b"P6\n7 5\n0\n" + px, b"P6x7 5\n255\n" + px,
        b"P6\n+7 5\n255\n" + px, b"P6\n7 5\n99999999999\n" + px, b"P6\n7 5\n65536\n" + px + px, b"P7\n7 5\n255\n" + px, b"P", b"P6", b"P6\n7",
        b"Q6\n7 5\n255\n" + px, b"P6\n7 5\n255\n" + px + px,
        # decodable by OpenCV but not CV_8U binary: the library refuses these instead of returning something else
        b"P6\n7 5\n65535\n" + px + px, b"P4\n8 2\n\xff\x00", b"P2\n2 2\n255\n1 2 3 4 ",
    ]


def main():
    cases = header_cases()
    verdict, shape, digest = [], [], []
    for c in cases:
        try:
            m = cv2.imdecode(np.frombuffer(c, np.uint8), cv2.IMREAD_UNCHANGED)
        except cv2.error:
            m = None
        if m is None:
            verdict.append(0); shape.append((0, 0, 0)); digest.append("")
        else:
            verdict.append(1 if m.dtype == np.uint8 and c[1:2] in (b"5", b"6") else 2)     # 2: decodes, outside the 8-bit binary forms
            shape.append((m.shape[0], m.shape[1], 1 if m.ndim == 2 else m.shape[2])); digest.append(sha(m))
    # one wire frame: colour image -> imencode(".ppm") (what Evision.imencode does) -> imdecode -> cvtColor, both Camera.RGB values
    w, h = 320, 240
    bgr = np.stack([synth.textured_frame(40 + k, w, h) for k in range(3)], axis=2)
    ok, enc = cv2.imencode(".ppm", bgr)
    assert ok
    wire = enc.tobytes()
    dec = cv2.imdecode(np.frombuffer(wire, np.uint8), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(dec, bgr)
    gray_rgb1 = cv2.cvtColor(dec, cv2.COLOR_RGB2GRAY)          # Tracking::GrabImageMonocular with mbRGB
    gray_rgb0 = cv2.cvtColor(dec, cv2.COLOR_BGR2GRAY)
    ok, encg = cv2.imencode(".pgm", bgr[:, :, 1])
    np.savez_compressed(os.path.join(HERE, "pnm_cases.npz"), n=len(cases), blob=np.frombuffer(b"".join(cases), np.uint8),
                        lengths=np.array([len(c) for c in cases]), verdict=np.array(verdict), shape=np.array(shape), digest=np.array(digest),
                        wire_seeds=np.array([40, 41, 42]), wire_size=np.array([w, h]), wire_header=np.frombuffer(wire[:15], np.uint8),
                        wire_sha=sha(np.frombuffer(wire, np.uint8)), decoded_sha=sha(dec), gray_rgb1_sha=sha(gray_rgb1), gray_rgb0_sha=sha(gray_rgb0),
                        pgm_header=np.frombuffer(encg.tobytes()[:15], np.uint8), cv2_version=cv2.__version__)
    print(len(cases), "cases;", sum(v == 1 for v in verdict), "decode;", "wire", len(wire), "bytes, header", wire[:15])


if __name__ == "__main__":
    main()
