"""Generates the golden fixtures under tests/golden/ from the cv2-backed restatement (oracle/orb_cv2.py).

Run here (needs cv2, which carries the REAL OpenCV resize / FAST / GaussianBlur / fastAtan2 / BFMatcher code):
    python tests/golden/make_golden.py
The reference's own ORB sources are not under /root/reference, so these vectors are the pin the
reference cannot provide: OpenCV primitives are the real thing, the ORB-SLAM3 control logic is the restatement.
Each .npz holds: frame recipe (seed, size, kind), extractor parameters, final keypoints + descriptors + monoIndex,
per-level SHA-256 of pyramid and blurred planes, per-level candidate / selected counts.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import orb_cv2 as oc  # noqa: E402
from send_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, seed, w, h, kind, nfeatures, lap
    ("tex_320x240_nf500", 11, 320, 240, "textured", 500, (0, 1000)),
    ("mixed_320x240_nf500", 12, 320, 240, "mixed", 500, (0, 1000)),
    ("sparse_320x240_nf500", 13, 320, 240, "sparse", 500, (0, 1000)),
    ("lowc_376x240_nf600", 14, 376, 240, "lowcontrast", 600, (0, 1000)),
    ("tex_640x480_nf1000", 15, 640, 480, "textured", 1000, (0, 1000)),
    ("tex_1280x720_nf1250_lap", 16, 1280, 720, "textured", 1250, (0, 1000)),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name, seed, w, h, kind, nf, lap in CASES:
        frame = synth.textured_frame(seed, w, h, kind)
        p = oc.Params(nf, 1.2, 8, 20, 7)
        r = oc.extract(frame, p, lap=lap, keep_stages=True)
        st = r["stages"]
        kps = r["kps"]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            seed=seed, width=w, height=h, kind=kind, nfeatures=nf, lap=np.array(lap), frame_sha=sha(frame),
            kps=kps.astype(np.float32), desc=r["desc"], mono_index=r["mono_index"],
            level_sha=np.array([sha(l) for l in st["levels"]]),
            blur_sha=np.array([sha(b) if b is not None else "" for b in st["blur"]]),
            ncand=np.array([len(c) for c in st["cand"]]), nsel=np.array([len(s) for s in st["sel"]]),
            cv2_version=oc.cv2.__version__,
        )
        print(name, "keypoints", len(kps), "mono", r["mono_index"], "cand", [len(c) for c in st["cand"]])
    # Hamming kNN (k=2) golden: cv2.BFMatcher on a small random database
    db = synth.descriptor_db(4096, seed=21)
    q, src = synth.queries_from_db(db, 128, seed=22)
    db[100] = db[7]            # force exact ties to pin the lowest-index rule
    db[2000] = db[7]
    idx, dist = oc.knn2_bf(q, db)
    np.savez_compressed(os.path.join(HERE, "knn2_4096x128.npz"), db_seed=21, q_seed=22, idx=idx, dist=dist, src=src)
    print("knn2", idx[:3].tolist(), dist[:3].tolist())


if __name__ == "__main__":
    main()
