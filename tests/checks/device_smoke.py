#!/usr/bin/env python
"""A small pass over every device path, each result checked against the oracle: single frame, a batch through submit/collect, a PPM
wire frame, windowed matching and kNN on both routes.  Small enough to run under a sanitizer or a debugger where those are available
(compute-sanitizer is closed on the pool this repository was developed on).  Usage (on a B200): python tests/checks/device_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from send_slam_b200 import orbx, synth
from oracle import oracle_lib as ol

w, h, nf, B = 320, 240, 500, 4
o = ol.Oracle(nf)
frames = np.stack([synth.textured_frame(50 + i, w, h, "textured" if i % 2 == 0 else "mixed") for i in range(B)])
e = orbx.ORBextractor(nf, 1.2, 8, 20, 7, device=0, max_width=w, max_height=h, max_batch=B)
mono, kps, desc = e(frames[0])
k_o, d_o, m_o = o.extract(frames[0])
assert mono == m_o and np.array_equal(desc, d_o)
e.extract_batch_submit(frames)
mb, nb, kb, db = e.extract_batch_collect()
for i in range(B):
    k_i, d_i, m_i = o.extract(frames[i])
    assert nb[i] == len(k_i) and np.array_equal(db[i, :nb[i]], d_i)
col = np.stack([frames[0], frames[1], frames[2]], -1)
ppm = b"P6\n%d %d\n255\n" % (w, h) + col.tobytes()
mp, kp_, dp, size = e.extract_pnm(ppm, camera_rgb=True)
assert np.array_equal(dp, o.extract(ol.gray(col, 2))[1])
# windowed matching against a shifted copy
f2 = synth.shifted_frame(frames[0], 3, -2)
mono2, kps2, desc2 = e(f2)
quvr = np.stack([kps["x"] + 3, kps["y"] - 2, 15.0 * np.float32(1.2) ** kps["octave"]], 1).astype(np.float32)
qlev = np.stack([kps["octave"] - 1, kps["octave"] + 1], 1).astype(np.int32)
m = orbx.ORBmatcher(extractor=e)
got = m.SearchInWindows(desc, quvr, qlev, kps2, desc2, np.array([0, 0, w, h], np.float32))
want = ol.match_windowed(desc, quvr, qlev, kps2, desc2, np.array([0, 0, w, h], np.float32))
assert all(np.array_equal(a, b) for a, b in zip(got, want))
# kNN, both routes
dbase = synth.descriptor_db(9000, seed=5)
q, _ = synth.queries_from_db(dbase, 130, seed=6)
idx_o, dist_o = ol.knn2(q, dbase)
ix = orbx.Knn2Index(dbase, device=0)
for backend in (orbx.Knn2Index.TENSOR, orbx.Knn2Index.POPC):
    ix.set_backend(backend)
    idx, dist = ix.knnMatch(q)
    assert np.array_equal(idx, idx_o) and np.array_equal(dist, dist_o), backend
e.close()
print("sanitize smoke ok")
