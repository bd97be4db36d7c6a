#!/usr/bin/env python
"""Per-stage device times of one extraction call shape (CUDA events inside the library), no torch needed.
Usage: python tools/stage_times.py [W H NFEAT BATCH REPS]   -> one JSON line.  Environment knobs (ORBX_FAST_V, ORBX_FAST_MIX, ...)
are read by the library once per process, so a sweep runs this script once per setting (tools/stage_sweep.py)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from send_slam_b200 import orbx, synth

def main():
    a = [int(x) for x in sys.argv[1:6]] + [None] * 5
    W, H, NF, B, REPS = a[0] or 640, a[1] or 480, a[2] or 1000, a[3] or 64, a[4] or 6
    frames = np.stack([synth.textured_frame(100 + i, W, H) for i in range(min(B, 16))])
    frames = frames[np.arange(B) % len(frames)].copy()
    ex = orbx.ORBextractor(NF, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
    ex.extract_batch(frames)
    ex.set_profiling(True)
    acc, n_kp = {}, 0
    for r in range(REPS + 1):
        mono, n, kps, desc = ex.extract_batch(frames)
        if r == 0:
            continue
        for k, v in ex.stage_times_ms().items():
            acc[k] = acc.get(k, 0.0) + v / REPS
        n_kp = int(n.sum())
    import hashlib
    hsh = hashlib.sha1(b"".join(desc[f, :n[f]].tobytes() + kps[f, :n[f]].tobytes() for f in range(B))).hexdigest()[:12]
    env = {k: v for k, v in os.environ.items() if k.startswith("ORBX_")}
    print(json.dumps({"shape": [W, H, NF, B], "env": env, "stage_us": {k: round(1e3 * v, 1) for k, v in acc.items()}, "keypoints": n_kp,
                      "result_sha1": hsh}), flush=True)

main()
