#!/usr/bin/env python
"""pairs/s of the three kNN backends on one GPU: 2000 queries x N rows (default 1 M and 10 M), device resident, checked against each other."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from send_slam_b200 import orbx, synth
dev = torch.device("cuda", 0)
for rows in [int(x) for x in (sys.argv[1:] or ["1000000", "10000000"])]:
    g = torch.Generator(device=dev); g.manual_seed(5)
    d_db = torch.randint(0, 256, (rows, 32), dtype=torch.uint8, device=dev, generator=g)
    src = torch.randint(0, rows, (2000,), device=dev, generator=g)
    d_q = d_db[src].clone()
    d_q[:, 3] ^= 0x5A
    ref = None
    for name, be in (("popc", 0), ("tensor_i8", 1), ("tensor_fp4", 2)):
        ix = orbx.Knn2Index(device=0, device_ptr=d_db.data_ptr(), nrows=rows)
        ix.set_backend(be)
        st = torch.cuda.Stream(); ix.set_stream(st.cuda_stream)
        out = torch.zeros((2000, 2), dtype=torch.int64, device=dev)
        for _ in range(3):
            ix.query_device(d_q.data_ptr(), 2000, out.data_ptr())
        ix.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if be == 0 else 10
        e0.record(st)
        for _ in range(n):
            ix.query_device(d_q.data_ptr(), 2000, out.data_ptr())
        e1.record(st); ix.sync()
        t = e0.elapsed_time(e1) * 1e-3 / n
        same = None if ref is None else bool(torch.equal(out, ref))
        if ref is None:
            ref = out.clone()
        print(json.dumps({"rows": rows, "backend": name, "ms": round(1e3 * t, 4), "T_pairs_per_s": round(2000 * rows / t / 1e12, 3), "equal_to_popc": same}), flush=True)
        ix.close()
    del d_db
    torch.cuda.empty_cache()
