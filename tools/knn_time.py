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

# ---- sustained leg: the FP4 and int8 backends back to back for ~2 s each on a 10 M-row shard, SM clock and board power sampled through NVML
if os.environ.get("KNN_SUSTAIN", "1") == "1":
    import threading, time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    rows = 10_000_000
    g = torch.Generator(device=dev); g.manual_seed(5)
    d_db = torch.randint(0, 256, (rows, 32), dtype=torch.uint8, device=dev, generator=g)
    d_q = d_db[torch.randint(0, rows, (2000,), device=dev, generator=g)].clone()
    for name, be in (("tensor_i8", 1), ("tensor_fp4", 2)):
        ix = orbx.Knn2Index(device=0, device_ptr=d_db.data_ptr(), nrows=rows)
        ix.set_backend(be)
        st = torch.cuda.Stream(); ix.set_stream(st.cuda_stream)
        out = torch.zeros((2000, 2), dtype=torch.int64, device=dev)
        ix.query_device(d_q.data_ptr(), 2000, out.data_ptr()); ix.sync()
        samples, stop = [], threading.Event()
        def sampler():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.02)
        th = threading.Thread(target=sampler); th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 700
        e0.record(st)
        for _ in range(n):
            ix.query_device(d_q.data_ptr(), 2000, out.data_ptr())
        e1.record(st); ix.sync()
        stop.set(); th.join()
        t = e0.elapsed_time(e1) * 1e-3 / n
        mid = samples[len(samples) // 4:]
        clk = sorted(s[0] for s in mid); pw = sorted(s[1] for s in mid)
        print(json.dumps({"sustained": name, "rows": rows, "seconds": round(t * n, 2), "T_pairs_per_s": round(2000 * rows / t / 1e12, 3),
                          "sm_mhz_median": clk[len(clk) // 2], "sm_mhz_min": clk[0], "power_w_median": pw[len(pw) // 2], "power_w_max": pw[-1]}), flush=True)
        ix.close()
