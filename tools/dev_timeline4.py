#!/usr/bin/env python
"""CUPTI timeline (torch.profiler) of the bench's steady state: four handles / streams, device-resident 64 x 640x480 batches, graph replay.
Prints per-kernel busy time, the union of busy intervals (time with at least one kernel running) and the concurrency histogram over a window of
8 steps.  Usage: python tools/dev_timeline4.py [lanes]"""
import json, os, sys, tempfile, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from send_slam_b200 import orbx, synth
W, H, B, RING = 640, 480, 64, 8
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
base = np.stack([synth.textured_frame(100 + i, W, H) for i in range(16)])
d_in = [torch.from_numpy(np.roll(base[np.arange(B) % 16], 5 * r, axis=2)).to(dev) for r in range(RING)]
lanes = []
for _ in range(NL):
    e = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
    cap = e.capacity
    s = torch.cuda.Stream()
    e.set_stream(s.cuda_stream)
    lanes.append((e, s, torch.zeros((B, cap, 7), device=dev), torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev),
                  torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)))
def step(i):
    e, s, k, d, n, m = lanes[i % NL]
    e.extract_batch_device(d_in[i % RING].data_ptr(), H * W, B, W, H, W, k.data_ptr(), d.data_ptr(), cap, n.data_ptr(), m.data_ptr())
for i in range(3 * RING + NL):
    step(i)
torch.cuda.synchronize()
NS = 16
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(NS):
        step(i)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
span = t1 - t0
busy = collections.Counter(); cnt = collections.Counter()
for e in ev:
    name = e["name"].split("(")[0].replace("void ", "").replace("orbx::", "")[:28]
    busy[name] += e["dur"]; cnt[name] += 1
print(f"{NL} lanes, {NS} steps: span {span:.1f} us = {span / NS:.1f} us per step (under the profiler)")
for k, v in busy.most_common():
    print(f"  {k:28s} launches {cnt[k]:4d}  busy {v:9.1f} us  = {v / NS:7.1f} us per step, avg {v / cnt[k]:6.1f}")
# concurrency: sweep
pts = []
for e in ev:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
hist = collections.Counter(); cur = 0; last = pts[0][0]
for t, d in pts:
    hist[cur] += t - last; last = t; cur += d
print("kernels running at once -> share of the span:", {k: round(v / span, 3) for k, v in sorted(hist.items())})
# gaps where nothing runs
print("idle share:", round(hist[0] / span, 3))
