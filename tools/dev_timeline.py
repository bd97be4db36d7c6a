#!/usr/bin/env python
"""CUPTI timeline (torch.profiler) of one device-resident orbx_extract_batch_device call (64 x 640x480), graph replay included.
Usage: python tools/dev_timeline.py"""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from send_slam_b200 import orbx, synth

B, W, H = 64, 640, 480
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
cap = ex.capacity
stream = torch.cuda.Stream()
ex.set_stream(stream.cuda_stream)
d_in = torch.from_numpy(np.stack([synth.textured_frame(i, W, H) for i in range(B)])).cuda()
d_kp = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
d_n = torch.zeros(B, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(B, dtype=torch.int32, device="cuda")
def step():
    ex.extract_batch_device(d_in.data_ptr(), W * H, B, W, H, W, d_kp.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), d_mono.data_ptr())
for _ in range(5):
    step()
ex.sync()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); step()
    ex.sync()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:8.1f}  stream {e['args'].get('stream', '?'):>3}  {e['name'][:48]}")
print("span us:", ev[-1]["ts"] + ev[-1]["dur"] - t0)
