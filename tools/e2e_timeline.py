#!/usr/bin/env python
"""CUPTI timeline (via torch.profiler) of one orbx_extract_batch call with pinned host buffers: prints every memcpy /
kernel with start offset and duration.  Usage: [ORBX_CHUNK=..] python tools/e2e_timeline.py [batch]"""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from send_slam_b200 import orbx, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = 640, 480
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
cap = ex.capacity
pin = torch.from_numpy(np.stack([synth.textured_frame(i, W, H) for i in range(B)])).pin_memory()
pk = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
pd = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
out = (pk.numpy().view(orbx.KP_DTYPE).reshape(B, cap), pd.numpy())
for i in range(4):
    ex.extract_batch(pin.numpy(), out=out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ex.extract_batch(pin.numpy(), out=out)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:8.1f}  stream {e['args'].get('stream', '?'):>3}  {e['name'][:60]}")
print("span us:", ev[-1]["ts"] + ev[-1]["dur"] - t0)
