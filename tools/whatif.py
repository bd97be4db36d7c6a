#!/usr/bin/env python
"""What-if timing of the device-resident step (four lanes, as bench.py's headline) with stages switched off (ORBX_SKIP_STAGES bit mask:
1 pyramid, 2 blur, 4 FAST, 8 quadtree + slots, 16 descriptors; results are meaningless then).  Shows what a stage costs the overlapped
step, as opposed to its own duration.  Usage: ORBX_SKIP_STAGES=1 python tools/whatif.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from send_slam_b200 import orbx, synth
W, H, B, RING, NL, STEPS = 640, 480, 64, 8, 4, 200
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
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = lanes[0][1]
ev0.record(main)
for i in range(STEPS):
    step(i)
for e, s, *_ in lanes[1:]:
    ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
ev1.record(main)
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / STEPS
print(json.dumps({"skip": os.environ.get("ORBX_SKIP_STAGES", "0"), "ms_per_step": round(ms, 4), "frames_per_s": round(B / ms * 1e3)}), flush=True)
