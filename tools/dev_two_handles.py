#!/usr/bin/env python
"""Device-resident throughput with one handle vs several handles used alternately on their own streams (cross-step overlap).
Usage: [ORBX_FAST_CTAS=5] python tools/dev_two_handles.py [nhandles]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from send_slam_b200 import orbx, synth

NH = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, W, H, RING = 64, 640, 480, 8
dev = torch.device("cuda", 0)
d_in = [torch.from_numpy(np.stack([synth.textured_frame(64 * r + i, W, H) for i in range(B)])).to(dev) for r in range(RING)]
lanes = []
for k in range(NH):
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
    cap = ex.capacity
    st = torch.cuda.Stream(device=dev)
    ex.set_stream(st.cuda_stream)
    bufs = (torch.zeros((B, cap, 7), dtype=torch.float32, device=dev), torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev),
            torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev))
    lanes.append((ex, st, bufs))

def step(i):
    ex, st, (kp, de, n, mo) = lanes[i % NH]
    t = d_in[i % RING]
    ex.extract_batch_device(t.data_ptr(), H * W, B, W, H, W, kp.data_ptr(), de.data_ptr(), ex.capacity, n.data_ptr(), mo.data_ptr())

for i in range(3 * RING * NH):
    step(i)
torch.cuda.synchronize()
K = 80
t0 = time.perf_counter()
for i in range(K):
    step(i)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"handles={NH} fast_ctas={os.environ.get('ORBX_FAST_CTAS', '-')} split={os.environ.get('ORBX_DEV_SPLIT', '-')}: {K * B / dt:.0f} frames/s, {1e3 * dt / K:.4f} ms/step, n={int(lanes[0][2][2].sum())}")
