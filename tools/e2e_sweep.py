#!/usr/bin/env python
"""e2e timing of orbx_extract_batch (pinned host buffers) for the ORBX_CHUNK given in the environment; checks the result
against the device-resident entry point.  Usage: ORBX_CHUNK=16 python tools/e2e_sweep.py [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from send_slam_b200 import orbx, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = 640, 480
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
cap = ex.capacity
batches = [np.stack([synth.textured_frame(100 * r + i, W, H) for i in range(B)]) for r in range(4)]
pin = [torch.from_numpy(b).pin_memory() for b in batches]
pk = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
pd = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
out = (pk.numpy().view(orbx.KP_DTYPE).reshape(B, cap), pd.numpy())
for i in range(3):
    ex.extract_batch(pin[i % 4].numpy(), out=out)
torch.cuda.synchronize()
n_it = 30
t0 = time.perf_counter()
for i in range(n_it):
    mono, n, kps, desc = ex.extract_batch(pin[i % 4].numpy(), out=out)
dt = (time.perf_counter() - t0) / n_it
# pageable in/out
t0 = time.perf_counter()
for i in range(10):
    mono2, n2, kps2, desc2 = ex.extract_batch(batches[i % 4])
dtp = (time.perf_counter() - t0) / 10
# parity of the pipelined result with a single-range run of the last batch (profiling mode forces one range)
last = (n_it - 1) % 4
k1, d1, nn = kps.copy(), desc.copy(), n.copy()
ex.set_profiling(True)
mono3, n3, kps3, desc3 = ex.extract_batch(batches[last])
ex.set_profiling(False)
ok = np.array_equal(nn, n3) and all(np.array_equal(k1[i, :nn[i]], kps3[i, :nn[i]]) and np.array_equal(d1[i, :nn[i]], desc3[i, :nn[i]]) for i in range(B))
print(f"chunk={os.environ.get('ORBX_CHUNK', 'default')} batch={B} pinned {1e3 * dt:.3f} ms/step {B / dt:.0f} frames/s | pageable {1e3 * dtp:.3f} ms/step {B / dtp:.0f} frames/s | identical={ok}")
