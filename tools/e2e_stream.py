#!/usr/bin/env python
"""Streaming e2e figure of bench.py in isolation: orbx_extract_batch_submit / _collect from one host thread over L handles, pinned
buffers, 64 x 640x480 per batch; prints frames/s and the plain-copy ceiling measured in the same process.
Knobs (environment, read once per process): ORBX_E2E_LANES, ORBX_CHUNK, ORBX_SUB, ORBX_STREAMS.  Usage: python tools/e2e_stream.py [steps]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from send_slam_b200 import orbx, synth

W, H, B, RING = 640, 480, 64, 8
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
NL = int(os.environ.get("ORBX_E2E_LANES", "4"))
base = np.stack([synth.textured_frame(100 + i, W, H) for i in range(16)])
pin = [torch.from_numpy(np.roll(base[np.arange(B) % 16], 5 * r, axis=2)).pin_memory() for r in range(RING)]
lanes = []
for _ in range(NL):
    e = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=B)
    cap = e.capacity
    pk = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
    pd = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    lanes.append((e, (pk.numpy().view(orbx.KP_DTYPE).reshape(B, cap), pd.numpy()), pk, pd))


host = {"submit": 0.0, "collect": 0.0, "n": 0}


def run(n):
    busy = [False] * NL
    for i in range(n):
        k = i % NL
        e, o, _, _ = lanes[k]
        ta = time.perf_counter()
        if busy[k]:
            e.extract_batch_collect()
        tb = time.perf_counter()
        e.extract_batch_submit(pin[i % RING].numpy(), out=o)
        tc = time.perf_counter()
        host["collect"] += tb - ta; host["submit"] += tc - tb; host["n"] += 1
        busy[k] = True
    for j in range(NL):
        k = (n + j) % NL
        if busy[k]:
            lanes[k][0].extract_batch_collect()


run(NL * RING + NL)
torch.cuda.synchronize()
host.update(submit=0.0, collect=0.0, n=0)
t0 = time.perf_counter()
run(steps)
dt = time.perf_counter() - t0
fps = steps * B / dt
# ceiling: plain copies of the same buffers
dev = torch.device("cuda", 0)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
d_tmp = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
def copies(n):
    for i in range(n):
        with torch.cuda.stream(sa if i & 1 else sb):
            d_tmp[i & 1].copy_(pin[i % RING], non_blocking=True)
    torch.cuda.synchronize()
copies(8)
t0 = time.perf_counter()
copies(160)
ceil = 160 * B / (time.perf_counter() - t0)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("ORBX_")}, "lanes": NL, "e2e_fps": round(fps), "h2d_ceiling_fps": round(ceil),
                  "ratio": round(fps / ceil, 3),
                  "host_us_per_step": {"submit": round(1e6 * host["submit"] / max(host["n"], 1), 1), "collect_incl_wait": round(1e6 * host["collect"] / max(host["n"], 1), 1)},
                  "step_us": round(1e6 * dt / steps, 1)}), flush=True)
