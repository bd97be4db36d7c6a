#!/usr/bin/env python
"""Platform ceiling of the frame upload at N ranks: every rank copies 64 x 640x480 batches from page-locked host memory to its GPU at
the same time, no kernels.  Compares ordinary page-locked memory with write-combined memory (orbx_host_alloc) and, per rank, cores bound
to the GPU's NUMA node or not.  Launch: python tools/h2d_probe.py  |  python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from send_slam_b200 import orbx

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, H, W, RING = 64, 480, 640, 8
res = {}
for kind in ("pinned", "write_combined"):
    bufs = [orbx.PinnedArray((B, H, W), np.uint8, write_combined=(kind == "write_combined")) for _ in range(RING)]
    for b in bufs:
        b.array[:] = 7
    d_tmp = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
    st = [torch.cuda.Stream(), torch.cuda.Stream()]
    def copies(n):
        for i in range(n):
            with torch.cuda.stream(st[i & 1]):
                # cudaMemcpyAsync through torch needs a tensor: wrap the page-locked numpy view (no copy)
                d_tmp[i & 1].copy_(torch.from_numpy(bufs[i % RING].array), non_blocking=True)
        torch.cuda.synchronize()
    copies(8)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    copies(200)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    res[kind] = {"frames_per_s": world * 200 * B / dt, "gb_per_s_total": world * 200 * B * H * W / dt / 1e9}
    for b in bufs:
        b.close()
if rank == 0:
    print(json.dumps({"ranks": world, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
