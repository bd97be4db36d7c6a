#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): (1) row-sharded brute-force kNN, top-2 merged with one NCCL
all-gather, compared on rank 0 with the CPU oracle on a bounded database; (2) frame-sharded extraction, every rank's
shard compared with the oracle on a few frames; (3) timing of the config-4 shape (2000 queries vs rows-per-GPU x world).
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/checks/multigpu_check.py [rows_per_gpu]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from send_slam_b200 import orbx, sharded, synth

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = {"world": world}

# (1) parity on a database the oracle finishes in seconds: 200 003 rows (ragged shards), duplicates across shards
N, NQ = 200_003, 500
db = synth.descriptor_db(N, seed=4242)
db[N - 5] = db[17]                      # a tie between the first and the last shard: lowest global row must win
q, _ = synth.queries_from_db(db, NQ, seed=43)
q[0] = db[17]
lo, hi = sharded.row_shard(N, world, rank)
index = orbx.Knn2Index(db[lo:hi], device=local, row_offset=lo)
d_q = torch.from_numpy(q).to(dev)
from oracle import oracle_lib as ol
idx_o, dist_o = ol.knn2(q, db) if rank == 0 else (None, None)
for trial, backend in enumerate((orbx.Knn2Index.TENSOR_FP4, orbx.Knn2Index.POPC, orbx.Knn2Index.TENSOR, orbx.Knn2Index.TENSOR_FP4)):
    index.set_backend(backend)
    packed = sharded.knn2_sharded(index, d_q).cpu().numpy()
    if rank == 0:
        got_idx, got_dist = (packed & 0xFFFFFFFF).astype(np.int64), (packed >> 32).astype(np.int64)
        bad = np.nonzero((got_idx != idx_o).any(1) | (got_dist != dist_o).any(1))[0]
        res[f"knn_parity_trial{trial}_backend{backend}"] = len(bad) == 0
        if len(bad):
            res[f"bad_trial{trial}"] = [[int(b), got_idx[b].tolist(), got_dist[b].tolist(), idx_o[b].tolist(), dist_o[b].tolist()] for b in bad[:4]]
            res[f"nbad_trial{trial}"] = int(len(bad))
        res["tie_rows"] = got_idx[0].tolist()
index.close()

# (2) frame shards: 16 frames over the ranks, each rank checks its own shard against the oracle
B = 16
f_lo, f_hi = sharded.frame_shard(B, world, rank)
frames = np.stack([synth.textured_frame(900 + i, 640, 480) for i in range(f_lo, f_hi)])
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=local, max_width=640, max_height=480, max_batch=max(1, f_hi - f_lo))
mono, n, kps, desc = ex.extract_batch(frames)
from oracle import oracle_lib as ol
o = ol.Oracle(1000, 1.2, 8, 20, 7)
ok = True
for i in range(len(frames)):
    k_o, d_o, m_o = o.extract(frames[i])
    ok = ok and n[i] == len(k_o) and np.array_equal(desc[i, :n[i]], d_o) and mono[i] == m_o
t = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
res["extract_shards_ok"] = bool(t.item())
ex.close()

# (3) config-4 shape: 2000 queries vs rows_per_gpu x world rows, k = 2, top-2 merge over NCCL
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
dbs = synth.descriptor_db(rows, seed=1234 + rank)
qs, _ = synth.queries_from_db(dbs, 2000, seed=99)
d_db = torch.from_numpy(dbs).to(dev)
d_qs = torch.from_numpy(qs).to(dev)
if world > 1:
    dist.broadcast(d_qs, 0)
index = orbx.Knn2Index(device=local, device_ptr=d_db.data_ptr(), nrows=rows, row_offset=rank * rows)
for _ in range(3):
    out = sharded.knn2_sharded(index, d_qs)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
ev0.record()
for _ in range(steps):
    out = sharded.knn2_sharded(index, d_qs)
ev1.record()
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
res["knn_cfg4"] = {"queries": 2000, "rows_total": rows * world, "ms_per_query_batch": float(ms.item()),
                   "pairs_per_s": 2000 * rows * world / (float(ms.item()) * 1e-3)}
index.close()
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
