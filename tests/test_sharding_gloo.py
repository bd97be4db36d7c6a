"""CPU, world_size 2 over gloo: the host-side logic of the sharded paths (frame shards; row-sharded kNN partial layout,
all-gather and tie rules).  The per-shard top-2 and the merge run as CUDA kernels in production; here the oracle stands
in for both so that only the plumbing in send_slam_b200/sharded.py is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from send_slam_b200 import sharded, synth

NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def test_frame_and_row_shards_cover_exactly():
    for n in (0, 1, 7, 64, 65, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharded.frame_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [sharded.frame_shard(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert sharded.row_shard(10_000_000, 8, 3) == (3_750_000, 5_000_000)
    with pytest.raises(ValueError):
        sharded.frame_shard(10, 2, 2)


def pack(idx, dist_):
    out = (dist_.astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)
    out[idx < 0] = NONE
    return out


def host_merge(gathered: torch.Tensor) -> torch.Tensor:
    g = gathered.numpy().view(np.uint64)                      # [world, nq, 2]
    allk = np.sort(np.concatenate(list(g), axis=1), axis=1)   # per query: ascending (dist, row) keys
    return torch.from_numpy(allk[:, :2].copy().view(np.int64))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_lib as ol
        db = synth.descriptor_db(3001, seed=77)
        db[2500] = db[10]                                     # cross-shard exact tie: lowest global row must win
        q, _ = synth.queries_from_db(db, 50, seed=78)
        q[0] = db[10]
        a, b = sharded.row_shard(len(db), world, rank)
        idx, dst = ol.knn2(q, db[a:b])
        idx = np.where(idx >= 0, idx + a, idx)
        local = torch.from_numpy(pack(idx, dst).view(np.int64))
        merged = sharded.knn2_all_gather_merge(local, host_merge)
        ret[rank] = merged.numpy().view(np.uint64).copy()
    finally:
        dist.destroy_process_group()


def test_row_sharded_knn_merge_world2(oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    db = synth.descriptor_db(3001, seed=77)
    db[2500] = db[10]
    q, _ = synth.queries_from_db(db, 50, seed=78)
    q[0] = db[10]
    idx, dst = oracle.knn2(q, db)
    want = pack(idx, dst)
    assert np.array_equal(ret[0], want) and np.array_equal(ret[1], want)
    assert idx[0].tolist() == [10, 2500] and dst[0].tolist() == [0, 0]
