"""Multi-GPU plumbing for the two places the hot path shards; one process per GPU, torch.distributed.

* Extraction: frames / camera streams are independent units -> contiguous frame shards per rank, NO collective.
* Brute-force kNN (k=2): the descriptor database is row-sharded; every rank computes the top-2 of all queries in its
  shard (CUDA, orbx_knn2_query_device), the 16-byte-per-query partials are exchanged with ONE all-gather over
  NCCL/NVLink (2000 queries x 2 x 8 B = 32 KB per rank) and merged on every rank by the CUDA merge kernel
  (orbx_knn2_merge_device).  Ties resolve to the lowest global row, as cv::BFMatcher does.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def frame_shard(nframes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of a batch for `rank`; the first nframes % world ranks take one extra frame."""
    if world < 1 or not (0 <= rank < world) or nframes < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(nframes, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def row_shard(nrows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) rows of the descriptor database held by `rank` (row_offset = start)."""
    return frame_shard(nrows, world, rank)


def knn2_all_gather_merge(local_packed, merge: Callable, group=None):
    """local_packed: tensor [nq, 2] int64 (dist << 32 | global row; -1 = missing) of this rank's shard, on the device the
    process group communicates on.  Returns merge(gathered [world, nq, 2]) -- `merge` is Knn2Index.merge_device-backed
    on a GPU (see knn2_sharded); the gloo tests inject a host merge to check layout and tie rules on CPU."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return merge(local_packed.unsqueeze(0))
    flat = local_packed.contiguous().view(-1)
    gathered = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(gathered, flat, group=group)      # rank-major: [world][nq][2]
    return merge(gathered.view((world,) + tuple(local_packed.shape)))


def knn2_sharded(index, d_queries, group=None, stream: Optional[int] = None):
    """Top-2 of every query over the WHOLE row-sharded database.  index: orbx.Knn2Index of this rank's shard (created
    with row_offset = row_shard(...)[0]); d_queries: uint8 CUDA tensor [nq, 32], identical on every rank.
    Returns an int64 CUDA tensor [nq, 2] (dist << 32 | global row), identical on every rank."""
    import torch
    nq = d_queries.shape[0]
    dev = d_queries.device
    if stream is None:
        # torch's default stream has handle 0, which the C ABI reads as "the handle's own stream": name the legacy default
        # stream explicitly (cudaStreamLegacy = 1) so that the query, the NCCL all-gather and the merge stay ordered
        stream = torch.cuda.current_stream(dev).cuda_stream or 1
    index.set_stream(stream)
    local = torch.empty((nq, 2), dtype=torch.int64, device=dev)
    index.query_device(d_queries.data_ptr(), nq, local.data_ptr())

    def merge(gathered):
        out = torch.empty((nq, 2), dtype=torch.int64, device=dev)
        index.merge_device(gathered.data_ptr(), gathered.shape[0], nq, out.data_ptr())
        return out

    return knn2_all_gather_merge(local, merge, group)
