"""Row-sharded (O)PQ search across the GPUs of one box: one process per GPU (torch.distributed /
NCCL for the plumbing), each rank owns a contiguous block of database rows as its own index,
queries are replicated, and the only exchange is ONE all-gather of the per-shard top-k records
(64-bit sortable keys) followed by a per-query merge kernel.

The merge order is the reference's (score, id) lexicographic order on GLOBAL ids, so the sharded
result is identical to a single-index search.  (The same shape as the unbuilt boost.MPI
shard+merge layer in the reference's vendored FLANN, retrieval/vlindex/lib/FLANN/mpi/index.h:177-214.)
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) of shard `rank`: contiguous blocks of ceil(n/world) rows."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


# ---- sortable records (host-side mirror of csrc/common.cuh, used by tests and for decoding) ----
def f32_orderable(d: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(d, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def f32_from_orderable(o: np.ndarray) -> np.ndarray:
    o = np.ascontiguousarray(o, dtype=np.uint32)
    b = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return b.view(np.float32)


def pack_keys(dist: np.ndarray, ids: np.ndarray) -> np.ndarray:
    return (f32_orderable(dist).astype(np.uint64) << np.uint64(32)) | (np.asarray(ids).astype(np.uint64) & np.uint64(0xFFFFFFFF))


def unpack_keys(keys: np.ndarray):
    keys = np.asarray(keys, dtype=np.uint64)
    return f32_from_orderable((keys >> np.uint64(32)).astype(np.uint32)), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)


def merge_keys_host(keys_all: np.ndarray, k: int) -> np.ndarray:
    """[L, nq, k] sorted key lists -> [nq, k] smallest keys (plain numpy; the CPU twin of
    topk_merge_kernel, used by the gloo tests of the exchange step)."""
    L, nq, _ = keys_all.shape
    flat = np.transpose(keys_all, (1, 0, 2)).reshape(nq, L * keys_all.shape[2])
    return np.sort(flat, axis=1)[:, :k]


class ShardedPQ:
    """Distributed search over per-rank PQIndex shards.

    `local_search(q, k) -> keys [nq, k] (uint64 tensor, ascending)` and `merge(keys_all [L,nq,k]) ->
    (dist, ids)` are injectable so that the exchange logic runs under gloo on CPU in the tests;
    on GPUs they are the C-ABI calls (pq_search_dev with out_key, topk_merge_dev)."""

    def __init__(self, dist_module, rank: int, world: int, local_search, merge):
        self.dist, self.rank, self.world = dist_module, rank, world
        self.local_search, self.merge = local_search, merge
        self._gathered = {}

    def search(self, q, k: int):
        import torch
        keys_local = self.local_search(q, k)  # [nq, k] int64 view of uint64 keys
        if self.world == 1:
            return self.merge(keys_local.unsqueeze(0))
        nq, kk = keys_local.shape
        gathered = self._gathered.get((nq, kk))
        if gathered is None:
            gathered = self._gathered[(nq, kk)] = torch.empty((self.world * nq, kk), dtype=keys_local.dtype, device=keys_local.device)
        self.dist.all_gather_into_tensor(gathered, keys_local)  # the single collective of the path
        return self.merge(gathered.view(self.world, nq, kk))


def make_gpu_sharded(ctx, index, dist_module, rank: int, world: int, id_base: int, nprobe: int = 1):
    """Wire a ShardedPQ to the CUDA library for torch CUDA tensors."""
    import torch

    bufs = {}

    def _buffers(nq, k, device):
        key = (nq, k)
        if key not in bufs:  # reused across steps: no allocator traffic inside the timed region
            bufs[key] = dict(keys=torch.empty((nq, k), dtype=torch.int64, device=device),
                             dist=torch.empty((nq, k), dtype=torch.float32, device=device),
                             ids=torch.empty((nq, k), dtype=torch.int64, device=device),
                             mdist=torch.empty((nq, k), dtype=torch.float32, device=device),
                             mids=torch.empty((nq, k), dtype=torch.int64, device=device))
        return bufs[key]

    def local_search(q, k):
        nq = q.shape[0]
        b = _buffers(nq, k, q.device)
        keys, dist, ids = b["keys"], b["dist"], b["ids"]
        index.search_dev(q.data_ptr(), nq, k, nprobe, dist.data_ptr(), ids.data_ptr(), keys.data_ptr(), id_base)
        local_search.last = (dist, ids)
        return keys

    def merge(keys_all):
        L, nq, k = keys_all.shape
        if L == 1 and hasattr(local_search, "last"):
            return local_search.last
        b = _buffers(nq, k, keys_all.device)
        dist, ids = b["mdist"], b["mids"]
        ctx.topk_merge_dev(keys_all.data_ptr(), L, nq, k, dist.data_ptr(), ids.data_ptr())
        return dist, ids

    return ShardedPQ(dist_module, rank, world, local_search, merge)
