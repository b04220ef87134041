"""Sharded (O)PQ search across the GPUs of one box: one process per GPU (torch.distributed / NCCL
for the plumbing).  The ranks form a (query chunk x row shard) grid: a rank owns a contiguous block
of database rows as its own index and serves one chunk of the query batch; the only exchange is ONE
all-gather of the per-rank top-k records (64-bit sortable keys) followed by a
per-query merge kernel over the row shards.  row_shards = world is the plain row-sharded layout
(queries replicated); plan_layout() chooses the grid.

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


def merge_grid_host(keys_grid: np.ndarray, nq: int, k: int) -> np.ndarray:
    """[Q, R, chunk_q, k] gathered records of a (query chunk x row shard) grid -> [nq, k]: the CPU twin of
    b200nn_topk_merge_grid_dev (rows past nq are padding)."""
    Q, R, cq, _ = keys_grid.shape
    return np.concatenate([merge_keys_host(keys_grid[c], k) for c in range(Q)], axis=0)[:nq]


# ---- layout: how the ranks of one box split (rows x queries) -----------------------------------
# Constants of the cost model, measured on B200 this round (profiles/r01_summary.md):
_SCAN_BYTES_PER_S = 6.0e12    # algorithmic code bytes per second of the streaming phase of the fused scan, one GPU
_WARMUP_S = 30e-6             # parked warm-up + selection (phases A + B) per (CTA, segment)
_CANDIDATE_S = 0.015e-6       # streamed candidate, per (query, segment): ~1.5 k ln(rows/2048) of them
_LUT_S_PER_QUERY = 0.02e-6    # rotate + LUT build per query
_EXCHANGE_S = 40e-6           # all-gather launch + rank skew floor
_LINK_BYTES_PER_S = 300e9     # effective all-gather bandwidth at these message sizes
_MERGE_BYTES_PER_S = 1.0e12   # merge kernel: bytes of gathered records it reads


def layout_cost(world: int, row_shards: int, n_rows: int, batch: int, M: int, k: int, sm_count: int = 148) -> float:
    """Modelled seconds per step of a (row_shards x world/row_shards) grid of ranks."""
    R, Q = row_shards, world // row_shards
    rows = -(-n_rows // R)
    bq = -(-batch // Q)
    qw = max(1, 128 // M)
    groups = -(-bq // qw)

    def segment(r):  # top-k warm-up of one (CTA, segment) over r rows
        return _WARMUP_S + qw * 1.5 * k * float(np.log(max(r, 4096.0) / 2048.0)) * _CANDIDATE_S

    waves, rem = divmod(groups, sm_count)
    # every SM runs `waves` whole-shard segments, then a piece of the last wave (1-2 segments, 1.5 on average)
    t_warm = waves * segment(rows) + (1.5 * segment(rem * rows / sm_count) if rem else 0.0)
    t_scan = bq * rows * M / _SCAN_BYTES_PER_S
    t_lut = bq * _LUT_S_PER_QUERY
    gathered = world * bq * k * 8
    t_x = 0.0 if world == 1 else _EXCHANGE_S + gathered / _LINK_BYTES_PER_S + (gathered / _MERGE_BYTES_PER_S if R > 1 else 0.0)
    return t_scan + t_warm + t_lut + t_x


def plan_layout(world: int, n_rows: int, batch: int, M: int, k: int, sm_count: int = 148,
                max_rows_per_gpu: int = 1 << 30, tolerance: float = 0.02) -> tuple[int, int]:
    """(row_shards R, query_chunks Q) with R * Q = world.

    R = world is the memory-minimal layout (every rank scans its block of rows for
    the whole batch).  It is kept whenever the model puts it within `tolerance` of the best layout; for a
    database so small that a 1/world shard no longer amortises the per-(query, CTA) top-k warm-up (cfg3:
    1M rows = 16 MB of codes), ranks instead replicate row blocks and split the query batch, which cuts
    the warm-ups, the LUT builds and the gathered payload by Q.  The collective is the same single
    all-gather either way."""
    cands = [r for r in range(1, world + 1) if world % r == 0 and -(-n_rows // r) <= max_rows_per_gpu]
    if not cands:
        return world, 1
    cost = {r: layout_cost(world, r, n_rows, batch, M, k, sm_count) for r in cands}
    best = min(cost.values())
    r = max(r for r in cands if cost[r] <= best * (1.0 + tolerance))
    return r, world // r


def grid_coords(rank: int, row_shards: int) -> tuple[int, int]:
    """rank -> (row shard r, query chunk c); rank = c * row_shards + r (row shards of one chunk are adjacent
    ranks, so the gathered buffer reads [Q][R][chunk_q][k])."""
    return rank % row_shards, rank // row_shards


def query_chunk(nq: int, n_chunks: int, c: int) -> tuple[int, int, int]:
    """Queries [lo, hi) of chunk c and the (padded) chunk size."""
    cq = -(-nq // n_chunks) if nq else 0
    lo = min(nq, c * cq)
    return lo, min(nq, lo + cq), cq


class ShardedPQ:
    """Distributed search over a (query chunk x row shard) grid of ranks, one PQIndex shard per rank.

    `local_search(q_chunk, k) -> keys [n, k]` (int64 view of the uint64 records, ascending, GLOBAL ids)
    and `merge(keys_grid [Q, R, chunk_q, k], nq) -> (dist, ids)` are injectable so that the exchange
    logic runs under gloo on CPU in the tests; on GPUs they are the C-ABI calls (pq_search_dev with
    out_key, topk_merge_grid_dev).  The collective is ONE all-gather of [chunk_q, k] records per rank."""

    def __init__(self, dist_module, rank: int, world: int, local_search, merge, row_shards: int | None = None):
        self.dist, self.rank, self.world = dist_module, rank, world
        self.R = world if row_shards is None else int(row_shards)
        if self.R < 1 or world % self.R:
            raise ValueError(f"row_shards={row_shards} must divide world={world}")
        self.Q = world // self.R
        self.r, self.c = grid_coords(rank, self.R)
        self.local_search, self.merge = local_search, merge
        self._gathered = {}

    def search(self, q, k: int):
        import torch
        nq = q.shape[0]
        lo, hi, cq = query_chunk(nq, self.Q, self.c)
        keys = self.local_search(q[lo:hi], k)
        if self.world == 1:
            return self.merge(keys.view(1, 1, nq, k), nq)
        if keys.shape[0] != cq:  # ragged last chunk: pad with empty records (all-gather needs equal parts)
            pad = torch.full((cq, k), -1, dtype=keys.dtype, device=keys.device)
            pad[: hi - lo] = keys
            keys = pad
        tag = (cq, k, keys.device)
        gathered = self._gathered.get(tag)
        if gathered is None:
            gathered = self._gathered[tag] = torch.empty((self.world * cq, k), dtype=keys.dtype, device=keys.device)
        self.dist.all_gather_into_tensor(gathered, keys)  # the single collective of the path
        return self.merge(gathered.view(self.Q, self.R, cq, k), nq)


def make_gpu_sharded(ctx, index, dist_module, rank: int, world: int, id_base: int, nprobe: int = 1,
                     row_shards: int | None = None):
    """Wire a ShardedPQ to the CUDA library for torch CUDA tensors.  `index` holds the rows of row shard
    rank % row_shards (first global row = id_base).  The caller's current torch stream must be the stream
    the context launches on (ctx.set_stream).

    (An exchange issued in wave-aligned query chunks on a side stream, overlapping the next chunk's scan,
    was built and measured earlier this round: 6.39 vs 6.18 ms per step on 2 B200s -- slower, because two
    scan launches expose the first launch's wave tail -- and removed.)"""
    import torch

    local_bufs, merged_bufs, state = {}, {}, {}  # reused across steps: no allocator traffic inside the timed region

    def local_search(q, k):
        n = q.shape[0]
        if (n, k) not in local_bufs:
            local_bufs[(n, k)] = (torch.full((max(n, 1), k), -1, dtype=torch.int64, device=q.device),
                                  torch.empty((max(n, 1), k), dtype=torch.float32, device=q.device),
                                  torch.empty((max(n, 1), k), dtype=torch.int64, device=q.device))
        keys, dist, ids = local_bufs[(n, k)]
        if n:
            index.search_dev(q.data_ptr(), n, k, nprobe, dist.data_ptr(), ids.data_ptr(), keys.data_ptr(), id_base)
        state["local"] = (dist[:n], ids[:n])
        return keys[:n]

    def merge(keys_grid, nq):
        Q, R, cq, k = keys_grid.shape
        if Q * R == 1:
            return state["local"]
        if (nq, k) not in merged_bufs:
            merged_bufs[(nq, k)] = (torch.empty((nq, k), dtype=torch.float32, device=keys_grid.device),
                                    torch.empty((nq, k), dtype=torch.int64, device=keys_grid.device))
        dist, ids = merged_bufs[(nq, k)]
        ctx.topk_merge_grid_dev(keys_grid.data_ptr(), Q, R, cq, nq, k, dist.data_ptr(), ids.data_ptr())
        return dist, ids

    return ShardedPQ(dist_module, rank, world, local_search, merge, row_shards=row_shards)
