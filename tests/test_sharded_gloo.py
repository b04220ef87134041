"""world_size-2 and -4 gloo tests (CPU) of the multi-GPU host logic: the (query chunk x row shard) grid
of ranks, shard bounds, global-id keys, the single all-gather of per-rank top-k records and the merge order.  The per-shard search is stood in by the
oracle (this is the exchange step's test; the CUDA kernels are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from cvt_b200 import sharded
from oracle import oracle as orc


def test_shard_bounds_cover_and_are_contiguous():
    for n in (0, 1, 7, 53, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            edges = [sharded.shard_bounds(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            per = -(-n // w) if n else 0
            assert all(hi - lo <= per for lo, hi in edges)


def test_grid_coords_and_query_chunks():
    # rank = chunk * R + shard: the row shards of one query chunk are adjacent ranks
    assert [sharded.grid_coords(r, 2) for r in range(8)] == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2), (0, 3), (1, 3)]
    assert [sharded.grid_coords(r, 8) for r in range(8)] == [(r, 0) for r in range(8)]
    for nq in (0, 1, 7, 24, 4096, 5001):
        for Q in (1, 2, 3, 8):
            ch = [sharded.query_chunk(nq, Q, c) for c in range(Q)]
            assert ch[0][0] == 0 and ch[-1][1] == nq and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert len({c[2] for c in ch}) == 1 and all(hi - lo <= cq for lo, hi, cq in ch)


def test_plan_layout_replicates_small_databases_and_row_shards_large_ones():
    # cfg3 (1M rows = 16 MB of codes): a 1/8 shard no longer amortises the per-(query, CTA) warm-up
    for world in (2, 4, 8):
        R, Q = sharded.plan_layout(world, 1_000_000, 4096, 16, 100)
        assert R * Q == world and Q > 1
        cost = {r: sharded.layout_cost(world, r, 1_000_000, 4096, 16, 100) for r in (1, world)}
        assert cost[1] < cost[world]
    # cfg5 (100M rows, batch 16384): warm-ups are noise, the memory-minimal row-sharded layout is kept
    for world in (2, 4, 8):
        assert sharded.plan_layout(world, 100_000_000, 16384, 16, 100) == (world, 1)
    assert sharded.plan_layout(1, 1_000_000, 4096, 16, 100) == (1, 1)
    # a memory cap forces row shards
    assert sharded.plan_layout(8, 1_000_000, 4096, 16, 100, max_rows_per_gpu=130_000) == (8, 1)


def test_merge_grid_host_equals_global_sort():
    rng = np.random.Generator(np.random.PCG64(7))
    Q, R, cq, k, nq = 3, 2, 5, 4, 13
    keys = np.sort(rng.integers(0, 1 << 62, size=(Q, R, cq, k), dtype=np.uint64), axis=3)
    out = sharded.merge_grid_host(keys, nq, k)
    assert out.shape == (nq, k)
    for q in range(nq):
        c, ql = divmod(q, cq)
        assert np.array_equal(out[q], np.sort(keys[c, :, ql, :].reshape(-1))[:k])


def test_key_packing_orders_like_pairs():
    rng = np.random.Generator(np.random.PCG64(1))
    d = rng.standard_normal(5000).astype(np.float32)
    d[:50] = d[50:100]  # exact ties
    d[100] = 0.0
    d[101] = -0.0
    ids = rng.permutation(5000).astype(np.int64)
    keys = sharded.pack_keys(d, ids)
    order_keys = np.argsort(keys, kind="stable")
    order_pairs = np.lexsort((ids, d))
    # -0.0 < +0.0 in key order, equal in float order: exclude that pair from the strict comparison
    mask = ~np.isin(order_pairs, [100, 101])
    assert np.array_equal(order_keys[np.isin(order_keys, order_pairs[mask])], order_pairs[mask])
    dd, ii = sharded.unpack_keys(keys)
    assert np.array_equal(dd.view(np.uint32), d.view(np.uint32)) and np.array_equal(ii, ids)


def _worker(rank, world, port, layouts, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = cases.opq_case("flat_m16")
        k = 100
        xr = orc.opq_reorder(c["db"], c["reorder"])
        codes = orc.opq_pq_encode(xr, c["coarse"], np.zeros(c["n"], np.int32), c["cb"])
        qr = orc.opq_reorder(c["q"], c["reorder"])[:21]  # 21 queries: ragged chunks for Q = 2 and 4
        Dref, Iref = orc.opq_search_flat(qr, c["coarse"][0], c["cb"], codes, k, clamp=1.0)
        ok = True
        for R in layouts:
            r, ch = sharded.grid_coords(rank, R)
            lo, hi = sharded.shard_bounds(c["n"], R, r)
            seen = []

            def local_search(q, kk):
                seen.append(q.shape[0])
                D, I = orc.opq_search_flat(q.numpy(), c["coarse"][0], c["cb"], codes[lo:hi], kk, clamp=1.0)
                keys = sharded.pack_keys(D, I + lo)  # global ids = id_base + local row
                return torch.from_numpy(keys.view(np.int64))

            def merge(keys_grid, nq):
                return sharded.unpack_keys(sharded.merge_grid_host(keys_grid.numpy().view(np.uint64), nq, k))

            sh = sharded.ShardedPQ(dist, rank, world, local_search, merge, row_shards=R)
            D, I = sh.search(torch.from_numpy(qr), k)
            qlo, qhi, _ = sharded.query_chunk(qr.shape[0], world // R, ch)
            ok = ok and seen == [qhi - qlo]  # the rank searched exactly its chunk of the batch
            ok = ok and bool(np.array_equal(I, Iref) and np.array_equal(D.view(np.uint32), Dref.view(np.uint32)))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def _run(world, layouts):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, layouts, ret), nprocs=world, join=True)
    assert all(ret.get(r) is True for r in range(world))


def test_sharded_search_equals_single_index_world2():
    _run(2, [2, 1])  # plain row shards (queries replicated); replicated rows, queries split


def test_sharded_search_equals_single_index_world4_grid():
    _run(4, [2, 4, 1])  # 2 row shards x 2 query chunks, and both pure layouts


def test_sharded_search_equals_single_index_world8_grid():
    _run(8, [2, 8])  # the grid planned for cfg3 on 8 GPUs (2 row shards x 4 query chunks) and plain row sharding
