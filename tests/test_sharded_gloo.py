"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, global-id keys, the single
all-gather of per-shard top-k records and the merge order.  The per-shard search is stood in by the
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


def test_wave_split_whole_waves_then_remainder():
    # batch 4096, 8 queries per CTA, 148 SMs: 512 groups = 3 whole waves (444 groups) + 68
    assert sharded.wave_split(4096, 8, 148) == [(0, 3552), (3552, 4096)]
    assert sharded.wave_split(1024, 8, 148) == [(0, 1024)]          # less than one wave
    assert sharded.wave_split(148 * 8 * 2, 8, 148) == [(0, 2368)]    # whole waves only
    assert sharded.wave_split(0, 8, 148) == [(0, 0)]
    for nq in (1, 7, 1185, 5000, 16384):
        ch = sharded.wave_split(nq, 4, 148)
        assert ch[0][0] == 0 and ch[-1][1] == nq and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))


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


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = cases.opq_case("flat_m16")
        k = 100
        xr = orc.opq_reorder(c["db"], c["reorder"])
        codes = orc.opq_pq_encode(xr, c["coarse"], np.zeros(c["n"], np.int32), c["cb"])
        qr = orc.opq_reorder(c["q"], c["reorder"])
        lo, hi = sharded.shard_bounds(c["n"], world, rank)

        seen_rows = []

        def local_search(q, kk, rows=None):
            seen_rows.append(rows)
            D, I = orc.opq_search_flat(q.numpy(), c["coarse"][0], c["cb"], codes[lo:hi], kk, clamp=1.0)
            keys = sharded.pack_keys(D, I + lo)  # global ids = id_base + local row
            return torch.from_numpy(keys.view(np.int64))

        def merge(keys_all, rows=None):
            m = sharded.merge_keys_host(keys_all.numpy().view(np.uint64), k)
            return sharded.unpack_keys(m)

        Dref, Iref = orc.opq_search_flat(qr, c["coarse"][0], c["cb"], codes, k, clamp=1.0)
        sh = sharded.ShardedPQ(dist, rank, world, local_search, merge)
        D, I = sh.search(torch.from_numpy(qr), k)
        ok = bool(np.array_equal(I, Iref) and np.array_equal(D.view(np.uint32), Dref.view(np.uint32)))
        # the same exchange issued in query chunks (what the GPU path overlaps with the next chunk's scan)
        nq = qr.shape[0]
        cut = max(1, nq // 3)
        sh2 = sharded.ShardedPQ(dist, rank, world, local_search, merge, split=lambda n: [(0, cut), (cut, n)])
        D2, I2 = sh2.search(torch.from_numpy(qr), k)
        ok = ok and seen_rows[-2:] == [(0, cut, nq), (cut, nq, nq)]
        ok = ok and bool(np.array_equal(I2.numpy(), Iref) and np.array_equal(D2.numpy().view(np.uint32), Dref.view(np.uint32)))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_sharded_search_equals_single_index_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) is True and ret.get(1) is True
