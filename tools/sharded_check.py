#!/usr/bin/env python
"""Parity of the one-process-per-GPU sharded search through the C ABI (b200nn_comm_* + b200nn_pq_search_sharded_dev):
run under torchrun with N ranks; every (row shards x query chunks) grid must return exactly what a single index holding
all rows returns.  torch.distributed is used for the rendezvous only (broadcast of the NCCL id, barriers); the exchange
itself is the library's own ncclAllGather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/sharded_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cases
    from cvt_b200 import capi, sharded

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # rendezvous only
    ctx = capi.Context(local)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm(ctx, rank, world, uid[0])
    assert comm.nccl_version() > 0
    for case, nprobe, k in (("flat_m16", 1, 100), ("ivf_m8", 3, 20)):
        c = cases.opq_case(case)
        n = c["n"]
        q = np.concatenate([c["q"], c["q"] * np.float32(3.0)])[:41]  # ragged vs. the chunk sizes
        single = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
        single.add(c["db"])
        D0, I0 = single.search(q, k, nprobe=nprobe)
        qd = torch.from_numpy(q).cuda()
        for R in [r for r in range(1, world + 1) if world % r == 0]:
            shard_i, chunk_i = sharded.grid_coords(rank, R)
            lo, hi = sharded.shard_bounds(n, R, shard_i)
            idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
            idx.add(c["db"][lo:hi])
            od = torch.empty((len(q), k), dtype=torch.float32, device="cuda")
            oi = torch.empty((len(q), k), dtype=torch.int64, device="cuda")
            for _ in range(2):
                idx.search_sharded_dev(comm, R, qd.data_ptr(), len(q), k, nprobe, lo, od.data_ptr(), oi.data_ptr())
                ctx.synchronize()
                assert np.array_equal(oi.cpu().numpy(), I0.astype(np.int64)), (case, R, rank)
                assert np.array_equal(od.cpu().numpy().view(np.uint32), D0.view(np.uint32)), (case, R, rank)
            idx.close()
            dist.barrier()
        single.close()
    comm.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("sharded_check ok: world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
