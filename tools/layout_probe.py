"""Development aid (ONE GPU): time the local work of one rank -- rotate + LUT + scan + slice merge over its
(query chunk, row shard) -- for every grid layout of 2, 4 and 8 ranks on the cfg3 workload, next to the
figure cvt_b200.sharded.layout_cost models for the whole step.  The exchange is not included (one GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvt_b200 import capi, sharded, synth

n, B, M, k, D = 1_000_000, 4096, 16, 100, 128
db = synth.sift_like(n, D)
q = synth.sift_like(B, D, seed=synth.SEED_QUERY)
perm = synth.SHIPPED_REORDER_128
coarse, cb = synth.train_pq_model(db[:20000][:, perm], M, 256, 1, iters=6)
ctx = capi.Context(0)
qd = torch.from_numpy(q).cuda()
od = torch.empty((B, k), dtype=torch.float32, device="cuda")
oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
indexes = {}
print("world R  Q  rows/GPU queries/GPU | rotate   lut    scan   merge  local_ms | model_step_ms  GB/s(alg)")
for world in (1, 2, 4, 8):
    for R in [r for r in (1, 2, 4, 8) if world % r == 0]:
        Q = world // R
        rows = -(-n // R)
        bq = -(-B // Q)
        if rows not in indexes:
            idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
            idx.add(db[:rows])
            indexes[rows] = idx
        idx = indexes[rows]
        ts = []
        for it in range(6):
            flush.zero_()
            idx.search_dev(qd.data_ptr(), bq, k, 1, od.data_ptr(), oi.data_ptr())
            ctx.synchronize()
            ts.append(idx.last_timing())
        t = {key: float(np.median([x[key] for x in ts[1:]])) for key in ts[0]}
        tot = sum(t.values())
        print(f"{world:5d} {R:2d} {Q:2d} {rows:9d} {bq:11d} | {t['rotate_ms']:6.3f} {t['lut_ms']:6.3f} {t['scan_ms']:7.3f} {t['merge_ms']:6.3f} {tot:8.3f} | "
              f"{sharded.layout_cost(world, R, n, B, M, k) * 1e3:8.3f}      {bq * rows * M / (t['scan_ms'] * 1e-3) / 1e9:6.0f}", flush=True)
