#!/usr/bin/env python
"""Launches the secondary kernels once at their representative shapes so that ncu can capture them (development aid):
   python tools/ncu_probe.py flat|encode|ivf|cfg5shard"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvt_b200 import capi, synth  # noqa: E402

what = sys.argv[1]
ctx = capi.Context(0)
if what == "flat":      # makeSearch shape: 125 402 x 128, 1 536 descriptors, k = 5, IP in SSE order
    x = synth.sift_like(125_402, 128, seed=7); q = synth.sift_like(1536, 128, seed=8)
    idx = capi.FlatIndex(ctx, "ip", 128, len(x), order=4); idx.add(x, np.arange(len(x), dtype=np.uint64))
    for _ in range(3):
        idx.search(q, 5)
    idx.search(q[:1], 5)
elif what == "encode":  # a3 at cfg3's model shape: 262 144 rows, M = 16, d_sub = 8 (and M = 32, d_sub = 16)
    for D, M in ((128, 16), (512, 32)):
        n = 262_144 if D == 128 else 65_536
        x = synth.sift_like(n, D, seed=3) if D == 128 else synth.cnn_like(n, D, seed=3)
        rng = np.random.Generator(np.random.PCG64(5))
        cb = (rng.standard_normal((M, 256, D // M)) * 0.05).astype(np.float32)
        pq = capi.PQIndex.create(ctx, np.zeros((1, D), np.float32), cb)
        for _ in range(3):
            pq.encode(x)
        pq.close()
elif what == "ivf":     # f-1 at the reference's index shape: K = 8192, M = 16, nprobe = 3, 1 M rows, 4096 queries
    n = 1_000_000
    db = synth.sift_like(n, 128, seed=11); q = synth.sift_like(4096, 128, seed=12)
    perm = synth.SHIPPED_REORDER_128
    coarse, cb, _ = capi.pq_train(ctx, db[:200_000], 8192, 16, 256, perm=perm, max_iter=4, seed=1)
    pq = capi.PQIndex.create(ctx, coarse, cb, perm=perm)
    pq.add(db)
    for _ in range(3):
        pq.search(q, 100, nprobe=3)
    pq.close()
ctx.synchronize()
print("probe done:", what)
