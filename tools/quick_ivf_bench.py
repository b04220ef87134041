"""Quick timing of the reference's actual index shape (f-1): IVF with K = 8192 coarse centroids, M = 16 x 256 PQ on the
residuals, nprobe = 3 -- trained, built and queried on the device (development aid)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvt_b200 import capi, synth

ctx = capi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K, M, nq, k = 8192, 16, 4096, 100
perm = synth.SHIPPED_REORDER_128
t0 = time.perf_counter()
db = synth.sift_like(n, 128, seed=synth.SEED_DB)
q = synth.sift_like(nq, 128, seed=synth.SEED_QUERY)
print(f"host: generated {n} + {nq} rows in {time.perf_counter() - t0:.1f} s")
capi.kmeans(ctx, db[:2000], 8, 2, 1)  # warm-up
t0 = time.perf_counter()
coarse, cb, mse = capi.pq_train(ctx, db[:200_000], K, M, 256, perm=perm, max_iter=10, seed=synth.SEED_KMEANS)
print(f"pq_train K={K}, M={M} x 256 on 200000 rows, 10 updates per stage: {time.perf_counter() - t0:.3f} s; "
      f"coarse mse {mse[0]:.5f}, residual quantisation error {mse[1:].sum():.5f}")
idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
t0 = time.perf_counter()
idx.add(db)
ctx.synchronize()
dt = time.perf_counter() - t0
print(f"IVFOPQ::Add of {n} rows from host memory (rotate + coarse assign + residual PQ encode): {dt:.3f} s = {n / dt / 1e6:.2f} M rows/s")
idx.search(q[:256], k, nprobe=3)
for nprobe in (1, 3, 8):
    idx.search(q, k, nprobe=nprobe)  # workspaces of this shape
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        d, i = idx.search(q, k, nprobe=nprobe)
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[2]
    found = float((i < n).mean())
    print(f"IVF search {nq} queries, top-{k}, nprobe={nprobe}: {dt * 1e3:.2f} ms = {nq / dt / 1e3:.0f} k QPS (host buffers); filled result slots {found:.3f}")
