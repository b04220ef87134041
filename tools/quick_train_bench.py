"""Quick timing of device training (development aid, f-4): the bench-shaped flat PQ model and a coarse quantizer."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvt_b200 import capi, synth

ctx = capi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
x = synth.sift_like(n, 128, seed=5)
perm = synth.SHIPPED_REORDER_128
capi.kmeans(ctx, x[:2000], 8, 2, 1)  # warm-up (module load)
for iters in (10,):
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    coarse, cb, mse = capi.pq_train(ctx, x, 0, 16, 256, perm=perm, max_iter=iters, seed=3)
    dt = time.perf_counter() - t0
    print(f"pq_train flat M=16 x 256, {n} x 128 rows, {iters} updates per sub-space: {dt:.3f} s, {ctx.launch_count() - l0} launches, "
          f"quantisation error {mse[1:].sum():.5f}")
t0 = time.perf_counter()
_, cb_np = synth.train_pq_model(x[:, perm], 16, 256, 1, iters=10, train_rows=n)
print(f"numpy Lloyd (synth.train_pq_model), same shape: {time.perf_counter() - t0:.3f} s")
for K, iters in ((1024, 5), (8192, 1)):
    t0 = time.perf_counter()
    c, a, d, it, m = capi.kmeans(ctx, x, K, iters, 4)
    dt = time.perf_counter() - t0
    print(f"kmeans K={K} over {n} x 128, {it} updates: {dt:.3f} s ({3.0 * n * K * 128 * (it + 1) / dt / 1e12:.2f} TFLOP/s fp32 in the assignment), mse {m:.5f}")
# a2 at the shipped K = 8192: IVFOPQ::Add's coarse assignment + PQ encode of 262144 rows (tiled distance kernel)
K = 8192
rng = np.random.Generator(np.random.PCG64(0))
coarse = x[rng.choice(n, K, replace=False)][:, perm].copy()
cb = rng.standard_normal((16, 256, 8)).astype(np.float32) * 0.05
idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm)
xs = synth.sift_like(1 << 18, 128, seed=6)
idx.add(xs[:4096])
ctx.synchronize()
t0 = time.perf_counter()
idx.add(xs)
ctx.synchronize()
dt = time.perf_counter() - t0
print(f"pq add (rotate + coarse assign K={K} + encode) of {len(xs)} rows from host: {dt:.3f} s ({len(xs) * K * 128 / dt / 1e12:.2f} T pair-elements/s incl. H2D)")
q = synth.sift_like(4096, 128, seed=7)
idx.search(q[:64], 10, nprobe=3)
t0 = time.perf_counter()
idx.search(q, 10, nprobe=3)
print(f"IVF search 4096 queries, nprobe=3, K={K}, {idx.n_rows} rows: {time.perf_counter() - t0:.4f} s")
