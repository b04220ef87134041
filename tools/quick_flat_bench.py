"""Quick timing of the flat index paths (development aid): cfg1 (fp32, 10k x 128, 100 queries, k=10) and
cfg2 (u8 L2, 1M x 128, batch 1024, k=10) plus the dense rotation GEMM."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvt_b200 import capi, synth

ctx = capi.Context(0)
def timeit(fn, it=5):
    fn(); ctx.synchronize()
    ctx.event_record(0)
    for _ in range(it): fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / it

rng = np.random.Generator(np.random.PCG64(0))
# cfg1
x = synth.sift_like(10_000, 128); q = synth.sift_like(100, 128, seed=2)
idx = capi.FlatIndex(ctx, "ip", 128, 10_000, order=4); idx.add(x, np.arange(10_000, dtype=np.uint64))
qd = torch.from_numpy(q).cuda(); od = torch.empty((100, 10), device="cuda"); ol = torch.empty((100, 10), dtype=torch.int64, device="cuda")
ms = timeit(lambda: idx.search_dev(qd.data_ptr(), 100, 10, od.data_ptr(), ol.data_ptr()))
print(f"cfg1 fp32 IP 10k x128, 100 q, k=10: {ms:.3f} ms  -> {100/ms*1e3:.0f} QPS")
idx.close()
# cfg2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
xu = rng.integers(0, 256, size=(n, 128), dtype=np.uint8); qu = rng.integers(0, 256, size=(1024, 128), dtype=np.uint8)
idx = capi.FlatIndex(ctx, "l2_u8", 128, n); idx.add(xu, np.arange(n, dtype=np.uint64))
qd = torch.from_numpy(qu).cuda(); od = torch.empty((1024, 10), dtype=torch.int32, device="cuda"); ol = torch.empty((1024, 10), dtype=torch.int64, device="cuda")
ms = timeit(lambda: idx.search_dev(qd.data_ptr(), 1024, 10, od.data_ptr(), ol.data_ptr()), it=3)
print(f"cfg2 u8 L2 {n} x128, 1024 q, k=10: {ms:.3f} ms -> {1024/ms*1e3:.0f} QPS, {2*1024*n*128/ms/1e9:.1f} Tops")
idx.close()
# rotation GEMM
D = 128
coarse = np.zeros((1, D), np.float32); cb = rng.standard_normal((16, 256, 8)).astype(np.float32)
pq = capi.PQIndex.create(ctx, coarse, cb, R=synth.dense_rotation(D))
xs = synth.sift_like(1 << 18, D)
t0 = time.time(); pq.add(xs); print(f"add 262144 rows with dense R (rotate GEMM + encode): {time.time()-t0:.3f}s")
pq.close()
