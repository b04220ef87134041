"""Quick device-side timing of the flat ADC search stages (development aid, not the contract bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvt_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
M = int(sys.argv[3]) if len(sys.argv) > 3 else 16
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
D = 128
t0 = time.time()
db = synth.sift_like(n, D)
q = synth.sift_like(B, D, seed=synth.SEED_QUERY)
perm = synth.SHIPPED_REORDER_128
coarse, cb = synth.train_pq_model(db[:20000][:, perm], M, 256, 1, iters=6)
print(f"data+codebooks: {time.time()-t0:.1f}s", flush=True)
ctx = capi.Context(0)
idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
t0 = time.time(); idx.add(db); print(f"add {n} rows: {time.time()-t0:.2f}s", flush=True)
qd = torch.from_numpy(q).cuda()
od = torch.empty((B, k), dtype=torch.float32, device="cuda")
oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
for it in range(5):
    idx.search_dev(qd.data_ptr(), B, k, 1, od.data_ptr(), oi.data_ptr())
    ctx.synchronize()
    t = idx.last_timing()
    tot = sum(t.values())
    print(f"iter {it}: {t}  total {tot:.3f} ms  QPS {B/tot*1e3:.0f}  alg GB/s {B*n*M/ (t['scan_ms']*1e-3)/1e9:.0f}", flush=True)
kth = od[:, k-1].cpu().numpy()
print("k-th score: median %.4f  frac<1.0: %.4f" % (np.median(kth), (kth < 1.0).mean()))
print("launches", ctx.launch_count())
import hashlib
print("ids sha1", hashlib.sha1(oi.cpu().numpy().tobytes()).hexdigest()[:16], "scores sha1", hashlib.sha1(od.cpu().numpy().tobytes()).hexdigest()[:16])
if os.environ.get("QUICK_STATS"):
    os.environ["B200NN_SCAN_STATS"] = "1"
    idx.search_dev(qd.data_ptr(), B, k, 1, od.data_ptr(), oi.data_ptr())
    ctx.synchronize()
