"""Development aid: sweep the scan kernel's runtime knobs (env) and print its candidate-path counters."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvt_b200 import capi, synth

B, M, k, D = 4096, 16, 100, 128
sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [125_000, 1_000_000]
knobs = sys.argv[2].split(";") if len(sys.argv) > 2 else ["", "B200NN_TAU_EVERY=8", "B200NN_TAU_EVERY=4"]
nmax = max(sizes)
db = synth.sift_like(nmax, D)
q = synth.sift_like(B, D, seed=synth.SEED_QUERY)
perm = synth.SHIPPED_REORDER_128
coarse, cb = synth.train_pq_model(db[:20000][:, perm], M, 256, 1, iters=6)
ctx = capi.Context(0)
qd = torch.from_numpy(q).cuda()
od = torch.empty((B, k), dtype=torch.float32, device="cuda")
oi = torch.empty((B, k), dtype=torch.int64, device="cuda")
ref_ids = {}
for n in sizes:
    idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
    idx.add(db[:n])
    for kn in knobs:
        sets = dict(kv.split("=") for kv in kn.split(",") if kv)
        for a, b in sets.items():
            os.environ[a] = b
        ts = []
        for it in range(6):
            idx.search_dev(qd.data_ptr(), B, k, 1, od.data_ptr(), oi.data_ptr())
            ctx.synchronize()
            ts.append(idx.last_timing()["scan_ms"])
        ids = oi.cpu().numpy().copy()
        same = True if n not in ref_ids else bool((ids == ref_ids[n]).all())
        ref_ids.setdefault(n, ids)
        print(f"n={n} knobs[{kn}] scan_ms min {min(ts[1:]):.3f} med {sorted(ts[1:])[2]:.3f}  ids_same_as_first={same}", flush=True)
        os.environ["B200NN_SCAN_STATS"] = "1"
        idx.search_dev(qd.data_ptr(), B, k, 1, od.data_ptr(), oi.data_ptr())
        ctx.synchronize()
        del os.environ["B200NN_SCAN_STATS"]
        for a in sets:
            del os.environ[a]
    idx.close()
