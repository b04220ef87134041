"""Quick device-side timing of the exact fp32 scan at makeSearch's shape (development aid): 125 402 x 128 rows,
1 536 descriptors per image, k = 5, 1 - <q,x> in SSE lane order (hnsw_sifts_retrieval/makeSearch.cpp:47-62)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hashlib

import numpy as np
import torch
from cvt_b200 import capi, synth

ctx = capi.Context(0)
x = synth.sift_like(125_402, 128, seed=7)
q = synth.sift_like(1536, 128, seed=8)
idx = capi.FlatIndex(ctx, "ip", 128, len(x), order=4)
idx.add(x, np.arange(len(x), dtype=np.uint64))
qd = torch.from_numpy(q).cuda()
od = torch.empty((1536, 5), device="cuda")
ol = torch.empty((1536, 5), dtype=torch.int64, device="cuda")
for _ in range(3):
    idx.search_dev(qd.data_ptr(), 1536, 5, od.data_ptr(), ol.data_ptr())
ctx.synchronize()
ctx.event_record(0)
for _ in range(10):
    idx.search_dev(qd.data_ptr(), 1536, 5, od.data_ptr(), ol.data_ptr())
ctx.event_record(1)
ms = ctx.event_elapsed_ms(0, 1) / 10
print(f"makeSearch shape, device-resident: {ms:.3f} ms per image = {1536 * 125402 * 128 / ms / 1e9:.2f} T pair-elements/s; "
      f"labels sha1 {hashlib.sha1(ol.cpu().numpy().tobytes()).hexdigest()[:12]}")
