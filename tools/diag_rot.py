import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from cvt_b200 import capi, synth
ctx = capi.Context(0)
D, M = 128, 16
rng = np.random.Generator(np.random.PCG64(0))
coarse = np.zeros((1, D), np.float32); cb = rng.standard_normal((M, 256, D // M)).astype(np.float32)
def run(R, x, tag):
    idx = capi.PQIndex.create(ctx, coarse, cb, R=R)
    y = idx.rotate(x); idx.close()
    y64 = x.astype(np.float64) @ R.astype(np.float64).T
    bad = y != y64.astype(np.float32)
    rel = np.abs(y - y64) / np.maximum(np.abs(y64), 1e-300)
    print(tag, "mismatch frac", bad.mean(), "max rel", rel.max(), "rows bad", bad.any(1).sum(), "cols bad", bad.any(0).sum())
    if bad.any():
        r, c = np.argwhere(bad)[0]
        print("   first bad", r, c, y[r, c], y64[r, c], hex(y[r,c:c+1].view(np.uint32)[0]), hex(np.float32(y64[r,c])[None].view(np.uint32)[0]))
        print("   bad by col (first 16):", bad.sum(0)[:16], " bad rows mod 8 hist:", np.bincount(np.argwhere(bad)[:,0] % 8, minlength=8))
    return y
n = 512
xi = rng.integers(-1000, 1000, size=(n, D)).astype(np.float32)            # 11-bit values: only the hi part is non-zero
xf = (rng.standard_normal((n, D)) * np.exp(rng.uniform(-8, 8, size=(n, D)))).astype(np.float32)
I = np.eye(D, dtype=np.float32)
perm = synth.SHIPPED_REORDER_128
P = np.zeros((D, D), np.float32); P[np.arange(D), perm] = 1
run(I, xi, "identity, small ints ")
run(P, xi, "perm,     small ints ")
run(I, xf, "identity, full floats")
run(P, xf, "perm,     full floats")
x1 = np.ones((n, D), np.float32) * np.float32(1.0 + 2**-12 + 2**-23)
run(I, x1, "identity, 1+2^-12+2^-23")
Rd = synth.dense_rotation(D, 3)
y = run(Rd, xf / np.linalg.norm(xf, axis=1, keepdims=True), "dense R, unit rows   ")
