"""Quick timing of the front-end kernels (development aid, f-3): the PCA projection GEMM of both shipped model shapes
(1024 -> 128, 2048 -> 256) with the fused L2 normalisation, and rootSIFT, on device-resident rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cvt_b200 import capi

ctx = capi.Context(0)


def timeit(fn, it=10):
    fn(); ctx.synchronize()
    ctx.event_record(0)
    for _ in range(it):
        fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / it


rng = np.random.Generator(np.random.PCG64(1))
for K, N, n in ((1024, 128, 1 << 17), (2048, 256, 1 << 16)):
    V = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    mean = (rng.standard_normal(K) * 0.3).astype(np.float32)
    p = capi.Projection(ctx, V, mean)
    x = torch.relu(torch.randn((n, K), device="cuda"))
    y = torch.empty((n, N), device="cuda")
    ms = timeit(lambda: p.reduce_dim_dev(x.data_ptr(), n, y.data_ptr(), True))
    flop = 2.0 * n * K * N * 5
    print(f"proj {K}->{N} x {n} rows (+L2 norm): {ms:.3f} ms  {flop/ms/1e9:.1f} TF/s tf32 (5 products), "
          f"{(n*K*4 + n*N*4)/ms/1e6:.0f} GB/s of X+Y")
    y0 = torch.empty((n, N), device="cuda")
    p.reduce_dim_dev(x.data_ptr(), n, y0.data_ptr(), False)
    ref = (x[:8192].double() - torch.from_numpy(mean).cuda().double()) @ torch.from_numpy(V).cuda().double().T
    scale = torch.linalg.norm(x[:8192].double() - torch.from_numpy(mean).cuda().double(), dim=1, keepdim=True)
    print("   un-normalised projection vs fp64: max |err| / |x - mean| = %.3e" % float(((y0[:8192].double() - ref).abs() / scale).max()))
    nrm = torch.linalg.norm(y, dim=1)
    print("   row norms in [%.7f, %.7f]" % (float(nrm.min()), float(nrm.max())))
    p.close()
d = torch.floor(torch.rand((1 << 20, 128), device="cuda") * 200)
lib = capi.load()
import ctypes as C
ms = timeit(lambda: capi._check(lib.b200nn_rootsift_dev(ctx.h, C.c_void_p(d.data_ptr()), C.c_size_t(d.shape[0]), C.c_int(128), C.c_float(1e-7)), "rootsift_dev"))
print(f"rootsift 1M x 128 in place: {ms:.3f} ms  {2*d.numel()*4/ms/1e6:.0f} GB/s")
