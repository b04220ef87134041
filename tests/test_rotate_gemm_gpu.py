"""a1 as a dense rotation on the tcgen05 tensor cores (split-TF32 GEMM, csrc/rotate_gemm.cu).

* a PERMUTATION given as a dense 0/1 matrix must come out bit-identical to IVFOPQ::reorder
  (the error-free split makes every partial sum exactly representable);
* a dense orthonormal R must be at fp32 accuracy (compared with float64 and with the oracle's
  sequential fp32 dot products), far inside the 1e-4 tolerance north_star states for fp32;
* downstream of the rotation the pipeline stays bit-exact (codes / top-k from the GPU's own
  rotated rows equal the oracle's on those rows)."""
import numpy as np
import pytest

from cvt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from cvt_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _model(D, M, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = synth.sift_like(3000, D, seed=seed) if D == 128 else synth.cnn_like(3000, D, seed=seed)
    coarse, cb = synth.train_pq_model(x[:1500], M, 256, 1, iters=2, seed=seed, train_rows=1500)
    coarse = (rng.standard_normal((1, D)) * 0.01).astype(np.float32)
    return x, coarse, cb


@pytest.mark.parametrize("D,M,n", [(128, 16, 1000), (64, 8, 777), (512, 32, 300), (256, 32, 129)])
def test_permutation_as_dense_R_is_bit_exact(ctx, D, M, n):
    from cvt_b200 import capi
    x, coarse, cb = _model(D, M, 100 + D)
    perm = synth.SHIPPED_REORDER_128 if D == 128 else synth.random_permutation(D, seed=D)
    R = np.zeros((D, D), dtype=np.float32)
    R[np.arange(D), perm] = 1.0  # y[i] = x[perm[i]]
    idx = capi.PQIndex.create(ctx, coarse, cb, R=R)
    # values with full 24-bit mantissas, mixed signs and magnitudes
    rng = np.random.Generator(np.random.PCG64(D))
    xs = (x[:n] * rng.choice([-1.0, 1.0], size=(n, D)) * np.exp(rng.uniform(-8, 8, size=(n, D)))).astype(np.float32)
    y = idx.rotate(xs)
    ref = orc.opq_reorder(xs, perm)
    # bit-identical, except that a GEMM returns +0.0 where the gather copies a -0.0 input
    nz = ref != 0
    assert np.array_equal(y[nz].view(np.uint32), ref[nz].view(np.uint32))
    assert np.all(y[~nz] == 0)
    idx.close()


@pytest.mark.parametrize("D,M", [(128, 16), (64, 8), (256, 32)])
def test_dense_rotation_fp32_accuracy_and_pipeline(ctx, D, M):
    from cvt_b200 import capi
    x, coarse, cb = _model(D, M, 200 + D)
    R = synth.dense_rotation(D, seed=D + 5)
    idx = capi.PQIndex.create(ctx, coarse, cb, R=R, clamp=np.inf)
    y = idx.rotate(x)
    y64 = x.astype(np.float64) @ R.astype(np.float64).T
    scale = np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)
    err = np.abs(y - y64) / scale
    assert err.max() < 2e-6, err.max()  # fp32-level accumulation error (a plain TF32 GEMM would be ~5e-4)
    yo = orc.opq_rotate_dense(x, R)
    assert (np.abs(y - yo) / scale).max() < 2e-6
    # the rest of the pipeline on the GPU's own rotated rows is bit-exact
    idx.add(x)
    _, _, codes = idx.get_rows()
    assert np.array_equal(codes, orc.opq_pq_encode(y, coarse, np.zeros(len(y), np.int32), cb))
    q = x[:16] + 0.01
    Dg, Ig = idx.search(q, k=10)
    qy = idx.rotate(q)
    Do, Io = orc.opq_search_flat(qy, coarse[0], cb, codes, 10)
    assert np.array_equal(Ig.astype(np.int64), Io) and np.array_equal(Dg.view(np.uint32), Do.view(np.uint32))
    # against the all-oracle pipeline (oracle rotation): codes may differ only at near-ties
    codes_o = orc.opq_pq_encode(yo, coarse, np.zeros(len(yo), np.int32), cb)
    assert (codes != codes_o).mean() < 5e-3
    idx.close()


def test_dense_rotation_errors(ctx):
    from cvt_b200 import capi
    x, coarse, cb = _model(128, 16, 7)
    with pytest.raises(capi.B200nnError):  # both a permutation and R
        capi.PQIndex.create(ctx, coarse, cb, perm=synth.SHIPPED_REORDER_128, R=np.eye(128, dtype=np.float32))
    rng = np.random.Generator(np.random.PCG64(1))
    c96 = np.zeros((1, 96), np.float32)
    cb96 = rng.standard_normal((12, 256, 8)).astype(np.float32)
    with pytest.raises(capi.B200nnError):  # D % 64 != 0
        capi.PQIndex.create(ctx, c96, cb96, R=np.eye(96, dtype=np.float32))
