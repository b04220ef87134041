"""GPU parity tests of training through the C ABI.  yael's kmeans (the reference's trainer) is
un-vendored and randomly initialised => parity unpinned there; the product's deterministic Lloyd iteration must equal
its CPU restatement in oracle/ BIT FOR BIT (centroids, assignment, distances, iteration count), and the models it
writes must be the reference's file format and usable by the search path."""
import os
import subprocess

import numpy as np
import pytest

import cases
from cvt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
BIN = os.path.join(cases.ROOT, "tools", "bin")


@pytest.fixture(scope="module")
def ctx():
    from cvt_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("n,d,k,max_iter,seed,dup", [
    (4000, 8, 64, 0, 3, False),      # a PQ sub-space shape, to convergence
    (3000, 128, 16, 5, 9, False),    # a coarse-quantizer shape (4 rows per warp)
    (900, 400, 7, 3, 4, False),      # d > 384: one row per warp
    (777, 3, 5, 0, 1, False),        # ragged everything
    (480, 4, 20, 0, 5, True),        # 12 distinct points, k = 20: empty clusters, donors, stop on zero error
    (64, 6, 64, 0, 2, False),        # k = n
    (5000, 16, 1, 2, 8, False),      # k = 1: one cluster of 5000 rows = 10 summation blocks
    (2000, 64, 96, 3, 11, False),    # d >= 32 and k >= 64: the register-tiled distance kernel (dist_tile.cu)
    (1501, 40, 130, 2, 12, False),   # tiled, ragged: d = 16 + 16 + 8, two centroid tiles (ldm = 132), rows not a multiple of 128
    (9000, 32, 4100, 1, 6, False),   # tiled, two row chunks of the distance matrix (128 MB / (4100 * 4 B) -> 8064 rows)
])
def test_kmeans_bit_exact_vs_oracle(ctx, n, d, k, max_iter, seed, dup):
    from cvt_b200 import capi
    if dup:
        x = np.repeat(cases.train_inputs(12, d, seed=seed), 40, axis=0)[:n]
    else:
        x = cases.train_inputs(n, d, seed=seed)
    c, a, dist, it, mse = capi.kmeans(ctx, x, k, max_iter, seed)
    oc, oa, od, oit, omse = orc.kmeans(x, k, max_iter, seed)
    assert it == oit
    assert np.array_equal(a, oa)
    assert np.array_equal(_bits(c), _bits(oc))
    assert np.array_equal(_bits(dist), _bits(od))
    assert mse == omse


def test_kmeans_errors(ctx):
    from cvt_b200 import capi
    with pytest.raises(capi.B200nnError):
        capi.kmeans(ctx, np.zeros((3, 4), np.float32), 5)          # fewer rows than centroids
    x = cases.train_inputs(100, 4)
    x[7, 2] = np.nan
    with pytest.raises(capi.B200nnError):
        capi.kmeans(ctx, x, 3)                                       # a NaN row has no nearest centroid (B-2)


@pytest.mark.parametrize("K", [6, 0])
def test_pq_train_bit_exact_vs_oracle(ctx, K):
    """TrainPQ::IFVPQ: reorder -> CoarseQuan -> residue -> ProdQuan; K = 0 is the flat-ADC model."""
    from cvt_b200 import capi
    x = cases.train_inputs(2500, 32, seed=9)
    perm = np.random.Generator(np.random.PCG64(1)).permutation(32).astype(np.int32)
    coarse, cb, mse = capi.pq_train(ctx, x, K, 4, 32, perm=perm, max_iter=5, seed=21)
    oc, ocb, omse = orc.pq_train(x, K, 4, 32, perm=perm, max_iter=5, seed=21)
    assert np.array_equal(_bits(coarse), _bits(oc))
    assert np.array_equal(_bits(cb), _bits(ocb))
    assert np.array_equal(mse, omse)


def test_trained_model_quality_file_format_and_search(ctx, tmp_path):
    """Train the bench-shaped flat model (M=16 x 256, 128-d SIFT-shaped rows) on the device; the quantisation error must
    not be worse than the numpy Lloyd used for synthetic models so far; the written file is byte-identical to the
    reference layout and drives the search path like any other model."""
    from cvt_b200 import capi
    perm = synth.SHIPPED_REORDER_128
    db = synth.sift_like(20000, 128, seed=31)
    coarse, cb, mse = capi.pq_train(ctx, db, 0, 16, 256, perm=perm, max_iter=6, seed=synth.SEED_KMEANS)
    assert coarse.shape == (1, 128) and not coarse.any() and mse[0] == 0.0
    xr = db[:, perm]
    _, cb_np = synth.train_pq_model(xr, 16, 256, 1, iters=6, train_rows=20000)

    def qerr(cbs):
        codes = orc.opq_pq_encode(xr[:4000], coarse, np.zeros(4000, np.int32), cbs)
        rec = np.concatenate([cbs[m][codes[:, m]] for m in range(16)], axis=1)
        return float(((xr[:4000] - rec) ** 2).sum(1).mean())
    e_gpu, e_np = qerr(cb), qerr(cb_np)
    assert e_gpu <= 1.05 * e_np, (e_gpu, e_np)
    assert abs(float(mse[1:].sum()) - qerr(cb)) < 0.05 * e_gpu  # reported error = error of the encoded rows (same distribution)
    path = str(tmp_path / "trained.model")
    capi.pq_write_model(path, coarse, cb, perm)
    ref_path = str(tmp_path / "np.model")
    synth.write_opq_model(ref_path, coarse, cb, np.asarray(perm, dtype=np.int32))
    assert open(path, "rb").read() == open(ref_path, "rb").read()
    idx = capi.PQIndex.load_model(ctx, path)
    idx.add(db[:5000])
    q = synth.sift_like(16, 128, seed=32)
    dist, ids = idx.search(q, 10)
    _, _, codes = idx.get_rows()
    od, oi = orc.opq_search_flat(orc.opq_reorder(q, perm), coarse[0], cb, codes, 10, clamp=1.0)
    assert np.array_equal(ids.astype(np.int64), oi) and np.array_equal(_bits(dist), _bits(od))
    idx.close()


def test_unmodified_reference_train_main_on_gpu(ctx, tmp_path):
    """opq/train_codebook/train_PQ.cpp, compiled unmodified against include/b200nn/compat (the TrainPQ shim over
    b200nn_kmeans), must write exactly the model b200nn_pq_train + b200nn_pq_write_model produce."""
    from cvt_b200 import capi
    exe = os.path.join(BIN, "ref_train_PQ_on_b200nn")
    if not os.path.exists(exe):
        pytest.skip("ref_train_PQ_on_b200nn not built (needs /root/reference at build time)")
    assert b"b200nn_kmeans" in open(exe, "rb").read()
    D, K, M, ksub, n = 32, 5, 4, 16, 1200
    x = cases.train_inputs(n + 100, D, seed=13)
    perm = np.random.Generator(np.random.PCG64(2)).permutation(D)
    perm.astype(np.int64).tofile(str(tmp_path / "reorder.bin"))       # raw `long int[D]` (train_PQ_codebook.cpp:16-19)
    x.tofile(str(tmp_path / "train.bin"))
    env = dict(os.environ, B200NN_KMEANS_MAX_ITER="4", B200NN_KMEANS_SEED="77")
    r = subprocess.run([exe, str(tmp_path / "reorder.bin"), str(tmp_path / "train.bin"), str(tmp_path), str(n), str(K), str(D), str(M), str(ksub)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    out = tmp_path / f"OPQ_db_{n}_dim_{D}_k_{K}_PQ_m{M}_k{ksub}.fvecs"   # the reference's file-name scheme (:272)
    assert out.exists(), r.stdout
    coarse, cb, _ = capi.pq_train(ctx, x[:n], K, M, ksub, perm=perm.astype(np.int32), max_iter=4, seed=77)
    capi.pq_write_model(str(tmp_path / "direct.model"), coarse, cb, perm.astype(np.int32))
    assert out.read_bytes() == (tmp_path / "direct.model").read_bytes()
    c2, cb2, p2 = synth.read_opq_model(str(out))
    assert np.array_equal(p2, perm) and np.array_equal(_bits(cb2), _bits(cb))
