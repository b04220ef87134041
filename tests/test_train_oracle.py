"""CPU tests of the training restatement.  The reference trains with yael's kmeans
(train_PQ_codebook.cpp:164,229), un-vendored and randomly initialised => PARITY UNPINNED; what is checked here are
the properties of the deterministic Lloyd iteration the product defines (and the oracle restates): a known answer,
determinism, centroids = cluster means, a non-increasing quantisation error, the empty-cluster rule, and the
structure CoarseQuan -> residue -> ProdQuan of TrainPQ::IFVPQ."""
import numpy as np
import pytest

import cases
from oracle import oracle as orc


def test_kmeans_known_answer():
    x = np.array([[0.0], [1.0], [10.0], [11.0]], dtype=np.float32)
    for seed in range(6):
        c, a, d, it, mse = orc.kmeans(x, 2, 0, seed)
        assert sorted(c[:, 0].tolist()) == [0.5, 10.5]
        assert a[0] == a[1] and a[2] == a[3] and a[0] != a[2]
        assert np.array_equal(d, np.full(4, 0.25, np.float32)) and mse == 0.25


def test_kmeans_properties():
    x = cases.train_inputs(3000, 16)
    c, a, d, it, mse = orc.kmeans(x, 24, 0, 11)
    c2, a2, d2, it2, mse2 = orc.kmeans(x, 24, 0, 11)
    assert np.array_equal(c, c2) and np.array_equal(a, a2) and it == it2  # deterministic
    assert 0 < it < 10000
    for j in range(24):  # converged: every centroid is the mean of its rows (double accumulation, one rounding)
        m = x[a == j].astype(np.float64).mean(0)
        assert np.allclose(c[j], m, rtol=0, atol=1e-6)
    # assignment = first minimum of the sequential fp32 distance (IVFOPQ.cpp:107-129)
    assert np.array_equal(a, orc.opq_coarse_assign(x, c))
    # the quantisation error never increases with more updates, and a different seed gives a different start
    errs = [orc.kmeans(x, 24, i, 11)[4] for i in range(1, 8)]
    assert all(e1 >= e2 - 1e-12 for e1, e2 in zip(errs, errs[1:])) and errs[-1] >= mse - 1e-12
    assert not np.array_equal(orc.kmeans(x, 24, 1, 12)[0], orc.kmeans(x, 24, 1, 11)[0])


def test_kmeans_empty_clusters_are_reseeded():
    # 12 distinct points, 40 copies each, k = 20: the start picks duplicates, so clusters run empty and take the
    # farthest rows; at the end no two centroids that own rows coincide and every distinct point is a centroid
    pts = cases.train_inputs(12, 4, seed=3)
    x = np.repeat(pts, 40, axis=0)
    c, a, d, it, mse = orc.kmeans(x, 20, 0, 5)
    assert mse == 0.0 and len(np.unique(a)) == 12
    assert {tuple(r) for r in pts} <= {tuple(r) for r in c}
    with pytest.raises(ValueError):
        orc.kmeans(x[:5], 20, 0, 5)  # fewer rows than centroids


def test_pq_train_structure():
    x = cases.train_inputs(2500, 32, seed=9)
    perm = np.random.Generator(np.random.PCG64(1)).permutation(32).astype(np.int32)
    coarse, cb, mse = orc.pq_train(x, 6, 4, 32, perm=perm, max_iter=5, seed=21)
    xr = x[:, perm]
    c0, a0, d0, _, m0 = orc.kmeans(xr, 6, 5, 21)
    assert np.array_equal(coarse, c0) and mse[0] == m0
    res = xr - c0[a0]
    for m in range(4):
        cm, _, _, _, mm = orc.kmeans(res[:, m * 8:(m + 1) * 8], 32, 5, 22 + m)
        assert np.array_equal(cb[m], cm) and mse[1 + m] == mm
    # K = 0: flat-ADC model, one zero centroid, codebooks trained on the reordered rows themselves
    coarse0, cb0, mse0 = orc.pq_train(x, 0, 4, 32, perm=perm, max_iter=5, seed=21)
    assert coarse0.shape == (1, 32) and not coarse0.any() and mse0[0] == 0.0
    assert np.array_equal(cb0[1], orc.kmeans(xr[:, 8:16], 32, 5, 23)[0])


# ---------------------------------------------------------------------------------------------------------------
# The product's HOST bookkeeping (init rows, donors for empty clusters, stable counting sort: exported host-only by the
# C ABI) driven here with numpy standing in for the device kernels -- elementwise fp32 / fp64 numpy operations round
# exactly like the kernels' __fsub_rn/__fmul_rn/__fadd_rn/__dadd_rn -- and compared bit for bit with the oracle.
def _assign_np(x, c):
    """sequential fp32 squared distance to every centroid, first minimum wins (= kmeans_assign_kernel)."""
    n, d = x.shape
    acc = np.zeros((n, c.shape[0]), dtype=np.float32)
    for t in range(d):
        df = x[:, t:t + 1] - c[None, :, t]
        acc = acc + df * df
    a = acc.argmin(1).astype(np.int32)
    return a, acc[np.arange(n), a]


def _update_np(x, rows, off, count, block=512):
    """blocked double sums in list order (= kmeans_partial_kernel + kmeans_finalize_kernel)."""
    k, d = len(count), x.shape[1]
    c = np.empty((k, d), dtype=np.float32)
    for j in range(k):
        tot = np.zeros(d, dtype=np.float64)
        for lo in range(int(off[j]), int(off[j + 1]), block):
            part = np.zeros(d, dtype=np.float64)
            for r in rows[lo:min(lo + block, int(off[j + 1]))]:
                part = part + x[r].astype(np.float64)
            tot = tot + part
        c[j] = (tot / np.float64(count[j])).astype(np.float32)
    return c


def _kmeans_with_product_host_logic(x, k, max_iter, seed):
    from cvt_b200 import capi
    n = x.shape[0]
    c = x[capi.kmeans_init_rows(n, k, seed)].copy()
    cap = max_iter if max_iter > 0 else 10000
    prev, it = None, 0
    while True:
        a, dist = _assign_np(x, c)
        if (it > 0 and np.array_equal(a, prev)) or not dist.any() or it == cap:
            return c, a, dist, it
        a2, count, rows, off = capi.kmeans_plan_update(a, dist, k)
        assert count.sum() == n and count.min() >= 1 and np.array_equal(np.bincount(a2, minlength=k), count)
        for j in range(k):  # stable: rows of a cluster ascending
            seg = rows[off[j]:off[j + 1]]
            assert np.all(a2[seg] == j) and np.all(np.diff(seg) > 0)
        c = _update_np(x, rows, off, count)
        prev, it = a2, it + 1


@pytest.mark.parametrize("n,d,k,max_iter,seed,dup", [(1500, 8, 24, 0, 3, False), (1300, 5, 16, 4, 9, False),
                                                     (480, 4, 20, 0, 5, True), (64, 3, 64, 0, 1, False), (700, 16, 300, 3, 2, True)])
def test_product_host_bookkeeping_equals_oracle(n, d, k, max_iter, seed, dup):
    if dup:  # few distinct points: clusters run empty, donors move
        x = np.repeat(cases.train_inputs(max(n // 40, 12), d, seed=seed), 40, axis=0)[:n]
    else:
        x = cases.train_inputs(n, d, seed=seed)
    c, a, dist, it = _kmeans_with_product_host_logic(x, k, max_iter, seed)
    oc, oa, od, oit, _ = orc.kmeans(x, k, max_iter, seed)
    assert it == oit
    assert np.array_equal(c.view(np.uint32), oc.view(np.uint32))
    assert np.array_equal(a, oa) and np.array_equal(dist.view(np.uint32), od.view(np.uint32))


def test_host_bookkeeping_argument_checks():
    from cvt_b200 import capi
    with pytest.raises(capi.B200nnError):
        capi.kmeans_init_rows(3, 5, 0)  # fewer rows than centroids
    with pytest.raises(capi.B200nnError):
        capi.kmeans_plan_update(np.array([0, -1, 1], np.int32), np.zeros(3, np.float32), 2)  # a row without a centroid
    r = capi.kmeans_init_rows(1000, 1000, 7)
    assert sorted(r.tolist()) == list(range(1000))  # k = n: a permutation


def test_plan_update_random_cases_against_a_python_restatement():
    """The empty-cluster rule and the stable counting sort of b200nn_kmeans_plan_update on 200 random small assignments
    (many empty clusters, distance ties, clusters of one row), against a direct Python statement of the rule."""
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(0xE5))
    for case in range(200):
        k = int(rng.integers(1, 12))
        n = int(rng.integers(k, 40))
        used = rng.choice(k, size=int(rng.integers(1, k + 1)), replace=False)       # clusters that own rows
        assign = used[rng.integers(0, len(used), n)].astype(np.int32)
        dist = (rng.integers(0, 4, n) * 0.5).astype(np.float32)                      # ties are the norm
        a2, count, rows, off = capi.kmeans_plan_update(assign, dist, k)
        # ---- the rule, stated directly
        ea = assign.copy()
        cnt = np.bincount(ea, minlength=k)
        taken = np.zeros(n, bool)
        for j in range(k):
            if cnt[j]:
                continue
            best = -1
            for i in range(n):
                if not taken[i] and cnt[ea[i]] >= 2 and (best < 0 or dist[i] > dist[best]):
                    best = i
            taken[best] = True
            cnt[ea[best]] -= 1
            ea[best] = j
            cnt[j] = 1
        assert np.array_equal(a2, ea), case
        assert np.array_equal(count, cnt) and count.min() >= 1
        assert off[0] == 0 and off[-1] == n and np.array_equal(np.diff(off), cnt)
        assert np.array_equal(rows, np.argsort(ea, kind="stable")), case
