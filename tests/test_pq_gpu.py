"""GPU parity tests of the (O)PQ path, through the C ABI, against the golden vectors produced by the
unmodified reference and against the oracle restatement.  Bit-exact: uint8 codes, list ids,
neighbour ids AND fp32 scores (the kernels reproduce the reference's summation order)."""
import os

import numpy as np
import pytest

import cases
from cvt_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
G = cases.GOLDEN


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx():
    from cvt_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _fixture_files():
    db = [np.fromfile(os.path.join(G, "opq_fixture", "db", f), dtype="<f4").reshape(-1, 128) for f in cases.FIXTURE_DB]
    q = [np.fromfile(os.path.join(G, "opq_fixture", "query", f), dtype="<f4").reshape(-1, 128) for f in cases.FIXTURE_QUERY]
    return db, q


def test_shipped_fixture_reduced_model(ctx):
    """opq/data fixtures + shipped codebooks: IndexDatabase + QueryThrehold."""
    from cvt_b200 import capi
    gold = np.load(os.path.join(G, "opq_shipped_k256.npz"))
    idx = capi.PQIndex.load_model(ctx, os.path.join(G, "opq_shipped_k256.model"))
    assert (idx.D, idx.K, idx.M, idx.ksub) == (128, 256, 16, 256)
    db, q = _fixture_files()
    for gi, rows in enumerate(db):  # one videoId per file, IVFOPQ.cpp:198-201
        idx.add(rows, np.full(len(rows), gi, dtype=np.int32))
    assert idx.n_rows == 53 and idx.n_groups == 5
    lists, groups, codes = idx.get_rows()
    assert np.array_equal(lists, gold["row_list"])
    assert np.array_equal(groups, gold["row_group"])
    assert np.array_equal(codes, gold["codes"])
    qall = np.concatenate(q)
    match = idx.scores(qall, nprobe=3)
    assert np.array_equal(_bits(match), _bits(gold["match"]))
    # multi_frame_index_test.cpp:56-68: frame-summed scores + get_sort_results, per query file
    off = 0
    for fi, rows in enumerate(q):
        total = np.zeros(5, dtype=np.float32)
        for f in range(len(rows)):
            total = total + match[off + f]
        off += len(rows)
        s, i = orc.topk_pairs(total, 5)
        assert np.array_equal(i, gold["file_topk_id"][fi])
        assert np.array_equal(_bits(s), _bits(gold["file_topk_score"][fi]))
    # the same on the device (f-5): frame sums + get_sort_results, only [n_videos, k] results come back
    frame_off = np.cumsum([0] + [len(rows) for rows in q])
    S5, G5 = idx.query_groups(qall, frame_off, k=5, nprobe=3)
    assert np.array_equal(G5.astype(np.int64), gold["file_topk_id"])
    assert np.array_equal(_bits(S5), _bits(gold["file_topk_score"]))
    S7, G7 = idx.query_groups(qall, frame_off, k=7, nprobe=3)  # more than the 5 indexed videos: padded tail
    assert np.array_equal(G7[:, :5], G5) and np.all(np.isinf(S7[:, 5:])) and np.all(G7[:, 5:] == np.uint64(0xFFFFFFFFFFFFFFFF))
    idx.close()

    # per-row ids (videoId = row): top-k through the generic IVF path (K = 256, nprobe = 3)
    idx = capi.PQIndex.load_model(ctx, os.path.join(G, "opq_shipped_k256.model"))
    idx.add(np.concatenate(db))
    D, I = idx.search(qall, k=int(gold["pr_topk"]), nprobe=3)
    assert np.array_equal(I.astype(np.int64), gold["pr_topk_id"])
    assert np.array_equal(_bits(D), _bits(gold["pr_topk_score"]))
    idx.close()


@pytest.mark.parametrize("name", list(cases.OPQ_CASES))
def test_opq_synthetic_vs_reference_golden(ctx, name):
    from cvt_b200 import capi
    c = cases.opq_case(name)
    gold = np.load(os.path.join(G, f"opq_{name}.npz"))
    assert str(gold["input_sha"]) == c["input_sha"]
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"], clamp=1.0)
    rpg = c["rows_per_group"]
    groups = None if rpg == 1 else (np.arange(c["n"]) // rpg).astype(np.int32)
    idx.add(c["db"], groups)
    lists, grp, codes = idx.get_rows()
    assert np.array_equal(lists, gold["row_list"])
    assert np.array_equal(grp, gold["row_group"])
    assert np.array_equal(codes, gold["codes"])
    match = idx.scores(c["q"], nprobe=c["nk"])
    assert np.array_equal(_bits(match), _bits(gold["match"]))
    if rpg == 1:
        D, I = idx.search(c["q"], k=int(gold["topk"]), nprobe=c["nk"])
        assert np.array_equal(I.astype(np.int64), gold["topk_id"]), name
        assert np.array_equal(_bits(D), _bits(gold["topk_score"])), name
    idx.close()


def test_query_groups_vs_oracle(ctx):
    """f-5 on synthetic data: 40 indexed videos of 50 frames, query videos of 1..260 frames (several frame
    chunks), IVF probing; device frame sums + top-k == the oracle's QueryThrehold scores summed in frame order."""
    from cvt_b200 import capi
    c = cases.opq_case("ivf_m8")
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"], clamp=1.0)
    groups = (np.arange(c["n"]) // 50).astype(np.int32)
    idx.add(c["db"], groups)
    lists, grp, codes = idx.get_rows()
    lens = [1, 12, 260, 0, 3]
    q = np.concatenate([c["q"]] * 23)[: sum(lens)] * np.float32(1.0)
    q[20:40] *= np.float32(2.5)  # some frames over the clamp everywhere
    frame_off = np.cumsum([0] + lens)
    k = 10
    S, G = idx.query_groups(q, frame_off, k=k, nprobe=3)
    dense = orc.opq_query_scores(orc.opq_reorder(q, c["reorder"]), c["coarse"], c["cb"], 3, lists, grp, codes, idx.n_groups, 1.0)
    for v in range(len(lens)):
        total = np.zeros(idx.n_groups, dtype=np.float32)
        for f in range(frame_off[v], frame_off[v + 1]):
            total = total + dense[f]
        s, i = orc.topk_pairs(total, k)
        assert np.array_equal(G[v].astype(np.int64), i), v
        assert np.array_equal(_bits(S[v]), _bits(s)), v
    idx.close()


def test_rotate_encode_lut_vs_oracle(ctx):
    from cvt_b200 import capi
    c = cases.opq_case("ivf_m8")
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
    xr = idx.rotate(c["db"])
    assert np.array_equal(_bits(xr), _bits(orc.opq_reorder(c["db"], c["reorder"])))
    lists, codes = idx.encode(xr)
    ol = orc.opq_coarse_assign(xr, c["coarse"])
    assert np.array_equal(lists, ol)
    assert np.array_equal(codes, orc.opq_pq_encode(xr, c["coarse"], ol, c["cb"]))
    qr = idx.rotate(c["q"])
    probes, lut = idx.build_lut(qr, nprobe=3)
    for i in range(len(qr)):
        op = orc.opq_coarse_probe(qr[i], c["coarse"], 3)
        assert np.array_equal(probes[i], op)
        for p in range(3):
            assert np.array_equal(_bits(lut[i, p]), _bits(orc.opq_build_lut(qr[i], c["coarse"][op[p]], c["cb"])))
    idx.close()


def _random_flat_index(ctx, n, D, M, seed, clamp=np.inf):
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(seed))
    db = synth.sift_like(n, D, seed=seed)
    perm = synth.SHIPPED_REORDER_128 if D == 128 else synth.random_permutation(D, seed)
    coarse, cb = synth.train_pq_model(db[:2000][:, perm], M, 256, 1, iters=2, seed=seed + 1, train_rows=2000)
    coarse = (rng.standard_normal((1, D)) * 0.01).astype(np.float32)
    idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=clamp)
    idx.add(db)
    return idx, db, perm, coarse, cb


@pytest.mark.parametrize("M,n,nq,k", [(16, 100_003, 67, 100), (8, 20_000, 19, 10), (32, 30_011, 21, 128), (4, 5_000, 40, 1),
                                       (16, 50, 9, 100), (16, 4097, 8, 7)])
def test_fast_scan_vs_oracle_and_generic(ctx, M, n, nq, k):
    """TMA-staged conflict-free scan == oracle restatement (ids and score bits), including ragged
    sizes (n % 64 != 0, nq % queries-per-CTA != 0, n < k) and every supported M."""
    D = 128
    idx, db, perm, coarse, cb = _random_flat_index(ctx, n, D, M, seed=1000 + M + n)
    q = synth.sift_like(nq, D, seed=77 + M)
    Dg, Ig = idx.search(q, k=k, nprobe=1)
    _, _, codes = idx.get_rows()
    xr = orc.opq_reorder(db, perm)
    assert np.array_equal(codes, orc.opq_pq_encode(xr, coarse, np.zeros(n, np.int32), cb))
    qr = orc.opq_reorder(q, perm)
    check = range(nq) if n <= 30_000 else range(0, nq, 8)
    for i in check:
        lut = orc.opq_build_lut(qr[i], coarse[0], cb)
        s = orc.opq_adc_scan(lut, codes)
        os_, oi = orc.topk_pairs(s, k)
        kk = min(k, n)
        assert np.array_equal(Ig[i, :kk].astype(np.int64), oi[:kk]), (M, n, i)
        assert np.array_equal(_bits(Dg[i, :kk]), _bits(os_[:kk])), (M, n, i)
        if kk < k:
            assert np.all(np.isinf(Dg[i, kk:])) and np.all(Ig[i, kk:] == np.uint64(0xFFFFFFFFFFFFFFFF))
    idx.close()


@pytest.mark.parametrize("M,n,nq,k", [(16, 20_000, 801, 100), (16, 20_000, 1424, 10), (32, 12_000, 720, 100), (8, 9_000, 2500, 16),
                                       (4, 6_000, 5000, 5)])
def test_fast_scan_tail_pieces_straddle_query_groups(ctx, M, n, nq, k):
    """Batches whose last wave is cut into equal pieces that straddle query-group boundaries (a tail CTA
    then runs two segments: two LUTs, two top-k lists, two output slices) == the oracle, every query."""
    from cvt_b200 import capi
    D = 128
    n_full, n_tail, slices, desc = capi.scan_plan(148, M, nq, n)
    assert n_tail > 0 and np.any(desc[:, 1, 3] > desc[:, 1, 2]), "the case must contain two-segment CTAs"
    idx, db, perm, coarse, cb = _random_flat_index(ctx, n, D, M, seed=2000 + M + n, clamp=1.0)
    q = synth.sift_like(nq, D, seed=99 + M)
    Dg, Ig = idx.search(q, k=k, nprobe=1)
    _, _, codes = idx.get_rows()
    od, oi = orc.opq_search_flat(orc.opq_reorder(q, perm), coarse[0], cb, codes, k, clamp=1.0)
    assert np.array_equal(Ig.astype(np.int64), oi)
    assert np.array_equal(_bits(Dg), _bits(od))
    # and again with a different batch size on the same index (the cached plan must follow the shape)
    D2, I2 = idx.search(q[: nq // 3], k=k, nprobe=1)
    assert np.array_equal(I2.astype(np.int64), oi[: nq // 3]) and np.array_equal(_bits(D2), _bits(od[: nq // 3]))
    idx.close()


def test_clamp_ties_fast_scan(ctx):
    """threhold = 1.0 (IVFOPQ.cpp:5): scores >= 1 collapse to exactly 1.0 and ids break the ties."""
    n, D, M, k = 3000, 128, 16, 64
    idx, db, perm, coarse, cb = _random_flat_index(ctx, n, D, M, seed=4242, clamp=1.0)
    q = synth.sift_like(12, D, seed=5) * np.float32(2.0)
    Dg, Ig = idx.search(q, k=k, nprobe=1)
    _, _, codes = idx.get_rows()
    qr = orc.opq_reorder(q, perm)
    ntie = 0
    for i in range(len(q)):
        s = np.minimum(orc.opq_adc_scan(orc.opq_build_lut(qr[i], coarse[0], cb), codes), np.float32(1.0))
        os_, oi = orc.topk_pairs(s, k)
        assert np.array_equal(Ig[i].astype(np.int64), oi)
        assert np.array_equal(_bits(Dg[i]), _bits(os_))
        ntie += int((os_ == 1.0).sum())
    assert ntie > 0, "the case must exercise clamp ties"
    idx.close()


def test_empty_and_errors(ctx):
    from cvt_b200 import capi
    c = cases.opq_case("flat_m16")
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
    D, I = idx.search(c["q"][:3], k=5)
    assert np.all(np.isinf(D)) and np.all(I == np.uint64(0xFFFFFFFFFFFFFFFF))
    with pytest.raises(capi.B200nnError):
        idx.search(c["q"][:3], k=129)
    with pytest.raises(capi.B200nnError):
        idx.search(c["q"][:3], k=5, nprobe=2)  # K == 1
    with pytest.raises(capi.B200nnError):
        capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=np.zeros(128, np.int32))  # not a permutation
    with pytest.raises(capi.B200nnError):
        capi.PQIndex.load_model(ctx, "/nonexistent.model")
    idx.close()


def test_save_load_index_roundtrip(ctx, tmp_path):
    from cvt_b200 import capi
    c = cases.opq_case("ivf_m8")
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
    groups = (np.arange(c["n"]) // 50).astype(np.int32)
    idx.add(c["db"], groups)
    ref_scores = idx.scores(c["q"], nprobe=3)
    idx.save_index(str(tmp_path), [f"/data/video_{i}.bin" for i in range(idx.n_groups)])
    path = tmp_path / f"OPQ_Index_db_{idx.n_groups}_dim_64_k_16_PQ_m8_k256.fvecs"
    assert path.exists()
    # byte layout of IVFOPQ::SaveIndex (IVFOPQ.cpp:544-580)
    raw = path.read_bytes()
    hdr = np.frombuffer(raw, dtype="<i4", count=5)
    assert hdr.tolist() == [64, 16, 8, 256, idx.n_groups]
    expect = 20 + 4 * (16 * 64 + 8 * 256 * 8) + 16 * 4 + c["n"] * (4 + 8) + idx.n_groups * 260
    assert len(raw) == expect
    idx2 = capi.PQIndex.load_index(ctx, str(path), perm=c["reorder"])
    assert idx2.n_rows == c["n"] and idx2.n_groups == idx.n_groups
    assert np.array_equal(_bits(idx2.scores(c["q"], nprobe=3)), _bits(ref_scores))
    idx.close(); idx2.close()


@pytest.mark.parametrize("clamp", [1.0, np.inf])
def test_ivf_fused_search_vs_oracle(ctx, clamp):
    """f-1: true IVF probing (K=16 lists, nprobe=3) with per-row ids through the fused LUT+scan+top-k
    kernel == the reference semantics restated by the oracle (dense clamp-initialised scores, then
    get_sort_results), including the clamp-valued tail that is filled by id order."""
    from cvt_b200 import capi
    c = cases.opq_case("ivf_m8")
    idx = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"], clamp=clamp)
    idx.add(c["db"])  # videoId = row
    lists, groups, codes = idx.get_rows()
    assert np.array_equal(groups, np.arange(c["n"], dtype=np.int32))
    q = np.concatenate([c["q"], c["q"] * np.float32(3.0)])  # the scaled copies push most scores over the clamp
    k = 50
    Dg, Ig = idx.search(q, k=k, nprobe=3)
    qr = orc.opq_reorder(q, c["reorder"])
    dense = orc.opq_query_scores(qr, c["coarse"], c["cb"], 3, lists, groups, codes, c["n"], clamp)
    ntail = 0
    for i in range(len(q)):
        os_, oi = orc.topk_pairs(dense[i], k)
        assert np.array_equal(Ig[i].astype(np.int64), oi), i
        assert np.array_equal(_bits(Dg[i]), _bits(os_)), i
        ntail += int((os_ == np.float32(clamp)).sum()) if np.isfinite(clamp) else int(np.isinf(os_).sum())
    assert ntail > 0 or not np.isfinite(clamp)  # the finite clamp case must exercise the id-ordered tail
    idx.close()


def test_coarse_assign_and_probes_large_K_tiled(ctx):
    """a2 + a4 on the register-tiled distance kernel (K >= 64): ragged K (not a multiple of the 128-centroid tile or of 4),
    rows not a multiple of 128, DUPLICATED centroids (exact distance ties: the lowest index must win, IVFOPQ.cpp:123) and
    probes in the reference's pop order."""
    from cvt_b200 import capi
    D, M, K, n, nq, nk = 64, 8, 1003, 3001, 77, 5
    rng = np.random.Generator(np.random.PCG64(77))
    x = synth.sift_like(n, D, seed=71)
    q = synth.sift_like(nq, D, seed=72)
    perm = synth.random_permutation(D, 73)
    coarse = x[rng.choice(n, K, replace=False)][:, perm].copy()
    coarse[500:520] = coarse[100:120]            # twenty exact duplicates at higher indices
    cb = (rng.standard_normal((M, 256, D // M)) * 0.05).astype(np.float32)
    idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm)
    xr = idx.rotate(x)
    lists, codes = idx.encode(xr)
    ol = orc.opq_coarse_assign(xr, coarse)
    assert np.array_equal(lists, ol)
    assert not np.isin(lists, np.arange(500, 520)).any() and np.isin(np.arange(100, 120), lists).any()
    assert np.array_equal(codes, orc.opq_pq_encode(xr, coarse, ol, cb))
    qr = idx.rotate(q)
    probes, lut = idx.build_lut(qr, nprobe=nk)
    for i in range(nq):
        assert np.array_equal(probes[i], orc.opq_coarse_probe(qr[i], coarse, nk)), i
    idx.close()


def test_reference_index_shape_K8192_end_to_end(ctx):
    """The reference's real index shape on the GPU: K = 8192 coarse centroids, M = 16, D = 128, nprobe = 3 (the shipped
    opq/model/*_k_8192_PQ_m16_k256_reorder.model has exactly this shape; the model file itself stays in /root/reference, so the
    centroids here are seeded).  Add -> coarse lists + codes, QueryThrehold scores per videoId and the fused top-k search are
    compared with the restatement bit for bit."""
    from cvt_b200 import capi
    D, M, K, n, nq, nk, k = 128, 16, 8192, 12_000, 24, 3, 40
    rng = np.random.Generator(np.random.PCG64(8192))
    x = synth.sift_like(n, D, seed=81)
    q = np.concatenate([x[:8] * np.float32(1.0), synth.sift_like(nq - 8, D, seed=82)])   # 8 queries are database rows
    perm = synth.SHIPPED_REORDER_128
    coarse = np.concatenate([x[rng.choice(n, 4096, replace=False)][:, perm],
                             synth.sift_like(4096, D, seed=83)[:, perm]]).astype(np.float32)
    cb = (rng.standard_normal((M, 256, D // M)) * 0.03).astype(np.float32)
    groups = (np.arange(n) // 12).astype(np.int32)
    idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
    idx.add(x, groups)
    lists, grp, codes = idx.get_rows()
    xr = orc.opq_reorder(x, perm)
    ol = orc.opq_coarse_assign(xr, coarse)
    assert np.array_equal(lists, ol) and np.array_equal(grp, groups)
    assert np.array_equal(codes, orc.opq_pq_encode(xr, coarse, ol, cb))
    qr = orc.opq_reorder(q, perm)
    S = idx.scores(q, nprobe=nk)
    So = orc.opq_query_scores(qr, coarse, cb, nk, lists, groups, codes, idx.n_groups, 1.0)
    assert np.array_equal(_bits(S), _bits(So))
    assert (S[:8].min(axis=1) < 0.5).all()       # a database row finds its own video
    idx2 = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
    idx2.add(x)                                   # videoId = row: the per-vector top-k path
    Dg, Ig = idx2.search(q, k, nprobe=nk)
    dense = orc.opq_query_scores(qr, coarse, cb, nk, lists, np.arange(n, dtype=np.int32), codes, n, 1.0)
    for i in range(nq):
        os_, oi = orc.topk_pairs(dense[i], k)
        assert np.array_equal(Ig[i].astype(np.int64), oi), i
        assert np.array_equal(_bits(Dg[i]), _bits(os_)), i
    idx.close(); idx2.close()


def test_ivf_long_lists_buffer_overflow_path(ctx):
    """f-1 with LONG lists (K = 4 lists over 20 000 rows, nprobe = 3, no clamp: ~15 000 candidates per query): the CTA-wide
    candidate buffer of ivf_search_topk_kernel overflows many times and is sorted/trimmed on the way (cta_sort_trim), and the
    LUTs of the three probes are built in one codebook pass == the oracle's dense restatement, ids and score bits."""
    from cvt_b200 import capi
    n, D, M, K, k = 20_000, 64, 8, 4, 100
    db = synth.sift_like(n, D, seed=31)
    q = synth.sift_like(24, D, seed=32)
    perm = synth.random_permutation(D, seed=33)
    coarse, cb = synth.train_pq_model(db[:, perm], M, 256, K, iters=3, seed=34, train_rows=1500)
    for clamp in (np.inf, 1.0):
        idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=clamp)
        idx.add(db)
        lists, groups, codes = idx.get_rows()
        assert np.bincount(lists, minlength=K).max() > 2048, "the case must hold lists longer than the candidate buffer"
        qr = orc.opq_reorder(q, perm)
        for nprobe in (1, 3):
            Dg, Ig = idx.search(q, k=k, nprobe=nprobe)
            dense = orc.opq_query_scores(qr, coarse, cb, nprobe, lists, groups, codes, n, clamp)
            for i in range(len(q)):
                os_, oi = orc.topk_pairs(dense[i], k)
                assert np.array_equal(Ig[i].astype(np.int64), oi), (clamp, nprobe, i)
                assert np.array_equal(_bits(Dg[i]), _bits(os_)), (clamp, nprobe, i)
        idx.close()


@pytest.mark.parametrize("M,n,nq,k", [(16, 70_001, 300, 100), (32, 40_003, 130, 64), (8, 30_000, 77, 10)])
def test_scan_kernel_variants_agree(ctx, M, n, nq, k, monkeypatch):
    """B200NN_SCAN_VAR: round 1's per-block threshold test (0), the two-slot code ring (bit 0, M = 32 only) and the grouped
    test (bit 1, the default) are the same function of their inputs: ids and score bits identical, and equal to the oracle."""
    D = 128
    idx, db, perm, coarse, cb = _random_flat_index(ctx, n, D, M, seed=3000 + M, clamp=1.0)
    q = synth.sift_like(nq, D, seed=55 + M)
    res = {}
    for var in ((0, 1, 2, 3) if M == 32 else (0, 2)):
        monkeypatch.setenv("B200NN_SCAN_VAR", str(var))
        res[var] = idx.search(q, k=k, nprobe=1)
    monkeypatch.delenv("B200NN_SCAN_VAR")
    res["default"] = idx.search(q, k=k, nprobe=1)
    _, _, codes = idx.get_rows()
    od, oi = orc.opq_search_flat(orc.opq_reorder(q[:16], perm), coarse[0], cb, codes, k, clamp=1.0)
    for var, (Dg, Ig) in res.items():
        assert np.array_equal(Ig, res[0][1]) and np.array_equal(_bits(Dg), _bits(res[0][0])), var
        assert np.array_equal(Ig[:16].astype(np.int64), oi) and np.array_equal(_bits(Dg[:16]), _bits(od)), var
    idx.close()
