"""CPU tests: the oracle restatement (oracle/cvt_oracle.c) against the committed golden vectors,
which were produced by the UNMODIFIED reference compiled in place (oracle/gen_golden.py), and --
when this container still has /root/reference and oracle/_ref -- against the reference itself."""
import hashlib
import os

import numpy as np
import pytest

import cases
from cvt_b200 import synth
from oracle import oracle as orc

G = cases.GOLDEN


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _check_opq(gold, prefix, coarse, cb, reorder, db_rows, q_rows, nk):
    g = lambda k: gold[prefix + k]
    x = orc.opq_reorder(db_rows, reorder)
    lists = orc.opq_coarse_assign(x, coarse)
    assert np.array_equal(lists, g("row_list"))
    codes = orc.opq_pq_encode(x, coarse, lists, cb)
    assert np.array_equal(codes, g("codes"))
    qr = orc.opq_reorder(q_rows, reorder)
    n_groups = int(gold[prefix + "n_groups"])
    match = orc.opq_query_scores(qr, coarse, cb, nk, lists, g("row_group"), codes, n_groups, 1.0)
    assert np.array_equal(_bits(match), _bits(g("match")))
    k = int(gold[prefix + "topk"])
    for f in range(match.shape[0]):
        s, i = orc.topk_pairs(match[f], k)
        assert np.array_equal(i, g("topk_id")[f])
        assert np.array_equal(_bits(s), _bits(g("topk_score")[f]))


def _fixture_rows():
    db = np.concatenate([np.fromfile(os.path.join(G, "opq_fixture", "db", f), dtype="<f4").reshape(-1, 128)
                         for f in cases.FIXTURE_DB])
    q = np.concatenate([np.fromfile(os.path.join(G, "opq_fixture", "query", f), dtype="<f4").reshape(-1, 128)
                        for f in cases.FIXTURE_QUERY])
    return db, q


def test_shipped_fixture_reduced_model():
    gold = np.load(os.path.join(G, "opq_shipped_k256.npz"))
    coarse, cb, reorder = synth.read_opq_model(os.path.join(G, "opq_shipped_k256.model"))
    db, q = _fixture_rows()
    assert db.shape == (53, 128) and q.shape == (10, 128)
    _check_opq(gold, "", coarse, cb, reorder, db, q, 3)
    _check_opq(gold, "pr_", coarse, cb, reorder, db, q, 3)
    # SURVEY.md App. C known answers (produced by the reference on the FULL shipped model; the
    # reduced model reproduces them because it keeps every centroid that can win)
    assert hashlib.sha256(gold["codes"].tobytes()).hexdigest() == \
        "9896a9b07826dadf9ae06a9159a2d37e2a1712bfc1d5456163de41e37338c3c7"
    assert gold["codes"][0].tolist() == [232, 88, 97, 41, 245, 248, 61, 222, 180, 193, 190, 48, 7, 128, 24, 86]
    np.testing.assert_allclose(gold["match"][2], [0.13036789, 0.766138613, 0.670435786, 0.666301548, 0.827702045],
                               rtol=1e-7)
    assert gold["file_topk_id"][0].tolist() == [0, 2, 3, 1, 4]
    np.testing.assert_allclose(gold["file_topk_score"][0], [1.25004315, 7.37535906, 7.3803978, 7.61460876, 7.98865891],
                               rtol=1e-7)
    assert gold["file_topk_id"][1].tolist() == [0, 1, 2, 3, 4]  # clamp ties at 1.0 broken by id


@pytest.mark.skipif(not os.path.isdir(orc.REFERENCE_ROOT), reason="full shipped model only in the build container")
def test_shipped_fixture_full_model():
    gold = np.load(os.path.join(G, "opq_shipped_full.npz"))
    model = os.path.join(orc.REFERENCE_ROOT, "opq/model/OPQ_db_5950000_dim_128_k_8192_PQ_m16_k256_reorder.model")
    assert hashlib.sha256(open(model, "rb").read()).hexdigest() == str(gold["model_sha256"])
    coarse, cb, reorder = synth.read_opq_model(model)
    db, q = _fixture_rows()
    _check_opq(gold, "", coarse, cb, reorder, db, q, 3)
    assert gold["row_list"][:9].tolist() == [1955, 1955, 1870, 1870, 1870, 5812, 5812, 71, 6987]


@pytest.mark.parametrize("name", list(cases.OPQ_CASES))
def test_opq_synthetic(name):
    c = cases.opq_case(name)
    gold = np.load(os.path.join(G, f"opq_{name}.npz"))
    assert str(gold["input_sha"]) == c["input_sha"], "input generator drifted; regenerate goldens"
    _check_opq(gold, "", c["coarse"], c["cb"], c["reorder"], c["db"], c["q"], c["nk"])


@pytest.mark.parametrize("name", list(cases.FLAT_CASES))
def test_flat(name):
    c = cases.flat_case(name)
    gold = np.load(os.path.join(G, f"flat_{name}.npz"))
    assert str(gold["input_sha"]) == c["input_sha"]
    runs = {"ip_sse": (0, 4), "ip_hnsw": (0, 4), "ip_avx": (0, 8), "l2_avx": (1, 8), "l2_sse": (1, 4), "l2i": (2, 0)}
    seen = 0
    for tag, (metric, lanes) in runs.items():
        if tag + "_dist" not in gold:
            continue
        seen += 1
        data, q = (c["xu"], c["qu"]) if metric == 2 else (c["x"], c["q"])
        d, l = orc.flat_search(metric, lanes, data, c["labels"], q, c["k"])
        assert np.array_equal(l, gold[tag + "_label"]), tag
        assert np.array_equal(_bits(d), _bits(gold[tag + "_dist"])), tag
    assert seen >= 4


def test_flat_tie_rule_is_lexicographic():
    # brutoforce.hpp:81-91 keeps the k smallest under (dist, label)
    c = cases.flat_case("ties_d16")
    d, l = orc.flat_search(0, 4, c["x"], c["labels"], c["q"], c["k"])
    for i in range(c["nq"]):
        alld = np.array([orc.lib().orc_flat_ip(orc._p(c["q"][i]), orc._p(c["x"][j]), 16, 4) for j in range(c["n"])],
                        dtype=np.float32)
        order = np.lexsort((c["labels"], alld))[:c["k"]]
        assert np.array_equal(c["labels"][order], l[i])


@pytest.mark.parametrize("d", [64, 128])
def test_sq_restatement_vs_reference_golden(d):
    # codes / x_normed / decode in the golden are REFERENCE-RUN: produced by the unmodified int8_quan.cc compiled against
    # stand-ins for faiss / nlohmann (oracle/Makefile, oracle/gen_golden.py::gen_sq).  decode_faiss (faiss's own
    # all-float decode) stays unpinned: faiss is absent, the golden there is the restatement's own output.
    assert bool(np.load(os.path.join(G, f"sq_d{d}.npz"))["pinned"])
    c = cases.sq_case(d)
    gold = np.load(os.path.join(G, f"sq_d{d}.npz"))
    assert str(gold["input_sha"]) == c["input_sha"]
    codes, xn = orc.sq_encode(c["x"], c["vmin"], c["vdiff"], True)
    assert np.array_equal(codes, gold["codes"])
    assert np.array_equal(_bits(xn), _bits(gold["x_normed"]))
    dec = orc.sq_decode(codes, c["vmin"], c["vdiff"])
    assert np.array_equal(_bits(dec), _bits(gold["decode"]))
    decf = orc.sq_decode(codes, c["vmin"], c["vdiff"], faiss_float=True)
    assert np.array_equal(_bits(decf), _bits(gold["decode_faiss"]))
    # properties: all-zero row stays zero; constant dim (vdiff==0) encodes 0; decode error <= half a bucket
    assert np.all(xn[1] == 0) and np.all(codes[:, 3] == 0)
    inside = (xn >= c["vmin"]) & (xn <= c["vmin"] + c["vdiff"])
    err = np.abs(dec - xn)[inside]
    bucket = np.broadcast_to(c["vdiff"] / 255.0, xn.shape)[inside]
    assert np.all(err <= bucket * 1.0 + 1e-6)
    vmin, vdiff = orc.sq_train_minmax(xn)
    assert np.array_equal(vmin, xn.min(0)) and np.array_equal(vdiff, xn.max(0) - xn.min(0))


@pytest.mark.skipif(not orc.have_ref("ref_opq"), reason="oracle/_ref not built")
def test_reference_binary_agrees_with_golden(tmp_path):
    # the compiled reference travels to the GPU box; make sure it still reproduces the goldens there
    gold = np.load(os.path.join(G, "opq_shipped_k256.npz"))
    gdb = [os.path.join(G, "opq_fixture", "db", f) for f in cases.FIXTURE_DB]
    gq = [os.path.join(G, "opq_fixture", "query", f) for f in cases.FIXTURE_QUERY]
    r = orc.run_ref_opq(os.path.join(G, "opq_shipped_k256.model"), gdb, gq, nk=3, topk=5, per_row=False)
    assert np.array_equal(r["codes"], gold["codes"])
    assert np.array_equal(_bits(r["match"]), _bits(gold["match"]))


def test_frontend_restatement_equals_cv2_golden():
    """f-3: the restatements of cvtk::PCAUtils::reduceDim and siftsIDX::rootSift against the committed cv2 outputs
    (oracle/gen_golden_frontend.py; cv2 is the dependency the reference calls, the model is the reference's own)."""
    gp = np.load(os.path.join(cases.GOLDEN, "frontend_pca.npz"))
    x = cases.frontend_pca_inputs(1024)
    assert str(gp["input_sha"]) == cases.sha(x)
    y = orc.pca_project(x, gp["mean"], gp["vectors"], True)
    assert np.array_equal(y.view(np.uint32), gp["y"].view(np.uint32))
    assert np.allclose(np.linalg.norm(y, axis=1), 1.0, atol=1e-6)
    gr = np.load(os.path.join(cases.GOLDEN, "frontend_rootsift.npz"))
    d = cases.frontend_sift_inputs()
    assert str(gr["input_sha"]) == cases.sha(d)
    r = orc.rootsift(d)
    assert np.array_equal(r.view(np.uint32), gr["y"].view(np.uint32))
    assert np.all(r[3] == 0)  # the all-zero descriptor stays zero (cv::normalize's epsilon guard)


# ---------------------------------------------------------------------------------------------------------------
# Differential pinning beyond the committed goldens: fresh random inputs every seed, the restatement against the
# UNMODIFIED reference compiled in place (oracle/_ref travels with the repo; skipped where it was never built).
@pytest.mark.skipif(not orc.have_ref("ref_flat_bf_sse"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flat_restatement_vs_reference_random(seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    for d in (16, 100, 128):
        n, nq, k = int(rng.integers(200, 700)), 7, int(rng.integers(1, 40))
        coarse_grid = seed == 3  # few distinct values: masses of exact ties, the (dist, label) rule decides
        x = (rng.integers(0, 4, (n, d)).astype(np.float32) * np.float32(0.25)) if coarse_grid else rng.standard_normal((n, d), dtype=np.float32)
        q = (rng.integers(0, 4, (nq, d)).astype(np.float32) * np.float32(0.25)) if coarse_grid else rng.standard_normal((nq, d), dtype=np.float32)
        labels = rng.permutation(n).astype(np.uint64) * np.uint64(3) + np.uint64(11)
        xu = rng.integers(0, 4 if coarse_grid else 256, (n, d)).astype(np.uint8)
        qu = rng.integers(0, 4 if coarse_grid else 256, (nq, d)).astype(np.uint8)
        runs = [("bf_sse", "ip", 0, 4), ("hnsw", "ip", 0, 4), ("hnsw", "l2i", 2, 0)]
        runs += [("bf_avx", "ip", 0, 8), ("hnsw", "l2", 1, 8)] if d % 16 == 0 else [("hnsw", "l2", 1, 4)]
        for flav, metric, mcode, lanes in runs:
            data, qq = (xu, qu) if metric == "l2i" else (x, q)
            dist, lab = orc.run_ref_flat(flav, metric, data, labels, qq, k)
            od, ol = orc.flat_search(mcode, lanes, data, labels, qq, k)
            assert np.array_equal(lab, ol), (seed, d, flav, metric)
            assert np.array_equal(dist.view(np.uint32), od.view(np.uint32)), (seed, d, flav, metric)


@pytest.mark.skipif(not orc.have_ref("ref_opq"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,D,M,K,nk", [(1, 64, 8, 12, 3), (2, 128, 16, 1, 1), (3, 32, 4, 40, 2)])
def test_opq_restatement_vs_reference_random(seed, D, M, K, nk, tmp_path):
    """IVFOPQ::Add + QueryThrehold + get_sort_results on a fresh random model / database / queries."""
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    n, nq, rows_per_group, topk = 700, 9, 7, 5
    db = synth.sift_like(n, D, seed=3000 + seed)
    q = synth.sift_like(nq, D, seed=4000 + seed)
    reorder = rng.permutation(D).astype(np.int32)
    coarse = db[rng.choice(n, K, replace=False)][:, reorder].copy() if K > 1 else (rng.standard_normal((1, D)) * 0.02).astype(np.float32)
    cb = (rng.standard_normal((M, 256, D // M)) * 0.08).astype(np.float32)
    model = str(tmp_path / "m.model")
    synth.write_opq_model(model, coarse, cb, reorder)
    dbf = []
    for gi, lo in enumerate(range(0, n, rows_per_group)):
        p = str(tmp_path / f"db{gi}.bin")
        synth.write_feat_file(p, db[lo:lo + rows_per_group])
        dbf.append(p)
    qp = str(tmp_path / "q.bin")
    synth.write_feat_file(qp, q)
    ref = orc.run_ref_opq(model, dbf, [qp], nk=nk, topk=topk, per_row=False)
    x = orc.opq_reorder(db, reorder)
    lists = orc.opq_coarse_assign(x, coarse)
    assert np.array_equal(lists, ref["row_list"])
    codes = orc.opq_pq_encode(x, coarse, lists, cb)
    assert np.array_equal(codes, ref["codes"])
    match = orc.opq_query_scores(orc.opq_reorder(q, reorder), coarse, cb, nk, lists, ref["row_group"], codes, ref["n_groups"], 1.0)
    assert np.array_equal(_bits(match), _bits(ref["match"]))
    for f in range(nq):
        s, i = orc.topk_pairs(match[f], topk)
        assert np.array_equal(i, ref["topk_id"][f]) and np.array_equal(_bits(s), _bits(ref["topk_score"][f]))


@pytest.mark.skipif(not orc.have_ref("ref_int8_quan"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_sq_restatement_vs_reference_random(seed):
    """Int8Quan::L2NormalizeVector / Int8Encode / Int8Decode(std::string&) of the unmodified int8_quan.cc (compiled against
    the faiss / nlohmann stand-ins) on fresh random rows and trained ranges: the restatement must equal it bit for bit.
    Seeds 3, 4 go through the multi-model constructor (JSON conf, source 1 / 2)."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    d = int(rng.choice([32, 64, 100, 128, 256]))
    n = int(rng.integers(40, 160))
    x = (rng.standard_normal((n, d), dtype=np.float32) * np.float32(rng.uniform(0.2, 3.0))).astype(np.float32)
    if seed % 2 == 0:
        x = np.maximum(x, 0)                       # ReLU-sparse rows, as the reference's own test vector
    x[0] = 0                                       # max(1e-12, norm) guard
    x[1] *= np.float32(1e-20)                      # denormal-range squares: accum is double, the quotient stays finite
    x[2] *= np.float32(1e6)
    xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    sub = xn[rng.choice(n, n // 2, replace=False)]  # range trained on a subset: some values fall outside (clamp to 0 / 255)
    vmin = sub.min(0).astype(np.float32)
    vdiff = (sub.max(0) - vmin).astype(np.float32)
    vdiff[int(rng.integers(d))] = 0                # constant dimension
    for l2norm in (True, False):
        ref = orc.run_ref_int8_quan(x, vmin, vdiff, l2norm=l2norm, via_conf=seed >= 3, source=seed - 2 if seed >= 3 else 0)
        codes, xa = orc.sq_encode(x, vmin, vdiff, l2norm=l2norm)
        assert np.array_equal(codes, ref["codes"]), (seed, l2norm)
        assert np.array_equal(_bits(xa), _bits(ref["x_after"])), (seed, l2norm)
        nrm = x.copy()
        for i in range(n):
            orc.lib().orc_sq_l2normalize(orc._p(nrm[i]), d)
        assert np.array_equal(_bits(nrm), _bits(ref["normed"]))
        assert np.array_equal(_bits(orc.sq_decode(codes, vmin, vdiff)), _bits(ref["decode"])), (seed, l2norm)
        assert ref["rc_bad_dims"] == 0             # n_dims % code_size != 0 -> 0 (int8_quan.cc:73-75)
