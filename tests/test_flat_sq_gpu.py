"""GPU parity tests of the exact flat index (BruteforceSearch drop-in) and of the scalar quantizer,
through the C ABI, against golden vectors from the unmodified reference (flat) and against the
oracle restatement (flat at larger sizes; SQ, whose parity is unpinned at the faiss boundary)."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
G = cases.GOLDEN


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def ctx():
    from cvt_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


RUNS = {"ip_sse": ("ip", 4), "ip_hnsw": ("ip", 4), "ip_avx": ("ip", 8), "l2_avx": ("l2", 8), "l2_sse": ("l2", 4),
        "l2i": ("l2_u8", 4)}


@pytest.mark.parametrize("name", list(cases.FLAT_CASES))
def test_flat_vs_reference_golden(ctx, name):
    from cvt_b200 import capi
    c = cases.flat_case(name)
    gold = np.load(os.path.join(G, f"flat_{name}.npz"))
    assert str(gold["input_sha"]) == c["input_sha"]
    seen = 0
    for tag, (metric, order) in RUNS.items():
        if tag + "_dist" not in gold:
            continue
        seen += 1
        data, q = (c["xu"], c["qu"]) if metric == "l2_u8" else (c["x"], c["q"])
        idx = capi.FlatIndex(ctx, metric, c["d"], c["n"] + 5, order=order)
        # two batches: addPoint order is preserved, labels are arbitrary (permuted, non-contiguous)
        h = c["n"] // 3
        idx.add(data[:h], c["labels"][:h])
        idx.add(data[h:], c["labels"][h:])
        assert len(idx) == c["n"]
        D, L = idx.search(q, c["k"])
        assert np.array_equal(L, gold[tag + "_label"]), (name, tag)
        assert np.array_equal(_bits(D), _bits(gold[tag + "_dist"])), (name, tag)
        idx.close()
    assert seen >= 4


def test_flat_scalar_order_and_odd_dim(ctx):
    # d % 4 != 0 -> the reference falls back to the scalar InnerProduct / L2Sqr loops (order 1)
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(9))
    n, d, nq, k = 700, 37, 9, 12
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    labels = rng.permutation(n).astype(np.uint64)
    for metric, m in (("ip", 0), ("l2", 1)):
        idx = capi.FlatIndex(ctx, metric, d, n, order=1)
        idx.add(x, labels)
        D, L = idx.search(q, k)
        od, ol = orc.flat_search(m, 1, x, labels, q, k)
        assert np.array_equal(L, ol) and np.array_equal(_bits(D), _bits(od))
        idx.close()


def test_flat_larger_vs_oracle_and_errors(ctx, tmp_path):
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(10))
    n, d, nq, k = 50_000, 128, 70, 100
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = x[rng.integers(0, n, nq)] + 0.05 * rng.standard_normal((nq, d), dtype=np.float32)
    labels = np.arange(n, dtype=np.uint64)
    idx = capi.FlatIndex(ctx, "ip", d, n, order=4)
    idx.add(x, labels)
    D, L = idx.search(q, k)
    for i in range(0, nq, 7):
        od, ol = orc.flat_search(0, 4, x, labels, q[i:i + 1], k)
        assert np.array_equal(L[i], ol[0]) and np.array_equal(_bits(D[i]), _bits(od[0]))
    # duplicate label / capacity -> the reference throws std::runtime_error (brutoforce.hpp:44-50)
    with pytest.raises(capi.B200nnError, match="Ids have to be unique"):
        idx.add(x[:1], labels[:1])
    with pytest.raises(capi.B200nnError, match="exceeds the specified limit"):
        idx.add(x[:1], np.array([n + 10], dtype=np.uint64))
    # removePoint: last row moves into the hole (brutoforce.hpp:58-70)
    idx.remove(5)
    assert len(idx) == n - 1
    D2, L2 = idx.search(q[:4], k)
    keep = labels != 5
    od, ol = orc.flat_search(0, 4, x[keep], labels[keep], q[:4], k)
    assert np.array_equal(L2, ol) and np.array_equal(_bits(D2), _bits(od))
    idx.close()
    # u8 at size
    xu = rng.integers(0, 256, size=(20_000, 128)).astype(np.uint8)
    qu = rng.integers(0, 256, size=(33, 128)).astype(np.uint8)
    lu = rng.permutation(20_000).astype(np.uint64)
    idx = capi.FlatIndex(ctx, "l2_u8", 128, 20_000)
    idx.add(xu, lu)
    D, L = idx.search(qu, 10)
    od, ol = orc.flat_search(2, 0, xu, lu, qu, 10)
    assert np.array_equal(L, ol) and np.array_equal(D, od)
    idx.close()


def test_flat_save_load_byte_format(ctx, tmp_path):
    from cvt_b200 import capi
    c = cases.flat_case("ties_d16")
    idx = capi.FlatIndex(ctx, "ip", 16, 40, order=4)
    idx.add(c["x"][:40], c["labels"][:40])
    p = str(tmp_path / "index.bin")
    idx.save(p)
    ref = open(os.path.join(G, "flat_saveindex_ties_d16_n40.bin"), "rb").read()  # written by the reference's saveIndex
    assert open(p, "rb").read() == ref
    idx2 = capi.FlatIndex.load_file(ctx, "ip", 16, os.path.join(G, "flat_saveindex_ties_d16_n40.bin"), order=4)
    assert len(idx2) == 40
    D1, L1 = idx.search(c["q"], 5)
    D2, L2 = idx2.search(c["q"], 5)
    assert np.array_equal(L1, L2) and np.array_equal(_bits(D1), _bits(D2))
    # fewer rows than k: tail is (+inf, UINT64_MAX) instead of the reference's uninitialised reads
    D3, L3 = idx.search(c["q"][:2], 50)
    assert np.all(np.isinf(D3[:, 40:])) and np.all(L3[:, 40:] == np.uint64(0xFFFFFFFFFFFFFFFF))
    idx.close(); idx2.close()


@pytest.mark.parametrize("d", [64, 128])
def test_sq_vs_reference_golden(ctx, d):
    """codes / x_normed / decode of the golden are REFERENCE-RUN (the unmodified int8_quan.cc compiled against stand-ins for
    its un-vendored dependencies, oracle/gen_golden.py::gen_sq): Int8Encode, L2NormalizeVector and Int8Decode(std::string&)
    on the GPU must equal them bit for bit.  decode_faiss (faiss's own all-float decode) is the one unpinned output."""
    from cvt_b200 import capi
    c = cases.sq_case(d)
    gold = np.load(os.path.join(G, f"sq_d{d}.npz"))
    assert bool(gold["pinned"]) and str(gold["input_sha"]) == c["input_sha"]
    sq = capi.SQ(ctx, c["vmin"], c["vdiff"])
    codes, xn = sq.encode(c["x"], l2norm=True)
    assert np.array_equal(codes, gold["codes"])
    assert np.array_equal(_bits(xn), _bits(gold["x_normed"]))  # caller's buffer is normalised in place
    codes_nn, x_same = sq.encode(c["x"], l2norm=False)
    assert np.array_equal(codes_nn, gold["codes_nonorm"]) and np.array_equal(_bits(x_same), _bits(c["x"]))
    assert np.array_equal(_bits(sq.decode(codes)), _bits(gold["decode"]))
    assert np.array_equal(_bits(sq.decode(codes, faiss_float=True)), _bits(gold["decode_faiss"]))
    vmin, vdiff = capi.SQ.train_minmax(ctx, xn)
    ov, od = orc.sq_train_minmax(xn)
    assert np.array_equal(_bits(vmin), _bits(ov)) and np.array_equal(_bits(vdiff), _bits(od))
    sq.close()


def test_sq_roundtrip_property_at_size(ctx):
    # encode -> decode stays within half a bucket inside the trained range; int8 scan of the codes
    from cvt_b200 import capi, synth
    x = synth.sift_like(100_000, 128, seed=31)
    vmin, vdiff = capi.SQ.train_minmax(ctx, x[:50_000])
    sq = capi.SQ(ctx, vmin, vdiff)
    codes, xn = sq.encode(x, l2norm=True)
    dec = sq.decode(codes)
    inside = (xn >= vmin) & (xn <= vmin + vdiff)
    err = np.abs(dec - xn)
    assert np.all(err[inside] <= (np.broadcast_to(vdiff / 255.0, xn.shape)[inside] + 1e-6))
    oc, _ = orc.sq_encode(x[:500], vmin, vdiff, True)
    assert np.array_equal(codes[:500], oc)
    # cfg2 shape in miniature: exact int32 L2 scan over the SQ codes
    idx = capi.FlatIndex(ctx, "l2_u8", 128, len(codes))
    idx.add(codes, np.arange(len(codes), dtype=np.uint64))
    D, L = idx.search(codes[:16], 10)
    assert np.array_equal(L[:, 0], np.arange(16, dtype=np.uint64)) and np.all(D[:, 0] == 0)
    od, ol = orc.flat_search(2, 0, codes, np.arange(len(codes), dtype=np.uint64), codes[:2], 10)
    assert np.array_equal(L[:2], ol) and np.array_equal(D[:2], od)
    sq.close(); idx.close()


@pytest.mark.parametrize("D,n,nq,k,hi", [(128, 33_333, 130, 10, 256), (32, 1000, 5, 1, 256), (64, 5000, 300, 32, 3), (256, 2049, 129, 7, 256),
                                          (128, 200, 17, 32, 2), (256, 3000, 40, 32, 256),
                                          # >= 262144 rows: the sample pass seeds the thresholds; tie-heavy so that rows AT the bound matter
                                          (32, 300_001, 70, 10, 3), (64, 270_000, 33, 32, 256)])
def test_u8_tensor_core_scan_vs_oracle_and_dp4a(ctx, D, n, nq, k, hi):
    """cfg2 path: tcgen05 kind::i8 GEMM + fused top-k == oracle (L2SqrI + (dist,label) heap rule) == dp4a kernel,
    on ragged sizes and tie-heavy data (hi=2,3: tiny alphabets -> masses of equal distances)."""
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(D * 7 + n))
    xu = rng.integers(0, hi, size=(n, D)).astype(np.uint8)
    qu = rng.integers(0, hi, size=(nq, D)).astype(np.uint8)
    labels = (rng.permutation(n).astype(np.uint64) * np.uint64(3)) + np.uint64(11)
    idx = capi.FlatIndex(ctx, "l2_u8", D, n)
    idx.add(xu[: n // 2], labels[: n // 2])
    idx.add(xu[n // 2:], labels[n // 2:])
    os.environ.pop("B200NN_NO_TC_U8", None)
    Dt, Lt = idx.search(qu, k)
    os.environ["B200NN_NO_TC_U8"] = "1"
    try:
        Dd, Ld = idx.search(qu, k)
    finally:
        os.environ.pop("B200NN_NO_TC_U8", None)
    assert np.array_equal(Lt, Ld) and np.array_equal(Dt, Dd)
    if n >= 262_144:  # the one-pass shared-bound scan (default when k <= 32) == the three passes with running thresholds
        os.environ["B200NN_U8_NO_SHARED_BOUND"] = "1"
        try:
            D3, L3 = idx.search(qu, k)
        finally:
            os.environ.pop("B200NN_U8_NO_SHARED_BOUND", None)
        assert np.array_equal(Lt, L3) and np.array_equal(Dt, D3)
    sel = range(nq) if n <= 5000 else range(0, nq, 9)
    for i in sel:
        od, ol = orc.flat_search(2, 0, xu, labels, qu[i:i + 1], k)
        assert np.array_equal(Lt[i], ol[0]) and np.array_equal(Dt[i], od[0]), (D, n, i)
    idx.close()
