"""Seeded input generators shared by oracle/gen_golden.py (which runs the compiled reference on
them and commits the outputs under tests/golden/) and by the parity tests (which regenerate the
same inputs and compare the CUDA path / the oracle with those committed outputs).

Each case stores a sha256 of its inputs in the golden file so generator drift is detected."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cvt_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURE_DB = ["6231519245_feat.bin", "6231075428_feat.bin", "6230951284_feat.bin", "6230880830_feat.bin",
              "6231307582_feat.bin"]  # order of opq/data/5_feats_list.txt
FIXTURE_QUERY = ["6231519245_feat.bin", "6231519245_6_feat.bin"]


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------- opq synthetic
OPQ_CASES = {
    # name: D, M, K, nk, n, rows_per_group, nq, topk, query_scale
    "flat_m16": dict(D=128, M=16, K=1, nk=1, n=3000, rows_per_group=1, nq=24, topk=100, qscale=1.0),
    "flat_m16_clamp": dict(D=128, M=16, K=1, nk=1, n=1500, rows_per_group=1, nq=8, topk=50, qscale=2.5),
    "ivf_m8": dict(D=64, M=8, K=16, nk=3, n=2000, rows_per_group=50, nq=12, topk=5, qscale=1.0),
    "flat_m8_d128": dict(D=128, M=8, K=1, nk=1, n=1000, rows_per_group=1, nq=8, topk=20, qscale=1.0),
}


def opq_case(name: str):
    c = dict(OPQ_CASES[name])
    D, M, K = c["D"], c["M"], c["K"]
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    db = synth.sift_like(c["n"], D, seed=seed)
    q = synth.sift_like(c["nq"], D, seed=seed + 1) * np.float32(c["qscale"])
    reorder = synth.SHIPPED_REORDER_128 if D == 128 else synth.random_permutation(D, seed=seed + 2)
    train = db[:, reorder]
    coarse, cb = synth.train_pq_model(train, M, 256, K, iters=4, seed=seed + 3, train_rows=1500)
    if K == 1:
        # a non-trivial single centroid exercises the residual path even in the flat configs
        coarse = (train[:64].mean(0, keepdims=True) * np.float32(0.25)).astype(np.float32)
    c.update(db=db, q=q.astype(np.float32), reorder=reorder.astype(np.int32), coarse=coarse, cb=cb,
             input_sha=sha(db, q, reorder, coarse, cb))
    return c


# ------------------------------------------------------------------------------- flat
FLAT_CASES = {
    # name: n, d, nq, k, kind
    "unit_d128": dict(n=2000, d=128, nq=16, k=10, kind="unit"),
    "ties_d16": dict(n=600, d=16, nq=12, k=25, kind="ties"),
    "unit_d100": dict(n=500, d=100, nq=6, k=5, kind="unit"),  # d%4==0, d%16!=0 -> SIMD4Ext
}


def flat_case(name: str):
    c = dict(FLAT_CASES[name])
    seed = int(hashlib.sha256(("flat" + name).encode()).hexdigest()[:8], 16)
    rng = np.random.Generator(np.random.PCG64(seed))
    n, d, nq = c["n"], c["d"], c["nq"]
    if c["kind"] == "unit":
        x = rng.standard_normal((n, d), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        q = rng.standard_normal((nq, d), dtype=np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    else:  # tiny alphabet -> masses of exactly equal distances, (dist,label) tie rule decides
        x = rng.integers(0, 3, size=(n, d)).astype(np.float32) * np.float32(0.25)
        q = rng.integers(0, 3, size=(nq, d)).astype(np.float32) * np.float32(0.25)
    labels = rng.permutation(n).astype(np.uint64) * np.uint64(7) + np.uint64(3)
    xu = rng.integers(0, 256 if c["kind"] == "unit" else 3, size=(n, d)).astype(np.uint8)
    qu = rng.integers(0, 256 if c["kind"] == "unit" else 3, size=(nq, d)).astype(np.uint8)
    c.update(x=x.astype(np.float32), q=q.astype(np.float32), labels=labels, xu=xu, qu=qu,
             input_sha=sha(x, q, labels, xu, qu))
    return c


# ------------------------------------------------------------------------------- sq
# the hard-coded vector of scalar_quantization/scalar_quantization/int8_quan_test.cpp:26
SQ_REF_TEST_VECTOR = np.array(
    [0.7678224, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.6331244, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.583638,
     0.76271933, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.21529453, 0.0, 0.0, 1.2015152, 0.0, 0.0, 0.0, 0.0, 0.0,
     0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.88310665, 0.0, 0.0, 0.19277531, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.5779805,
     0.0, 0.0, 0.7728174, 0.0, 2.21898, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], dtype=np.float32)


def sq_case(d: int = 64, n: int = 400):
    rng = np.random.Generator(np.random.PCG64(0x5C0DE + d))
    x = np.maximum(rng.standard_normal((n, d), dtype=np.float32) * np.float32(1.3), 0).astype(np.float32)
    if d == 64:
        x[0] = SQ_REF_TEST_VECTOR
    x[1] = 0  # all-zero row: max(1e-12, norm) guard
    train = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    vmin = train.min(0).astype(np.float32)
    vdiff = (train.max(0) - vmin).astype(np.float32)
    vdiff[3] = 0  # a constant dimension: the vdiff == 0 branch
    return dict(d=d, n=n, x=x, vmin=vmin, vdiff=vdiff, input_sha=sha(x, vmin, vdiff))


# ------------------------------------------------------------------------------- CLI record files
def write_record_file(path: str, ids, feats: np.ndarray) -> None:
    """brute_force.cpp's db.bin / querys.bin format:
    int32 num; num x { int32 idLen; char id[idLen]; int32 dim; float32 feat[dim] }."""
    with open(path, "wb") as f:
        np.array([len(ids)], dtype="<i4").tofile(f)
        for i, s in enumerate(ids):
            b = s.encode()
            np.array([len(b)], dtype="<i4").tofile(f)
            f.write(b)
            np.array([feats.shape[1]], dtype="<i4").tofile(f)
            np.ascontiguousarray(feats[i], dtype="<f4").tofile(f)


def cli_case():
    c = flat_case("unit_d128")
    db_ids = [f"img_{i:06d}" for i in range(c["n"])]
    q_ids = [f"query_{i:03d}" for i in range(c["nq"])]
    return dict(db=c["x"], q=c["q"], db_ids=db_ids, q_ids=q_ids)


# ------------------------------------------------------------------------------- front end (f-3)
def frontend_pca_inputs(K: int = 1024, n: int = 96):
    """ReLU-sparse "CNN pool feature" rows (what pca_train_project reduces), plus edge rows: the all-zero row and a row
    equal to nothing-but-noise at 1e-3 scale (small norms exercise the max(1e-12, norm) guard's neighbourhood)."""
    rng = np.random.Generator(np.random.PCG64(0xF3000 + K))
    x = np.maximum(rng.standard_normal((n, K), dtype=np.float32) * np.float32(0.8) + np.float32(0.2), 0).astype(np.float32)
    x[1] = 0
    x[2] *= np.float32(1e-3)
    return x


def frontend_sift_inputs(n: int = 80, d: int = 128):
    """SIFT-like integer-valued descriptors 0..255 (some negated: rootSift takes abs first), one all-zero row."""
    rng = np.random.Generator(np.random.PCG64(0x51F7))
    v = np.minimum(np.floor(np.abs(rng.standard_normal((n, d), dtype=np.float32)) * np.float32(45.0)), np.float32(255.0))
    v[::7] *= np.float32(-1.0)
    v[3] = 0
    return v.astype(np.float32)


# ------------------------------------------------------------------------------- training (f-4)
def train_inputs(n: int, d: int, seed: int = 0, centres: int = 40):
    """Clustered rows (Gaussian blobs around seeded centres) for the k-means tests."""
    rng = np.random.Generator(np.random.PCG64(0x7A11 + seed))
    c = rng.standard_normal((centres, d), dtype=np.float32) * np.float32(2.0)
    x = c[rng.integers(0, centres, n)] + rng.standard_normal((n, d), dtype=np.float32) * np.float32(0.5)
    return np.ascontiguousarray(x, dtype=np.float32)


# ------------------------------------------------------------------------------- scalar quantizer fixture
def int8_quan_test_vector():
    """The 64-d input pinned by the reference's own test (scalar_quantization/scalar_quantization/int8_quan_test.cpp:26)."""
    return SQ_REF_TEST_VECTOR.copy()


def int8_quan_test_model():
    """The seeded trained range the int8_quan_test mains (reference and drop-in) load as their model."""
    rng = np.random.Generator(np.random.PCG64(64))
    vmin = (-rng.random(64) * 0.02).astype(np.float32)
    vdiff = (rng.random(64) * 0.6 + 0.05).astype(np.float32)
    vdiff[5] = 0.0                                        # a constant dimension: code 0 (int8_quan.cc:82)
    return vmin, vdiff


def write_ixsq_file(path: str, vmin, vdiff) -> None:
    """faiss 1.5.3 IndexScalarQuantizer file as described in SURVEY.md App. A-8 (QT_8bit, RS_minmax, no stored codes)."""
    import struct
    d = len(vmin)
    with open(path, "wb") as f:
        f.write(b"IxSQ")
        f.write(struct.pack("<iqqqBi", d, 0, 1 << 20, 1 << 20, 1, 1))           # index header
        f.write(struct.pack("<iifQQ", 0, 0, 0.0, d, d))                          # qtype, rangestat, arg, d, code_size
        f.write(struct.pack("<Q", 2 * d))
        np.concatenate([np.asarray(vmin, np.float32), np.asarray(vdiff, np.float32)]).astype("<f4").tofile(f)
        f.write(struct.pack("<Q", 0))
