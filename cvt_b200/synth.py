"""Seeded synthetic inputs for tests and bench.

Everything here is plain numpy on the host: it produces the *inputs* (vectors, codebooks, model
files in the reference's byte formats) that both the CUDA path and the CPU checker consume.  No
search/encode arithmetic lives here.

Formats follow the reference:
  * OPQ model file   -- IVFOPQ::LoadModel, opq/src/IVFOPQ.cpp:75-95
  * raw feature file -- IVFOPQ::LoadSingleFeatFile, opq/src/IVFOPQ.cpp:441-458 (App. A-2)
"""
from __future__ import annotations

import numpy as np

SEED_DB = 0x5EED0001
SEED_QUERY = 0x5EED0002
SEED_KMEANS = 0x5EED0003
SEED_DENSE_R = 0x5EED0004

# The permutation shipped at the tail of opq/model/*.model: a valid
# permutation of 0..127 used as the "rotation" for D=128 synthetic models.
SHIPPED_REORDER_128 = np.array(
    [0, 31, 47, 63, 79, 95, 111, 127, 1, 30, 46, 62, 78, 94, 110, 119, 2, 29, 44, 61, 77, 80, 107, 113,
     3, 28, 45, 60, 74, 86, 103, 116, 4, 27, 43, 57, 66, 91, 99, 123, 5, 26, 42, 55, 68, 88, 100, 122,
     6, 25, 41, 53, 70, 89, 98, 126, 7, 24, 39, 54, 69, 92, 97, 124, 8, 23, 40, 50, 73, 83, 106, 118,
     9, 22, 38, 48, 76, 81, 108, 117, 10, 21, 37, 49, 75, 82, 109, 112, 11, 20, 36, 52, 71, 85, 104, 115,
     12, 19, 34, 59, 64, 93, 96, 125, 13, 18, 33, 58, 65, 90, 101, 121, 14, 17, 32, 56, 67, 87, 102, 120,
     15, 16, 35, 51, 72, 84, 105, 114], dtype=np.int32)


def sift_like(n: int, d: int = 128, seed: int = SEED_DB, noise: float = 40.0, chunk: int = 1 << 18) -> np.ndarray:
    """"SIFT-shaped" unit-norm fp32 vectors: clipped non-negative integer histograms around 1024
    cluster centres, then rootSIFT (abs -> /L1 -> sqrt) and L2 normalisation, the reference's own
    descriptor normalisation (hnsw_sifts_retrieval/siftsIndex.cpp:54-71)."""
    rng_c = np.random.Generator(np.random.PCG64(0xC0FFEE ^ d))
    centres = rng_c.random((1024, d), dtype=np.float32)
    out = np.empty((n, d), dtype=np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        g = np.abs(rng.standard_normal((hi - lo, d), dtype=np.float32)) * np.float32(noise)
        g += np.float32(30.0) * centres[np.arange(lo, hi) % 1024]
        v = np.minimum(np.floor(g), np.float32(255.0))
        v = v / (np.abs(v).sum(axis=1, keepdims=True) + np.float32(1e-7))
        v = np.sqrt(v)
        v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), np.float32(1e-12))
        out[lo:hi] = v.astype(np.float32)
    return out


def cnn_like(n: int, d: int = 512, seed: int = SEED_DB) -> np.ndarray:
    """ReLU-sparse "CNN fc feature" vectors, L2-normalised (utils/math_util.h:29-39 formula)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = np.maximum(rng.standard_normal((n, d), dtype=np.float32), 0)
    v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), np.float32(1e-12))
    return v.astype(np.float32)


def random_permutation(d: int, seed: int = SEED_DENSE_R) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(seed)).permutation(d).astype(np.int32)


def dense_rotation(d: int, seed: int = SEED_DENSE_R) -> np.ndarray:
    """Orthonormal R (row-major [d,d]) from the QR of a seeded Gaussian; y = R @ x."""
    g = np.random.Generator(np.random.PCG64(seed)).standard_normal((d, d))
    q, r = np.linalg.qr(g)
    q = q * np.sign(np.diag(r))
    return np.ascontiguousarray(q.astype(np.float32))


def _kmeans(x: np.ndarray, k: int, iters: int, rng: np.random.Generator) -> np.ndarray:
    """Plain Lloyd k-means (fp32 distances via the expanded form; training is unpinned in the
    reference -- yael with random init -- so only the *output* matters)."""
    n = x.shape[0]
    c = x[rng.choice(n, size=k, replace=n < k)].copy()
    for _ in range(iters):
        d2 = (x * x).sum(1, keepdims=True) - 2.0 * (x @ c.T) + (c * c).sum(1)[None, :]
        a = d2.argmin(1)
        for j in range(k):
            m = a == j
            if m.any():
                c[j] = x[m].mean(0)
            else:
                c[j] = x[rng.integers(n)]
    return c.astype(np.float32)


def train_pq_model(x_rot: np.ndarray, M: int, ksub: int = 256, K: int = 1, iters: int = 10,
                   seed: int = SEED_KMEANS, train_rows: int = 20000):
    """Seeded codebooks for synthetic models.  x_rot must already be in the rotated/permuted
    space (training happens there, opq/train_codebook/train_PQ_codebook.cpp:80,98,112).
    Returns (coarse [K,D], codebooks [M,ksub,D/M]).  K=1 uses the zero centroid (flat ADC)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.ascontiguousarray(x_rot[:train_rows], dtype=np.float32)
    n, D = x.shape
    if K == 1:
        coarse = np.zeros((1, D), dtype=np.float32)
        res = x
    else:
        coarse = _kmeans(x, K, iters, rng)
        d2 = (x * x).sum(1, keepdims=True) - 2.0 * (x @ coarse.T) + (coarse * coarse).sum(1)[None, :]
        res = x - coarse[d2.argmin(1)]
    ds = D // M
    cb = np.empty((M, ksub, ds), dtype=np.float32)
    for m in range(M):
        cb[m] = _kmeans(np.ascontiguousarray(res[:, m * ds:(m + 1) * ds]), ksub, iters, rng)
    return coarse, cb


def _kmeans_fast(x: np.ndarray, k: int, iters: int, rng: np.random.Generator) -> np.ndarray:
    """Lloyd k-means with a vectorised update (sort by assignment + segmented sums in double): the SURVEY.md 8(d)
    protocol (25 iterations, 256 centroids per sub-space, 100 k rows) in seconds.  Used for the BENCH models; the test
    models keep _kmeans (their golden files pin its output)."""
    n, d = x.shape
    c = x[rng.choice(n, size=k, replace=n < k)].copy()
    for _ in range(iters):
        g = x @ (np.float32(-2.0) * c).T          # |x|^2 is constant per row: argmin of |c|^2 - 2<x,c>
        g += (c * c).sum(1)[None, :]
        a = g.argmin(1)
        cnt = np.bincount(a, minlength=k)
        nz = cnt > 0
        order = np.argsort(a, kind="stable")
        starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        sums = np.add.reduceat(x[order].astype(np.float64), starts[nz], axis=0)
        c[nz] = (sums / cnt[nz, None]).astype(np.float32)
        if (~nz).any():
            c[~nz] = x[rng.integers(n, size=int((~nz).sum()))]
    return c.astype(np.float32)


def train_pq_model_bench(x_rot: np.ndarray, M: int, ksub: int = 256, iters: int = 25, seed: int = SEED_KMEANS):
    """Flat-ADC (K = 1, zero coarse centroid) codebooks for the benchmark workloads: seeded Lloyd, `iters` iterations over
    all rows given (the first 100 k rows of the database, already in the permuted space).  Sub-spaces are
    independent (own generator, seed + m) and trained on a thread pool."""
    import concurrent.futures as cf
    import os
    x = np.ascontiguousarray(x_rot, dtype=np.float32)
    D = x.shape[1]
    ds = D // M

    def one(m):
        return _kmeans_fast(np.ascontiguousarray(x[:, m * ds:(m + 1) * ds]), ksub, iters, np.random.Generator(np.random.PCG64(seed + 7919 * m)))
    threads = max(1, min(M, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        cb = np.stack(list(ex.map(one, range(M)))).astype(np.float32)
    return np.zeros((1, D), dtype=np.float32), cb


def write_opq_model(path: str, coarse: np.ndarray, cb: np.ndarray, reorder: np.ndarray) -> None:
    """int32 D,K,M,ksub | f32 coarse[K][D] | f32 cb[M][ksub][D/M] | int32 reorder[D]."""
    K, D = coarse.shape
    M, ksub, ds = cb.shape
    assert M * ds == D and reorder.shape == (D,)
    with open(path, "wb") as f:
        np.array([D, K, M, ksub], dtype="<i4").tofile(f)
        np.ascontiguousarray(coarse, dtype="<f4").tofile(f)
        np.ascontiguousarray(cb, dtype="<f4").tofile(f)
        np.ascontiguousarray(reorder, dtype="<i4").tofile(f)


def read_opq_model(path: str):
    with open(path, "rb") as f:
        D, K, M, ksub = np.fromfile(f, dtype="<i4", count=4)
        coarse = np.fromfile(f, dtype="<f4", count=K * D).reshape(K, D)
        cb = np.fromfile(f, dtype="<f4", count=M * ksub * (D // M)).reshape(M, ksub, D // M)
        reorder = np.fromfile(f, dtype="<i4", count=D)
    if sorted(reorder.tolist()) != list(range(D)):
        raise ValueError("model tail is not a permutation of 0..D-1")
    return coarse, cb, reorder


def write_feat_file(path: str, x: np.ndarray) -> None:
    np.ascontiguousarray(x, dtype="<f4").tofile(path)


def sq_minmax(x_normed: np.ndarray):
    """faiss RS_minmax, rs_arg=0 (scalar_quantization/train/src/sq_train.cpp:100-132)."""
    vmin = x_normed.min(0).astype(np.float32)
    vdiff = (x_normed.max(0) - vmin).astype(np.float32)
    return vmin, vdiff
