"""ctypes front-end of the CPU checker (oracle/cvt_oracle.c) and runner of the compiled reference
(oracle/_ref/*).  TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

_lib = None


def build(verbose: bool = False) -> None:
    """Compile the C restatement (always) and, when /root/reference is present, oracle/_ref."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_flat_ip.restype = C.c_float
        _lib.orc_flat_l2.restype = C.c_float
        _lib.orc_flat_l2_u8.restype = C.c_int32
    return _lib


def have_ref(name: str = "ref_opq") -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------- opq
def opq_reorder(x, perm):
    x = _f32(x)
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    y = np.empty_like(x)
    lib().orc_opq_reorder(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(perm), _p(y))
    return y


def opq_rotate_dense(x, R):
    x, R = _f32(x), _f32(R)
    y = np.empty_like(x)
    lib().orc_opq_rotate_dense(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(R), _p(y))
    return y


def opq_coarse_assign(x_rot, coarse):
    x_rot, coarse = _f32(x_rot), _f32(coarse)
    out = np.empty(x_rot.shape[0], dtype=np.int32)
    lib().orc_opq_coarse_assign(_p(x_rot), C.c_int64(x_rot.shape[0]), C.c_int(x_rot.shape[1]), _p(coarse),
                                C.c_int(coarse.shape[0]), _p(out))
    return out


def opq_pq_encode(x_rot, coarse, lists, cb):
    x_rot, coarse, cb = _f32(x_rot), _f32(coarse), _f32(cb)
    lists = np.ascontiguousarray(lists, dtype=np.int32)
    M, ksub, _ = cb.shape
    codes = np.empty((x_rot.shape[0], M), dtype=np.uint8)
    lib().orc_opq_pq_encode(_p(x_rot), C.c_int64(x_rot.shape[0]), C.c_int(x_rot.shape[1]), _p(coarse), _p(lists),
                            _p(cb), C.c_int(M), C.c_int(ksub), _p(codes))
    return codes


def opq_coarse_probe(q_rot_row, coarse, nk):
    q, coarse = _f32(q_rot_row), _f32(coarse)
    out = np.empty(nk, dtype=np.int32)
    lib().orc_opq_coarse_probe(_p(q), C.c_int(coarse.shape[1]), _p(coarse), C.c_int(coarse.shape[0]), C.c_int(nk),
                               _p(out))
    return out


def opq_build_lut(q_rot_row, centroid, cb):
    q, centroid, cb = _f32(q_rot_row), _f32(centroid), _f32(cb)
    M, ksub, _ = cb.shape
    lut = np.empty((M, ksub), dtype=np.float32)
    lib().orc_opq_build_lut(_p(q), C.c_int(q.shape[0]), _p(centroid), _p(cb), C.c_int(M), C.c_int(ksub), _p(lut))
    return lut


def opq_adc_scan(lut, codes):
    lut = _f32(lut)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    M, ksub = lut.shape
    out = np.empty(codes.shape[0], dtype=np.float32)
    lib().orc_opq_adc_scan(_p(lut), C.c_int(M), C.c_int(ksub), _p(codes), C.c_int64(codes.shape[0]), _p(out))
    return out


def opq_query_scores(q_rot, coarse, cb, nk, row_list, row_group, codes, n_groups, clamp=1.0):
    q_rot, coarse, cb = _f32(q_rot), _f32(coarse), _f32(cb)
    row_list = np.ascontiguousarray(row_list, dtype=np.int32)
    row_group = np.ascontiguousarray(row_group, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    M, ksub, _ = cb.shape
    match = np.empty((q_rot.shape[0], n_groups), dtype=np.float32)
    lib().orc_opq_query_scores(_p(q_rot), C.c_int64(q_rot.shape[0]), C.c_int(q_rot.shape[1]), _p(coarse),
                               C.c_int(coarse.shape[0]), _p(cb), C.c_int(M), C.c_int(ksub), C.c_int(nk),
                               _p(row_list), _p(row_group), _p(codes), C.c_int64(codes.shape[0]),
                               C.c_int64(n_groups), C.c_float(clamp), _p(match))
    return match


def topk_pairs(score, k):
    score = _f32(score)
    out_s = np.empty(k, dtype=np.float32)
    out_i = np.empty(k, dtype=np.int64)
    lib().orc_topk_pairs(_p(score), C.c_int64(score.shape[0]), C.c_int(k), _p(out_s), _p(out_i))
    return out_s, out_i


def opq_search_flat(q_rot, coarse0, cb, codes, k, clamp=np.inf):
    """Flat (K=1) ADC top-k per query through the restatement: LUT -> scan -> clamp -> top-k."""
    q_rot = _f32(q_rot)
    D = np.empty((q_rot.shape[0], k), dtype=np.float32)
    I = np.empty((q_rot.shape[0], k), dtype=np.int64)
    for i in range(q_rot.shape[0]):
        lut = opq_build_lut(q_rot[i], coarse0, cb)
        s = opq_adc_scan(lut, codes)
        if np.isfinite(clamp):
            s = np.minimum(s, np.float32(clamp))
        D[i], I[i] = topk_pairs(s, k)
    return D, I


# ---------------------------------------------------------------------------- flat
def flat_search(metric, lanes, data, labels, queries, k):
    """metric 0 = 1-IP, 1 = L2 (fp32, `lanes` accumulators), 2 = L2SqrI (uint8 -> int32)."""
    labels = np.ascontiguousarray(labels, dtype=np.uint64)
    if metric == 2:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
    else:
        data, queries = _f32(data), _f32(queries)
    n, d = data.shape
    nq = queries.shape[0]
    Df = np.zeros((nq, k), dtype=np.float32)
    Di = np.zeros((nq, k), dtype=np.int32)
    L = np.zeros((nq, k), dtype=np.uint64)
    for i in range(nq):
        lib().orc_flat_search(C.c_int(metric), C.c_int(lanes), _p(data), _p(labels), C.c_int64(n), C.c_int64(d),
                              _p(queries[i]), C.c_int(k), _p(Df[i]), _p(Di[i]), _p(L[i]))
    return (Di if metric == 2 else Df), L


# ---------------------------------------------------------------------------- sq
def sq_encode(x, vmin, vdiff, l2norm=True):
    """Row-wise Int8Encode; returns (codes, x_after) -- the reference normalises x in place."""
    x = _f32(x).copy()
    vmin, vdiff = _f32(vmin), _f32(vdiff)
    codes = np.empty(x.shape, dtype=np.uint8)
    for i in range(x.shape[0]):
        lib().orc_sq_encode(_p(x[i]), _p(codes[i]), C.c_int(x.shape[1]), _p(vmin), _p(vdiff), C.c_int(int(l2norm)))
    return codes, x


def sq_decode(codes, vmin, vdiff, faiss_float=False):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    vmin, vdiff = _f32(vmin), _f32(vdiff)
    x = np.empty(codes.shape, dtype=np.float32)
    fn = lib().orc_sq_decode_faiss if faiss_float else lib().orc_sq_decode
    for i in range(codes.shape[0]):
        fn(_p(codes[i]), _p(x[i]), C.c_int(codes.shape[1]), _p(vmin), _p(vdiff))
    return x


def sq_train_minmax(x):
    x = _f32(x)
    vmin = np.empty(x.shape[1], dtype=np.float32)
    vdiff = np.empty(x.shape[1], dtype=np.float32)
    lib().orc_sq_train_minmax(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(vmin), _p(vdiff))
    return vmin, vdiff


# ---------------------------------------------------------------------------- front end (f-3)
def pca_project(x, mean, vectors, l2norm=True):
    """cvtk::PCAUtils::reduceDim (pca_train_project/pca_online/pca_utils.cc:25-35)."""
    x, vectors = _f32(x), _f32(vectors)
    mean_a = None if mean is None else _f32(mean).reshape(-1)
    y = np.empty((x.shape[0], vectors.shape[0]), dtype=np.float32)
    lib().orc_pca_project(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), None if mean_a is None else _p(mean_a), _p(vectors),
                          C.c_int(vectors.shape[0]), C.c_int(int(l2norm)), _p(y))
    return y


def rootsift(x, eps=1e-7):
    """siftsIDX::rootSift (hnsw_sifts_retrieval/siftsIndex.cpp:54-71); returns a new array."""
    y = _f32(x).copy()
    lib().orc_rootsift(_p(y), C.c_int64(y.shape[0]), C.c_int(y.shape[1]), C.c_float(eps))
    return y


# ---------------------------------------------------------------------------- training (f-4)
def kmeans(x, k, max_iter=0, seed=0):
    """The product's deterministic Lloyd iteration restated (yael's kmeans is un-vendored: parity unpinned there).
    Returns (centroids, assign, dist, iterations, mse)."""
    x = _f32(x)
    n, d = x.shape
    c = np.empty((k, d), dtype=np.float32)
    a = np.empty(n, dtype=np.int32)
    dist = np.empty(n, dtype=np.float32)
    mse = C.c_double(0.0)
    it = lib().orc_kmeans(_p(x), C.c_int64(n), C.c_int64(d), C.c_int(0), C.c_int(d), C.c_int(k), C.c_int(max_iter), C.c_uint64(seed),
                          _p(c), _p(a), _p(dist), C.byref(mse))
    if it < 0:
        raise ValueError("orc_kmeans: bad input")
    return c, a, dist, it, mse.value


def pq_train(x_raw, K, M, ksub=256, perm=None, max_iter=0, seed=0):
    """TrainPQ::IFVPQ restated over the Lloyd iteration above.  Returns (coarse, codebooks, mse[1+M])."""
    x_raw = _f32(x_raw)
    n, D = x_raw.shape
    perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    coarse = np.empty((max(K, 1), D), dtype=np.float32)
    cb = np.empty((M, ksub, D // M), dtype=np.float32)
    mse = np.zeros(1 + M, dtype=np.float64)
    rc = lib().orc_pq_train(_p(x_raw), C.c_int64(n), C.c_int(D), C.c_int(K), C.c_int(M), C.c_int(ksub),
                            None if perm_a is None else _p(perm_a), C.c_int(max_iter), C.c_uint64(seed), _p(coarse), _p(cb), _p(mse))
    if rc != 0:
        raise ValueError("orc_pq_train: bad input")
    return coarse, cb, mse


# ---------------------------------------------------------------------------- compiled reference
def parse_ref_opq(path: str) -> dict:
    b = open(path, "rb").read()
    assert b[:4] == b"ROPQ"
    h = np.frombuffer(b, dtype="<i4", count=12, offset=4)
    ver, D, K, M, ksub, n_rows, n_groups, nq, nk, kk, cons, nqf = [int(v) for v in h]
    off = 4 + 48
    fpf = np.frombuffer(b, dtype="<i4", count=nqf, offset=off).copy(); off += 4 * nqf
    lst = np.frombuffer(b, dtype="<i4", count=n_rows, offset=off).copy(); off += 4 * n_rows
    grp = np.frombuffer(b, dtype="<i4", count=n_rows, offset=off).copy(); off += 4 * n_rows
    codes = np.frombuffer(b, dtype="u1", count=n_rows * M, offset=off).reshape(n_rows, M).copy(); off += n_rows * M
    ms = np.frombuffer(b, dtype="<f4", count=nq * n_groups, offset=off).reshape(nq, n_groups).copy()
    off += 4 * nq * n_groups
    rec = np.dtype([("s", "<f4"), ("i", "<u4")])
    tk = np.frombuffer(b, dtype=rec, count=nq * kk, offset=off).reshape(nq, kk).copy(); off += 8 * nq * kk
    tf = np.frombuffer(b, dtype=rec, count=nqf * kk, offset=off).reshape(nqf, kk).copy(); off += 8 * nqf * kk
    assert off == len(b)
    return dict(D=D, K=K, M=M, ksub=ksub, n_rows=n_rows, n_groups=n_groups, nq=nq, nk=nk, topk=kk,
                consistent=cons, frames_per_file=fpf, row_list=lst, row_group=grp, codes=codes, match=ms,
                topk_score=tk["s"].copy(), topk_id=tk["i"].astype(np.int64), file_topk_score=tf["s"].copy(),
                file_topk_id=tf["i"].astype(np.int64))


def run_ref_opq(model: str, db_files, query_files, nk: int, topk: int, per_row: bool) -> dict:
    """Run the unmodified reference (Add / QueryThrehold / get_sort_results) and parse its dump."""
    exe = os.path.join(REF_DIR, "ref_opq")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.bin")
        cmd = [exe, "run", model, out, str(nk), str(topk), "1" if per_row else "0", str(len(db_files)), *db_files,
               str(len(query_files)), *query_files]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"ref_opq failed: {r.stderr}")
        return parse_ref_opq(out)


def bench_ref_opq(model: str, db_file: str, query_file: str, nk: int, topk: int, n_queries: int, threads: int,
                  tmpdir: str, repeat: int = 1) -> dict:
    """Time the unmodified reference's QueryThrehold + get_sort_results on n_queries rows over
    `threads` host threads; also returns its top-k (scores, ids) for the parity check."""
    exe = os.path.join(REF_DIR, "ref_opq")
    dump = os.path.join(tmpdir, "ref_topk.bin")
    cmd = [exe, "bench", model, db_file, query_file, str(nk), str(topk), str(n_queries), str(threads), tmpdir,
           str(repeat), dump]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"ref_opq bench failed: {r.stderr}")
    out = json.loads(r.stdout.strip().splitlines()[-1])
    raw = np.fromfile(dump, dtype=np.uint8)
    half = n_queries * topk * 4
    out["topk_score"] = raw[:half].view("<f4").reshape(n_queries, topk).copy()
    out["topk_id"] = raw[half:].view("<u4").reshape(n_queries, topk).astype(np.int64)
    os.unlink(dump)
    return out


def run_ref_flat(flavour: str, metric: str, data, labels, queries, k: int, save_index: str | None = None):
    """flavour: bf_sse | bf_avx | hnsw;  metric: ip | l2 | l2i.  Returns (dist[nq,k], label[nq,k])."""
    exe = os.path.join(REF_DIR, f"ref_flat_{flavour}")
    u8 = metric == "l2i"
    data = np.ascontiguousarray(data, dtype=np.uint8 if u8 else np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.uint8 if u8 else np.float32)
    n, d = data.shape
    nq = queries.shape[0]
    with tempfile.TemporaryDirectory() as td:
        dp, qp, lp, op = (os.path.join(td, s) for s in ("d.bin", "q.bin", "l.bin", "o.bin"))
        data.tofile(dp); queries.tofile(qp)
        np.ascontiguousarray(labels, dtype=np.uint64).tofile(lp)
        cmd = [exe, metric, dp, lp, qp, str(n), str(d), str(nq), str(k), op]
        if save_index:
            cmd.append(save_index)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"ref_flat failed: {r.stderr}")
        rec = np.dtype([("d", "<i4" if u8 else "<f4"), ("pad", "V4"), ("l", "<u8")]) if False else None
        raw = np.fromfile(op, dtype=np.uint8)
    # records are (dist 4B, label 8B) packed without padding
    raw = raw.reshape(nq, k, 12)
    dist = raw[:, :, :4].copy().view("<i4" if u8 else "<f4").reshape(nq, k)
    lab = raw[:, :, 4:].copy().view("<u8").reshape(nq, k)
    return dist, lab


def run_ref_int8_quan(x, vmin, vdiff, l2norm=True, via_conf: bool = False, source: int = 0) -> dict:
    """Run the UNMODIFIED cvtk::quant::Int8Quan (scalar_quantization/scalar_quantization/int8_quan.cc compiled in place
    against the stub faiss / nlohmann headers, oracle/Makefile) on rows x with the trained range [vmin | vdiff] written as
    an IxSQ model file.  via_conf drives the multi-model constructor (a JSON conf naming `source`+1 models; the queried
    one is the last).  Returns the reference's own outputs: codes / x_after (Int8Encode, :72-94), normed
    (L2NormalizeVector, :46-56), decode (Int8Decode(std::string&), :117-132) -- reference-run -- and codes_faiss /
    decode_faiss, which pass through the STUB codec (faiss itself is absent) and stay unpinned."""
    import struct
    exe = os.path.join(REF_DIR, "ref_int8_quan")
    x = _f32(x)
    n, d = x.shape
    vmin, vdiff = _f32(vmin), _f32(vdiff)

    def write_model(path, vm, vd):
        with open(path, "wb") as f:  # faiss 1.5.x IndexScalarQuantizer layout, SURVEY.md App. A-8
            f.write(b"IxSQ")
            f.write(struct.pack("<iqqqBi", d, 0, 1 << 20, 1 << 20, 1, 1))
            f.write(struct.pack("<iifQQ", 0, 0, 0.0, d, d))
            f.write(struct.pack("<Q", 2 * d))
            np.concatenate([vm, vd]).astype("<f4").tofile(f)
            f.write(struct.pack("<Q", 0))

    with tempfile.TemporaryDirectory() as td:
        model = os.path.join(td, "m.bin")
        write_model(model, vmin, vdiff)
        arg_model, num_source = model, 0
        if via_conf:
            conf = {}
            for i in range(source):  # decoys in front of the model under test
                dp = os.path.join(td, f"decoy{i}.bin")
                write_model(dp, vmin + np.float32(1.0 + i), vdiff * np.float32(2.0))
                conf[str(i)] = {"model_path": dp}
            conf[str(source)] = {"model_path": model}
            arg_model = os.path.join(td, "conf.json")
            json.dump(conf, open(arg_model, "w"))
            num_source = source + 1
        rows, out = os.path.join(td, "rows.f32"), os.path.join(td, "out.bin")
        x.tofile(rows)
        r = subprocess.run([exe, arg_model, str(num_source), str(source if via_conf else 0), rows, str(n), str(d), "1" if l2norm else "0", out],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"ref_int8_quan failed: {r.stderr}")
        b = open(out, "rb").read()
    status, n2, d2 = np.frombuffer(b, "<i4", 3)
    assert status == 1 and n2 == n and d2 == d, f"ref_int8_quan status {status}"
    off = 12
    nd = n * d

    def take(dtype, size):
        nonlocal off
        a = np.frombuffer(b, dtype, nd, off).reshape(n, d).copy()
        off += nd * size
        return a
    res = dict(codes=take("u1", 1), x_after=take("<f4", 4), normed=take("<f4", 4), decode=take("<f4", 4), codes_faiss=take("u1", 1),
               decode_faiss=take("<f4", 4))
    res["rc_bad_dims"] = int(np.frombuffer(b, "<i4", 1, off)[0])
    return res
