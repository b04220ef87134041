#!/usr/bin/env python
"""Small-batch / per-call latencies of the entry points the reference interface forces, next to the
reference's own CPU code on the same inputs (oracle/_ref, per-call time = difference of two runs so that load/build time
cancels).  Development aid (lives under tests/ because it times the CPU checker binaries); prints one JSON document.

  makeSearch shape   125 402 x 128 rootSIFT rows (hnsw_sifts_retrieval/makeIdx.cpp:303-309), 1 536 descriptors, k = 5, 1 - <a,b>:
                     one batch (searchKnnBatch) and one query per call (searchKnn as makeSearch.cpp:52 calls it)
  brute-force CLI    10 000 x 128, 100 queries, k = 100 (brute_force.cpp:15), fp32 and int8
  Int8Encode         one 128-d vector per call (int8_quan.cc:72-94)
  IVFOPQ::Query      one 8-frame query over a K = 256, 20 000-row index (scores for all groups)
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvt_b200 import capi, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def wall(fn, it, warm=3):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    return (time.perf_counter() - t0) / it


def ref_flat_per_query(flavour, metric, data, q, k):
    """seconds per searchKnn of the reference's BruteforceSearch on the host (single thread, as the reference runs it)."""
    exe = os.path.join(REF, f"ref_flat_{flavour}")
    if not os.path.exists(exe):
        return None
    n, d = data.shape
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        dp, qp, op = (os.path.join(td, s) for s in ("d.bin", "q.bin", "o.bin"))
        data.tofile(dp); q.tofile(qp)
        ts = []
        for nq in (1, 1 + min(len(q) - 1, 32)):
            t0 = time.perf_counter()
            subprocess.run([exe, metric, dp, "-", qp, str(n), str(d), str(nq), str(k), op], check=True, capture_output=True)
            ts.append((nq, time.perf_counter() - t0))
        return (ts[1][1] - ts[0][1]) / (ts[1][0] - ts[0][0])


def main():
    out = {}
    ctx = capi.Context(0)
    # ---- makeSearch shape
    n, nq, k = 125_402, 1_536, 5
    x = synth.sift_like(n, 128, seed=7)
    q = synth.sift_like(nq, 128, seed=8)
    idx = capi.FlatIndex(ctx, "ip", 128, n, order=4)
    idx.add(x, np.arange(n, dtype=np.uint64))
    idx.search(q[:64], k)
    t_batch = wall(lambda: idx.search(q, k), 5)
    t_one = wall(lambda: idx.search(q[:1], k), 200, warm=20)
    cpu = ref_flat_per_query("hnsw", "ip", x, q, k)
    out["makeSearch_shape"] = {"rows": n, "dim": 128, "descriptors": nq, "k": k,
                               "gpu_batch_ms_per_image": 1e3 * t_batch, "gpu_pair_elements_per_s": n * nq * 128 / t_batch,
                               "gpu_single_query_call_us": 1e6 * t_one,
                               "cpu_reference_bruteforce_ms_per_query": None if cpu is None else 1e3 * cpu,
                               "cpu_reference_bruteforce_ms_per_image": None if cpu is None else 1e3 * cpu * nq,
                               "note": "host buffers in and out (H2D + D2H + sync inside); CPU = the reference's BruteforceSearch::searchKnn, 1 thread"}
    idx.close()
    # ---- brute-force CLI shape, k = 100
    n, nq, k = 10_000, 100, 100
    x = synth.sift_like(n, 128, seed=9); q = synth.sift_like(nq, 128, seed=10)
    idx = capi.FlatIndex(ctx, "ip", 128, n, order=4)
    idx.add(x, np.arange(n, dtype=np.uint64))
    t = wall(lambda: idx.search(q, k), 50)
    cpu = ref_flat_per_query("bf_sse", "ip", x, q, k)
    out["brute_force_cli_fp32"] = {"rows": n, "queries": nq, "k": k, "gpu_ms": 1e3 * t, "cpu_reference_ms": None if cpu is None else 1e3 * cpu * nq}
    idx.close()
    rng = np.random.Generator(np.random.PCG64(3))
    xu = rng.integers(0, 256, (1_000_000, 128), dtype=np.uint8); qu = rng.integers(0, 256, (1024, 128), dtype=np.uint8)
    idx = capi.FlatIndex(ctx, "l2_u8", 128, len(xu))
    idx.add(xu, np.arange(len(xu), dtype=np.uint64))
    res = {}
    for kk in (10, 32, 100):
        res[f"k{kk}_ms_per_1024_queries"] = 1e3 * wall(lambda: idx.search(qu, kk), 5)
    out["int8_scan_1M_x_128"] = res
    idx.close()
    # ---- Int8Encode, one vector per call
    v = synth.sift_like(4096, 128, seed=11)
    vmin, vdiff = capi.SQ.train_minmax(ctx, v)
    sq = capi.SQ(ctx, vmin, vdiff)
    one = v[:1].copy()
    t1 = wall(lambda: sq.encode(one, l2norm=True), 300, warm=20)
    tb = wall(lambda: sq.encode(v, l2norm=True), 50)
    cpu_enc = None
    try:  # the reference's Int8Quan::Int8Encode on the host: per-vector time = difference of two runs of the reference-run driver
        from oracle import oracle as orc
        if orc.have_ref("ref_int8_quan"):
            big = synth.sift_like(200_000, 128, seed=15)
            t0 = time.perf_counter(); orc.run_ref_int8_quan(big[:1000], vmin, vdiff); ta = time.perf_counter() - t0
            t0 = time.perf_counter(); orc.run_ref_int8_quan(big, vmin, vdiff); tb2 = time.perf_counter() - t0
            cpu_enc = (tb2 - ta) / (len(big) - 1000) / 4.0   # the driver runs encode, normalise, decode and the faiss-path encode per row
    except Exception:
        pass
    out["int8_encode"] = {"single_vector_call_us": 1e6 * t1, "batch_4096_us_per_vector": 1e6 * tb / 4096,
                          "cpu_reference_us_per_vector_approx": None if cpu_enc is None else 1e6 * cpu_enc}
    sq.close()
    # ---- IVFOPQ::Query, 8 frames
    n = 20_000
    db = synth.sift_like(n, 128, seed=12)
    perm = synth.SHIPPED_REORDER_128
    coarse, cb = synth.train_pq_model(db[:, perm], 16, 256, 256, iters=3, train_rows=4000)
    pq = capi.PQIndex.create(ctx, coarse, cb, perm=perm)
    pq.add(db, (np.arange(n) // 20).astype(np.int32))
    frames = synth.sift_like(8, 128, seed=13)
    t8 = wall(lambda: pq.scores(frames, nprobe=3), 100, warm=10)
    t64 = wall(lambda: pq.search(synth.sift_like(64, 128, seed=14), 100, nprobe=1), 50, warm=5)
    cpu_q = None
    try:  # the unmodified IVFOPQ::QueryThrehold + get_sort_results on the host, 1 thread, 8 frames
        import shutil
        from oracle import oracle as orc
        if orc.have_ref("ref_opq"):
            td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                synth.write_opq_model(os.path.join(td, "m.model"), coarse, cb, perm)
                synth.write_feat_file(os.path.join(td, "db.bin"), db)
                synth.write_feat_file(os.path.join(td, "q.bin"), frames)
                r = orc.bench_ref_opq(os.path.join(td, "m.model"), os.path.join(td, "db.bin"), os.path.join(td, "q.bin"), nk=3, topk=5,
                                      n_queries=8, threads=1, tmpdir=td, repeat=5)
                cpu_q = r["query_s_mean"]
            finally:
                shutil.rmtree(td, ignore_errors=True)
    except Exception:
        pass
    out["ivfopq_query"] = {"rows": n, "K": 256, "groups": pq.n_groups, "frames": 8, "query_8_frames_us": 1e6 * t8,
                           "search_64_queries_top100_nprobe1_us": 1e6 * t64,
                           "cpu_reference_query_8_frames_us": None if cpu_q is None else 1e6 * cpu_q,
                           "cpu_note": "reference index with one videoId per ROW (20 000 groups), 1 thread; the GPU index has 1 000 videos"}
    pq.close()
    ctx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
