#!/usr/bin/env python
"""bench.py -- the contract benchmark of the quantized nearest-neighbour path.

    python bench.py --gpus N --steps K --warmup W                 # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU scan

Workload (default `cfg3`, BASELINE.json configs[2], the configuration the metric is quoted on):
OPQ M=16 8-bit sub-codes, 1M x 128-d SIFT-shaped unit-norm vectors, flat (K=1) ADC scan, top-100
(recall@10 reported), batch 4096, clamp threshold 1.0 as in the reference.  A "step" = one pass of
the hot path over one query batch: rotate -> LUT build -> ADC scan + top-k -> slice merge
[-> all-gather of per-rank top-k + merge when N > 1].  With N > 1 the SAME database and batch are split
over a (query chunk x row shard) grid of ranks ("scaling": "strong"): cvt_b200.sharded.plan_layout picks
the grid (row shards = N is the plain row-sharded layout; --row-shards forces it), and when the planner
picks something else the plain row-sharded layout is timed as well and reported under "row_sharded".

Prints ONE JSON line on rank 0 (see README/DESIGN for the fields).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cvt_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: n_rows, D, M, batch, k
    "cfg3": dict(n=1_000_000, D=128, M=16, B=4096, k=100, desc="cfg3: OPQ M=16 8-bit, 1M x 128 SIFT-shaped, flat ADC top-100, batch 4096"),
    "cfg3_small": dict(n=100_000, D=128, M=16, B=512, k=100, desc="cfg3 shape at 1/10 size (development)"),
    "cfg2": dict(n=1_000_000, D=128, M=0, B=1024, k=10, desc="cfg2: int8 scalar-quantized exact L2 scan, 1M x 128 SIFT-shaped, batch 1024, top-10"),
    "cfg4_shard": dict(n=1_250_000, D=512, M=32, B=4096, k=100, desc="cfg4 per-GPU shard: OPQ M=32, 1.25M x 512 CNN-like, ADC top-100, batch 4096"),
}
METRIC = "queries/sec, batched ADC top-k scan (HBM GB/s in roofline; recall@10 vs CPU reference)"


def env_int(name, default):
    return int(os.environ.get(name, default))


def host_cpu_model():
    """Model name of the host CPU the CPU baseline runs on (core count and model stated)."""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def make_inputs(wl):
    """Seeded host inputs, identical on every rank and for both arms."""
    n, D, M, B = wl["n"], wl["D"], wl["M"], wl["B"]
    if D == 128:
        db = synth.sift_like(n, D, seed=synth.SEED_DB)
        q = synth.sift_like(B, D, seed=synth.SEED_QUERY)
        perm = synth.SHIPPED_REORDER_128
    else:
        db = synth.cnn_like(n, D, seed=synth.SEED_DB)
        q = synth.cnn_like(B, D, seed=synth.SEED_QUERY)
        perm = synth.random_permutation(D)
    coarse, cb = synth.train_pq_model(db[:20000][:, perm], M, 256, 1, iters=6, seed=synth.SEED_KMEANS, train_rows=20000)
    return db, q, perm.astype(np.int32), coarse, cb


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def write_reference_inputs(td, db, q, perm, coarse, cb):
    """The reference's own byte formats: .model (IVFOPQ::LoadModel) + raw feature files."""
    model = os.path.join(td, "bench.model")
    synth.write_opq_model(model, coarse, cb, perm)
    dbf, qf = os.path.join(td, "db.bin"), os.path.join(td, "q.bin")
    synth.write_feat_file(dbf, db)
    synth.write_feat_file(qf, q)
    return model, dbf, qf


def run_reference_sample(wl, inputs, n_queries, repeat, threads=0):
    """Time the UNMODIFIED reference (oracle/_ref/ref_opq: IVFOPQ::Add to build, QueryThrehold +
    get_sort_results to search) on a bounded query sample over all host threads."""
    from oracle import oracle as orc  # test infrastructure: the checker / CPU baseline, never the product path
    if not orc.have_ref("ref_opq"):
        raise RuntimeError("oracle/_ref/ref_opq is missing (it is built in the container where /root/reference exists)")
    db, q, perm, coarse, cb = inputs
    if threads <= 0:
        # all host threads this process may use, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers,
        # which would otherwise silently turn the reference arm into a single-thread run
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    td = tempfile.mkdtemp(prefix="b200nn_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        model, dbf, qf = write_reference_inputs(td, db, q, perm, coarse, cb)
        r = orc.bench_ref_opq(model, dbf, qf, nk=1, topk=wl["k"], n_queries=n_queries, threads=threads, tmpdir=td, repeat=repeat)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    return r


def run_cfg2(args, wl, local_rank):
    """Secondary workload (BASELINE.json configs[1]): exact L2 scan over 8-bit scalar-quantized codes.
    SQ encode on the GPU (Int8Quan arithmetic), then the tcgen05 kind::i8 scan with fused top-k."""
    import torch
    from cvt_b200 import capi
    n, D, B, k = wl["n"], wl["D"], wl["B"], wl["k"]
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    x = synth.sift_like(n, D, seed=synth.SEED_DB)
    vmin, vdiff = capi.SQ.train_minmax(ctx, x[:200_000])
    sq = capi.SQ(ctx, vmin, vdiff)
    codes, _ = sq.encode(x, l2norm=True)
    qc, _ = sq.encode(synth.sift_like(B, D, seed=synth.SEED_QUERY), l2norm=True)
    labels = np.arange(n, dtype=np.uint64)
    idx = capi.FlatIndex(ctx, "l2_u8", D, n)
    idx.add(codes, labels)
    q_pinned = torch.from_numpy(qc).pin_memory()
    q_dev = q_pinned.cuda()
    od = torch.empty((B, k), dtype=torch.int32, device="cuda")
    ol = torch.empty((B, k), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(args.warmup):
        idx.search_dev(q_dev.data_ptr(), B, k, od.data_ptr(), ol.data_ptr())
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        idx.search_dev(q_dev.data_ptr(), B, k, od.data_ptr(), ol.data_ptr())
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    ms_per_step = float(np.mean(ms))
    res_l, res_d = ol.cpu().numpy(), od.cpu().numpy()
    Dh = np.empty((B, k), dtype=np.int32)
    Lh = np.empty((B, k), dtype=np.uint64)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Dh, Lh = idx.search(qc, k)
    e2e_s = (time.perf_counter() - t0) / args.steps
    assert np.array_equal(Lh.astype(np.int64), res_l) and np.array_equal(Dh, res_d)
    ops = 2.0 * B * n * D
    peak = None
    try:
        peak = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        peak = 2.0 * 1590.0
    line = {"metric": "queries/sec, batched exact int8 L2 top-10 scan", "value": B / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8 x u8 -> s32", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_rows": n, "dim": D, "batch": B, "k": k, "l2": "256 MB buffer written between timed iterations"},
            "roofline": {"bound": "tensor", "kernel": "u8_scan_tc_kernel (+merge)", "achieved": ops / (ms_per_step * 1e-3) / 1e12, "peak": peak,
                         "unit": "TOP/s", "frac": ops / (ms_per_step * 1e-3) / 1e12 / peak, "traffic": None,
                         "peak_source": "2 x measured bf16 cuBLAS TF/s (MEASURED_PEAKS.json): dense int8 is nominally twice bf16"},
            "e2e": {"value": B / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(B * D), "d2h_bytes_per_step": int(B * k * 12)},
            "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline:
        from oracle import oracle as orc  # checker / CPU baseline only
        ns = 4
        t0 = time.perf_counter()
        odist, olab = orc.flat_search(2, 0, codes, labels, qc[:ns], k)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"first {ns} queries, oracle restatement of BruteforceSearch<int> + L2SqrI (scalar, 1 thread)"}
        line["parity"] = {"queries_checked": ns, "labels_identical": bool(np.array_equal(olab.astype(np.int64), res_l[:ns])),
                          "dists_identical": bool(np.array_equal(odist, res_d[:ns]))}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=64, help="queries of the batch timed on the host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--row-shards", type=int, default=0, help="row shards R of the rank grid (0: planned; N: plain row sharding)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n, D, M, B, k = wl["n"], wl["D"], wl["M"], wl["B"], wl["k"]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.workload == "cfg2":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is wired for the ADC workloads; cfg2's CPU baseline is in its own line"}))
            return 0
        inputs = make_inputs(wl)
        ns = min(B, args.cpu_sample)
        r = run_reference_sample(wl, inputs, ns, repeat=args.steps + args.warmup)
        qps = ns / r["query_s_mean"]
        line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["query_s_mean"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "n_rows": n, "dim": D, "M": M, "batch": B, "k": k, "nprobe": 1,
                           "step": f"{ns}-query sample of the batch per step"},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": r["threads"], "kind": "reference", "host": host_cpu_model(),
                                 "sample": f"{ns} of {B} queries per step x {args.steps + args.warmup} steps, unmodified IVFOPQ::QueryThrehold + "
                                           f"get_sort_results over {r['threads']} OpenMP threads; index built by IVFOPQ::Add in {r['build_s']:.1f}s"},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA)
    if args.workload == "cfg2":
        if rank != 0:
            return 0
        return run_cfg2(args, wl, local_rank)
    import torch
    import torch.distributed as dist
    from cvt_b200 import capi, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # a dedicated non-default stream shared by torch (flush, events, NCCL ordering) and the library
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    inputs = make_inputs(wl)
    db, q, perm, coarse, cb = inputs
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    R, Qc = (args.row_shards, world // max(1, args.row_shards)) if args.row_shards else sharded.plan_layout(world, n, B, M, k, sms)
    if R < 1 or R * Qc != world:
        raise SystemExit(f"bench.py: --row-shards {args.row_shards} must divide --gpus {world}")
    ctx = capi.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)

    def make_layout(row_shards):
        r, _ = sharded.grid_coords(rank, row_shards)
        s_lo, s_hi = sharded.shard_bounds(n, row_shards, r)
        idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
        idx.add(db[s_lo:s_hi])
        return idx, sharded.make_gpu_sharded(ctx, idx, dist, rank, world, id_base=s_lo, nprobe=1, row_shards=row_shards), s_lo, s_hi

    index, sh, lo, hi = make_layout(R)
    q_lo, q_hi, _ = sharded.query_chunk(B, Qc, sharded.grid_coords(rank, R)[1])

    q_pinned = torch.from_numpy(q).pin_memory()
    q_dev = q_pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_d_host = torch.empty((B, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((B, k), dtype=torch.int64).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(sh_, index_, steps):
        """`steps` timed steps (CUDA events on the launching stream, L2 flushed before each, barrier +
        synchronize on both sides) -> (ms per step = max over ranks, per-stage ms of this rank, result)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stages = []
        barrier()
        for s_ in range(steps):
            flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            ev[s_][0].record()
            dd_, ii_ = sh_.search(q_dev, k)
            ev[s_][1].record()
            ev[s_][1].synchronize()
            stages.append(index_.last_timing())
        barrier()
        total = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()) / steps, stages, (dd_, ii_)

    # ---- device-resident timing ("value"): inputs in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs a moment to come up: start it before the warm-up, read it after the timed region
    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        dd, ii = sh.search(q_dev, k)
    barrier()
    # nvidia-smi delivers a sample every ~50-100 ms; when the whole timed region is shorter than that (multi-GPU
    # steps of ~2 ms) keep the GPUs under the same load for ~1 s more so the clocks line holds samples taken under load
    t_step = torch.tensor([(time.perf_counter() - t_w0) / args.warmup], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
    extra_warm = 0
    if float(t_step.item()) * args.steps < 1.0:
        extra_warm = int(min(2000, 1.0 / max(float(t_step.item()), 1e-4)))
        for _ in range(extra_warm):
            dd, ii = sh.search(q_dev, k)
        barrier()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ms_per_step, stage_ms, (dd, ii) = timed_steps(sh, index, args.steps)
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    qps = B / (ms_per_step * 1e-3)
    result_ids = ii.cpu().numpy()
    result_d = dd.cpu().numpy()
    scan_ms = float(np.mean([t["scan_ms"] for t in stage_ms]))

    # ---- the plain row-sharded layout (row shards = N, queries replicated) beside the planned grid
    row_sharded = None
    if world > 1 and R != world:
        index_rs, sh_rs, lo_rs, hi_rs = make_layout(world)
        for _ in range(args.warmup):
            sh_rs.search(q_dev, k)
        ms_rs, stages_rs, (d_rs, i_rs) = timed_steps(sh_rs, index_rs, args.steps)
        assert np.array_equal(i_rs.cpu().numpy(), result_ids) and np.array_equal(d_rs.cpu().numpy().view(np.uint32), result_d.view(np.uint32)), \
            "row-sharded layout and planned grid disagree"
        scan_rs = float(np.mean([t["scan_ms"] for t in stages_rs]))
        row_sharded = {"value": B / (ms_rs * 1e-3), "unit": "queries/s", "ms_per_step": ms_rs, "rows_per_gpu": hi_rs - lo_rs,
                       "scan_kernel_ms": scan_rs, "roofline_frac": float(B) * (hi_rs - lo_rs) * M / (scan_rs * 1e-3) / 1e9 / measured_peak_hbm()[0]}
        index_rs.close()

    # ---- end-to-end through the public C-ABI call with HOST buffers (H2D + D2H inside the timed region)
    def e2e_step():
        if world == 1:
            index.search_host_ptr(q_pinned.data_ptr(), B, k, 1, out_d_host.data_ptr(), out_i_host.data_ptr())
        else:
            qd = q_pinned.to(dev, non_blocking=True)
            d2, i2 = sh.search(qd, k)
            out_d_host.copy_(d2, non_blocking=True)
            out_i_host.copy_(i2, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_qps = B / float(e2e_s.item())
    assert np.array_equal(out_i_host.numpy(), result_ids), "e2e path and device path disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_hbm()
    alg_bytes = float(q_hi - q_lo) * (hi - lo) * M  # every query of this rank's chunk "reads" every code byte of its shard once
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (LUT sums) over u8 codes", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_rows": n, "rows_per_gpu": hi - lo, "queries_per_gpu": q_hi - q_lo, "dim": D, "M": M, "ksub": 256,
                       "batch": B, "k": k, "nprobe": 1, "clamp": 1.0,
                       "parallelism": (f"{R} row shard(s) x {Qc} query chunk(s) ({'planned' if not args.row_shards else 'forced'}), one all-gather of top-k keys"
                                       if world > 1 else "single GPU"),
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "extra_untimed_warmup_steps_for_clock_sampling": extra_warm, "seeds": "SURVEY.md §8(d)"},
            "roofline": {"bound": "hbm", "kernel": "adc_scan_topk_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_ms, "scan_launches_per_step": 1,
                         "note": "algorithmic bytes = batch x shard rows x M; binding unit is the shared-memory gather pipe (DESIGN.md)"},
            "stage_ms": {kk: float(np.mean([t[kk] for t in stage_ms])) for kk in stage_ms[0]},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(B * D * 4), "d2h_bytes_per_step": int(B * k * 12)},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall}
    if row_sharded is not None:
        line["row_sharded"] = row_sharded

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on a bounded sample, + parity
    if world == 1 and not args.no_cpu_baseline:
        try:
            ns = min(B, args.cpu_sample)
            r = run_reference_sample(wl, inputs, ns, repeat=1)
            cpu_qps = ns / r["query_s_mean"]
            ids_equal = bool(np.array_equal(r["topk_id"], result_ids[:ns]))
            rel = np.abs(r["topk_score"] - result_d[:ns]) / np.maximum(np.abs(r["topk_score"]), 1e-30)
            recall10 = float(np.mean([len(set(r["topk_id"][i, :10]) & set(result_ids[i, :10])) / 10.0 for i in range(ns)]))
            line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": r["threads"], "kind": "reference", "host": host_cpu_model(),
                                    "sample": f"first {ns} of {B} queries, unmodified IVFOPQ::QueryThrehold + get_sort_results over "
                                              f"{r['threads']} OpenMP threads (index built by IVFOPQ::Add, {r['build_s']:.1f}s, untimed)"}
            line["parity"] = {"vs": "unmodified reference (oracle/_ref/ref_opq)", "queries_checked": ns, "topk_ids_identical": ids_equal,
                              "recall_at_10": recall10, "max_rel_dist_err": float(rel.max()),
                              "scores_bit_identical": bool(np.array_equal(r["topk_score"].view(np.uint32), result_d[:ns].view(np.uint32)))}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    else:
        line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference",
                                "sample": "timed at N=1 only (see the N=1 line)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
