#!/usr/bin/env python
"""bench.py -- the contract benchmark of the quantized nearest-neighbour path.

    python bench.py --gpus N --steps K --warmup W                   # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU scan

Workloads (`--workload`, BASELINE.json configs; default `cfg3`, the configuration the metric is quoted on):
    cfg3  OPQ M=16 8-bit sub-codes, 1M x 128-d SIFT-shaped unit-norm rows, flat (K=1) ADC scan, top-100, batch 4096
    cfg4  OPQ M=32, 10M x 512-d ReLU-sparse "CNN fc" rows, ADC top-100, batch 4096 (BASELINE leaves the batch open)
    cfg5  OPQ M=16, 100M x 128-d SIFT-shaped rows, ADC top-100, batch 16384
    cfg2  exact L2 scan over int8 scalar-quantized codes (tensor cores), 1M x 128, batch 1024, top-10 (1 GPU)
A "step" = one pass of the hot path over one query batch: rotate -> LUT build -> ADC scan + top-k -> slice merge
[-> ONE ncclAllGather of the per-rank top-k records + merge when N > 1].  With N > 1 the SAME database and batch are split
over a (row shard x query chunk) grid of ranks ("scaling": "strong"): b200nn_plan_layout picks the grid (row shards = N is
the plain row-sharded layout of SURVEY.md 8(e); --row-shards forces it) and when the planner picks something else the plain
row-sharded layout is timed in the same run and reported under "row_sharded".  The whole multi-GPU step runs inside the
C library (b200nn_comm_* / b200nn_pq_search_sharded_dev); torch.distributed only launches the ranks, carries the NCCL id
and reduces the timings.

cfg4 / cfg5 never materialise their raw vectors (51 GB / 20 GB): every rank generates its rows on the device in seeded
chunks (torch, counter = chunk index, so any N produces the same database) and encodes them chunk by chunk; only the codes
stay.  The CPU reference arm and the parity check then work on a stated row subsample (the scan is linear
in rows) plus a numpy restatement of the FULL scan for two queries.

Prints ONE JSON line on rank 0 (fields: README / DESIGN).
"""
from __future__ import annotations

import argparse
import json
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
    # name: n_rows, D, M, batch, k, row generation ("host": numpy, identical on every rank; "device": seeded torch chunks)
    "cfg3": dict(n=1_000_000, D=128, M=16, B=4096, k=100, gen="host", desc="cfg3: OPQ M=16 8-bit, 1M x 128 SIFT-shaped, flat ADC top-100, batch 4096"),
    "cfg3_small": dict(n=100_000, D=128, M=16, B=512, k=100, gen="host", desc="cfg3 shape at 1/10 size (development)"),
    "cfg2": dict(n=1_000_000, D=128, M=0, B=1024, k=10, gen="host", desc="cfg2: int8 scalar-quantized exact L2 scan, 1M x 128 SIFT-shaped, batch 1024, top-10"),
    "cfg4": dict(n=10_000_000, D=512, M=32, B=4096, k=100, gen="device", desc="cfg4: OPQ M=32 8-bit, 10M x 512 CNN-fc-shaped, flat ADC top-100, batch 4096"),
    "cfg4_shard": dict(n=1_250_000, D=512, M=32, B=4096, k=100, gen="device", desc="cfg4 per-GPU shard: OPQ M=32, 1.25M x 512 CNN-like, ADC top-100, batch 4096"),
    "cfg5": dict(n=100_000_000, D=128, M=16, B=16384, k=100, gen="device", desc="cfg5: OPQ M=16 8-bit, 100M x 128 SIFT-shaped, flat ADC top-100, batch 16384"),
    "cfg5_small": dict(n=4_000_000, D=128, M=16, B=2048, k=100, gen="device", desc="cfg5 shape at 1/25 size (development)"),
}
METRIC = "queries/sec, batched ADC top-k scan (HBM GB/s in roofline; recall@10 vs CPU reference)"
TRAIN_ROWS = 100_000   # codebooks: seeded Lloyd, 25 iterations, first 100 k rows
TRAIN_ITERS = 25
GEN_CHUNK = 500_000    # rows per generated chunk of a device-generated database (divides every shard size used)
SUBSAMPLE_ROWS = 1_000_000  # rows the CPU reference scans for a device-generated workload (the first rows of the database)
SUBSAMPLE_ROWS_PORT = 200_000  # ... when the CPU arm is the scalar oracle port (M = 32: it also has to encode the rows, 1 thread)


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


# ----------------------------------------------------------------------------------------------- inputs
class DeviceRows:
    """Seeded rows generated on the device, chunk by chunk.  Chunk c = global rows [c*GEN_CHUNK, (c+1)*GEN_CHUNK)
    comes from a generator seeded with (seed, c), so every rank count sees the same database."""

    def __init__(self, kind, D, seed, dev):
        import torch
        self.torch, self.kind, self.D, self.seed, self.dev = torch, kind, D, seed, dev
        if kind == "sift":
            rng_c = np.random.Generator(np.random.PCG64(0xC0FFEE ^ D))  # the cluster centres of synth.sift_like
            self.centres = torch.from_numpy(rng_c.random((1024, D), dtype=np.float32)).to(dev)

    def chunk(self, c, rows=GEN_CHUNK):
        torch = self.torch
        g = torch.Generator(device=self.dev)
        g.manual_seed((self.seed * 1_000_003 + c) & 0x7FFFFFFFFFFFFFFF)
        x = torch.randn((rows, self.D), generator=g, device=self.dev, dtype=torch.float32)
        if self.kind == "sift":  # synth.sift_like: clipped integer histograms around 1024 centres, rootSIFT, L2 norm
            x.abs_().mul_(40.0)
            idx = (torch.arange(rows, device=self.dev) + c * rows) % 1024
            x.add_(self.centres[idx], alpha=30.0)
            x.floor_().clamp_(max=255.0)
            x.div_(x.sum(1, keepdim=True).add_(1e-7)).sqrt_()
        else:                    # synth.cnn_like: ReLU-sparse, L2 norm
            x.clamp_(min=0.0)
        x.div_(x.norm(dim=1, keepdim=True).clamp_(min=1e-12))
        return x

    def rows(self, lo, hi):
        """Yields (first global row, tensor) pieces covering [lo, hi)."""
        for c in range(lo // GEN_CHUNK, (hi + GEN_CHUNK - 1) // GEN_CHUNK):
            x = self.chunk(c)
            a, b = max(lo, c * GEN_CHUNK) - c * GEN_CHUNK, min(hi, (c + 1) * GEN_CHUNK) - c * GEN_CHUNK
            yield c * GEN_CHUNK + a, (x if (a == 0 and b == GEN_CHUNK) else x[a:b].contiguous())

    def host(self, lo, hi):
        return np.concatenate([x.cpu().numpy() for _, x in self.rows(lo, hi)])


def make_inputs(wl, dev=None):
    """Seeded inputs, identical on every rank and for both arms.  Returns a dict: q [B, D] (host), perm, coarse, cb and
    either db (host rows, gen == "host") or rows (a DeviceRows factory) + db_head (the first SUBSAMPLE_ROWS rows on the host)."""
    n, D, M, B = wl["n"], wl["D"], wl["M"], wl["B"]
    kind = "sift" if D == 128 else "cnn"
    perm = synth.SHIPPED_REORDER_128 if D == 128 else synth.random_permutation(D)
    inp = dict(perm=perm.astype(np.int32))
    if wl["gen"] == "host":
        db = synth.sift_like(n, D, seed=synth.SEED_DB) if kind == "sift" else synth.cnn_like(n, D, seed=synth.SEED_DB)
        inp["q"] = synth.sift_like(B, D, seed=synth.SEED_QUERY) if kind == "sift" else synth.cnn_like(B, D, seed=synth.SEED_QUERY)
        inp["db"] = db
        train = db[:TRAIN_ROWS]
    else:
        import torch
        if dev is None:
            if not torch.cuda.is_available():
                raise SystemExit("bench.py: the device-generated workloads need a CUDA device for the row generator")
            dev = torch.device("cuda", env_int("LOCAL_RANK", 0))
        rows = DeviceRows(kind, D, synth.SEED_DB, dev)
        inp["rows"] = rows
        inp["q"] = DeviceRows(kind, D, synth.SEED_QUERY, dev).chunk(0, rows=B).cpu().numpy()
        inp["db_head"] = None  # fetched on demand (host_head)
        train = rows.host(0, min(n, TRAIN_ROWS))
    inp["coarse"], inp["cb"] = synth.train_pq_model_bench(train[:, perm], M, 256, iters=TRAIN_ITERS, seed=synth.SEED_KMEANS)
    return inp


def host_head(wl, inp):
    """The first min(n, SUBSAMPLE_ROWS) rows of the database on the host (what the CPU reference scans)."""
    m = min(wl["n"], SUBSAMPLE_ROWS if wl["M"] <= 16 else SUBSAMPLE_ROWS_PORT)
    if "db" in inp:
        return inp["db"][:m]
    if inp.get("db_head") is None:
        inp["db_head"] = inp["rows"].host(0, m)
    return inp["db_head"]


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


# ----------------------------------------------------------------------------------------------- CPU arms
def run_reference_sample(wl, inp, n_queries, repeat, threads=0):
    """Time the UNMODIFIED reference (oracle/_ref/ref_opq: IVFOPQ::Add to build, QueryThrehold + get_sort_results to search)
    on a bounded sample -- n_queries queries over the first min(n, SUBSAMPLE_ROWS) rows -- over all host threads.
    M = 32 has no reference implementation (`uchar PQindex[16]`, IVFOPQ.h:28): there the oracle restatement (1 thread) runs."""
    from oracle import oracle as orc  # test infrastructure: the checker / CPU baseline, never the product path
    head = host_head(wl, inp)
    q = inp["q"][:n_queries]
    if wl["M"] > 16:
        xr = orc.opq_reorder(head, inp["perm"])
        codes = orc.opq_pq_encode(xr, inp["coarse"], np.zeros(len(head), np.int32), inp["cb"])
        qr = orc.opq_reorder(q, inp["perm"])
        best, total = 1e30, 0.0
        for _ in range(repeat):
            t0 = time.perf_counter()
            od, oi = orc.opq_search_flat(qr, inp["coarse"][0], inp["cb"], codes, wl["k"], clamp=1.0)
            dt = time.perf_counter() - t0
            total += dt
            best = min(best, dt)
        return dict(kind="port", threads=1, build_s=0.0, query_s_mean=total / repeat, topk_score=od, topk_id=oi, rows=len(head))
    if not orc.have_ref("ref_opq"):
        raise RuntimeError("oracle/_ref/ref_opq is missing (it is built in the container where /root/reference exists)")
    if threads <= 0:
        # all host threads this process may use, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers,
        # which would otherwise silently turn the reference arm into a single-thread run
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    td = tempfile.mkdtemp(prefix="b200nn_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        model, dbf, qf = os.path.join(td, "bench.model"), os.path.join(td, "db.bin"), os.path.join(td, "q.bin")
        synth.write_opq_model(model, inp["coarse"], inp["cb"], inp["perm"])
        synth.write_feat_file(dbf, head)
        synth.write_feat_file(qf, q)
        r = orc.bench_ref_opq(model, dbf, qf, nk=1, topk=wl["k"], n_queries=n_queries, threads=threads, tmpdir=td, repeat=repeat)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    r["kind"] = "reference"
    r["rows"] = len(head)
    return r


def cpu_sample_text(wl, r, ns, steps=None):
    n, m = wl["n"], r["rows"]
    what = (f"unmodified IVFOPQ::QueryThrehold + get_sort_results over {r['threads']} OpenMP threads (index built by IVFOPQ::Add, "
            f"{r['build_s']:.1f}s, untimed)") if r["kind"] == "reference" else \
        "oracle restatement of the ADC scan + get_sort_results, 1 thread (M = 32 has no reference implementation: uchar PQindex[16])"
    rows = f"all {n} rows" if m == n else f"the first {m} of {n} rows (value scaled by {m}/{n}: the scan is linear in rows)"
    return f"first {ns} of {wl['B']} queries{'' if steps is None else f' per step x {steps} steps'} over {rows}; {what}"


def numpy_adc_topk(lut, codes, k, clamp, id_base):
    """Plain numpy restatement of IVFOPQ.cpp:300-309 + common.h:25-37 for ONE query over a block of rows: sequential fp32 sum
    over m from 0.0f, clamp, k smallest (score, id).  lut [M, 256] f32, codes [n, M] u8."""
    s = np.zeros(codes.shape[0], dtype=np.float32)
    for m in range(codes.shape[1]):
        s = s + lut[m][codes[:, m]]
    s = np.minimum(s, np.float32(clamp))
    ids = np.arange(codes.shape[0], dtype=np.int64) + id_base
    if len(s) > 4 * k:
        part = np.argpartition(s, 4 * k)[:4 * k]
        thr = s[part].max()
        part = np.nonzero(s <= thr)[0]
        s, ids = s[part], ids[part]
    order = np.lexsort((ids, s))[:k]
    return s[order], ids[order]


# ----------------------------------------------------------------------------------------------- cfg2
def run_cfg2(args, wl, local_rank):
    """Secondary workload (BASELINE.json configs[1]): exact L2 scan over 8-bit scalar-quantized codes.
    SQ encode on the GPU (Int8Quan arithmetic), then the tcgen05 kind::i8 scan with fused top-k."""
    import torch
    from cvt_b200 import capi
    n, D, B, k = wl["n"], wl["D"], wl["B"], args.k or wl["k"]
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
    ms_per_step = float(np.median(ms))
    res_l, res_d = ol.cpu().numpy(), od.cpu().numpy()
    Dh = np.empty((B, k), dtype=np.int32)
    Lh = np.empty((B, k), dtype=np.uint64)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Dh, Lh = idx.search(qc, k)
    e2e_s = (time.perf_counter() - t0) / args.steps
    assert np.array_equal(Lh.astype(np.int64), res_l) and np.array_equal(Dh, res_d)
    ops = 2.0 * B * n * D
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk_ghz = (clocks.get("sm_mhz") or 1965.0) / 1e3
    tmem_floor_ms = (-(-B // 128)) * (-(-n // 256)) * (128 * 256 * 4 / 64.0) / sms / (clk_ghz * 1e9) * 1e3
    peak, peak_src = 2.0 * 1590.0, "fallback: 2 x 1590 TF/s bf16 (B200_PROFILING.md)"
    pp = os.path.join(ROOT, "profiles", "i8_mma_peak.json")  # a bare tcgen05 kind::i8 MMA loop, measured on this pool (tools/i8_peak)
    try:
        if os.path.exists(pp):
            peak, peak_src = float(json.load(open(pp))["tops"]), "measured: bare tcgen05.mma kind::i8 loop on all SMs (profiles/i8_mma_peak.json)"
        else:
            peak = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            peak_src = "2 x measured bf16 cuBLAS TF/s (MEASURED_PEAKS.json): dense int8 is nominally twice bf16"
    except Exception:
        pass
    traffic = None  # dram bytes of ONE scan launch of exactly this shape from an ncu --set full capture, else null
    try:
        for e in json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json"))).get("u8_captures", []):
            if e.get("rows") == n and e.get("queries") == B and e.get("k") == k and e.get("dim") == D:
                traffic = e.get("dram_bytes_per_launch")
    except Exception:
        traffic = None
    line = {"metric": "queries/sec, batched exact int8 L2 top-k scan", "value": B / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_per_step_mean": float(np.mean(ms)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 x u8 -> s32", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_rows": n, "dim": D, "batch": B, "k": k, "l2": "256 MB buffer written between timed iterations"},
            "roofline": {"bound": "tensor", "kernel": "u8_scan_tc_kernel (+merge)", "achieved": ops / (ms_per_step * 1e-3) / 1e12, "peak": peak,
                         "unit": "TOP/s", "frac": ops / (ms_per_step * 1e-3) / 1e12 / peak, "traffic": traffic, "peak_source": peak_src,
                         # what physically binds the fused-epilogue GEMM: every int32 accumulator is read back through tcgen05.ld
                         # (TMEM read port: 64 B/clk per SM, B300_MICROARCH.md) -- 2048 clk per 128 x 256 tile against 512 clk of MMA
                         "binding_unit": "TMEM read port (tcgen05.ld 64 B/clk/SM): 128 KB of accumulators per 128x256 tile",
                         "binding_floor_ms": tmem_floor_ms, "binding_frac": tmem_floor_ms / ms_per_step},
            "e2e": {"value": B / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(B * D), "d2h_bytes_per_step": int(B * k * 12)},
            "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline:
        from oracle import oracle as orc  # checker / CPU baseline only
        ns = 4
        t0 = time.perf_counter()
        odist, olab = orc.flat_search(2, 0, codes, labels, qc[:ns], k)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": 1, "kind": "port", "host": host_cpu_model(),
                                "sample": f"first {ns} queries, oracle restatement of BruteforceSearch<int> + L2SqrI (scalar, 1 thread)"}
        line["parity"] = {"queries_checked": ns, "labels_identical": bool(np.array_equal(olab.astype(np.int64), res_l[:ns])),
                          "dists_identical": bool(np.array_equal(odist, res_d[:ns]))}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- main
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
    ap.add_argument("--k", type=int, default=0, help="neighbours (default: the workload's)")
    ap.add_argument("--no-full-parity", action="store_true", help="skip the numpy restatement of the full scan (two queries)")
    ap.add_argument("--no-k10", action="store_true", help="skip the extra k = 10 timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = dict(WORKLOADS[args.workload])
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    n, D, M, B = wl["n"], wl["D"], wl["M"], wl["B"]
    k = args.k or wl["k"]
    wl["k"] = k

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.workload == "cfg2":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm is wired for the ADC workloads; cfg2's CPU baseline is in its own line"}))
            return 0
        inp = make_inputs(wl)
        ns = min(B, args.cpu_sample)
        r = run_reference_sample(wl, inp, ns, repeat=args.steps + args.warmup)
        qps = ns / r["query_s_mean"] * (r["rows"] / n)
        line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["query_s_mean"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "n_rows": n, "dim": D, "M": M, "batch": B, "k": k, "nprobe": 1,
                           "step": f"{ns}-query sample of the batch per step" + ("" if r["rows"] == n else f" over the first {r['rows']} rows")},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": r["threads"], "kind": r["kind"], "host": host_cpu_model(),
                                 "sample": cpu_sample_text(wl, r, ns, args.steps + args.warmup)},
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
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # a dedicated non-default stream shared by torch (flush, events) and the library (scan, ncclAllGather, merge)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    inp = make_inputs(wl, dev)
    q, perm, coarse, cb = inp["q"], inp["perm"], inp["coarse"], inp["cb"]
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    R, Qc = (args.row_shards, world // max(1, args.row_shards)) if args.row_shards else capi.plan_layout(world, n, B, M, k, sms)
    if R < 1 or R * Qc != world:
        raise SystemExit(f"bench.py: --row-shards {args.row_shards} must divide --gpus {world}")
    ctx = capi.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    comm = None
    if world > 1:  # the library's own communicator: rank 0 draws the NCCL id, torch.distributed only carries it
        idt = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        torch.cuda.synchronize()
        comm = capi.Comm(ctx, rank, world, bytes(idt.cpu().numpy().tobytes()))

    t_build0 = time.perf_counter()

    def make_layout(row_shards):
        r, _ = sharded.grid_coords(rank, row_shards)
        s_lo, s_hi = sharded.shard_bounds(n, row_shards, r)
        idx = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
        if "db" in inp:
            idx.add(inp["db"][s_lo:s_hi])
        else:
            for _, x in inp["rows"].rows(s_lo, s_hi):  # generate -> rotate -> encode, chunk by chunk: only the codes stay
                torch.cuda.current_stream().synchronize()
                idx.add_dev(x.data_ptr(), x.shape[0])
                ctx.synchronize()
        return idx, s_lo, s_hi

    index, lo, hi = make_layout(R)
    build_s = time.perf_counter() - t_build0
    q_lo, q_hi, _ = sharded.query_chunk(B, Qc, sharded.grid_coords(rank, R)[1])

    q_pinned = torch.from_numpy(q).pin_memory()
    q_dev = q_pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_d = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    out_d_host = torch.empty((B, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((B, k), dtype=torch.int64).pin_memory()

    def step(index_, row_shards, id_base, qd=q_dev):
        """One step through the C ABI: everything (local scan, the all-gather, the merge) is enqueued by the library."""
        if world == 1:
            index_.search_dev(qd.data_ptr(), B, k, 1, out_d.data_ptr(), out_i.data_ptr(), 0, 0)
        else:
            index_.search_sharded_dev(comm, row_shards, qd.data_ptr(), B, k, 1, id_base, out_d.data_ptr(), out_i.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(index_, row_shards, id_base, steps):
        """`steps` timed steps (CUDA events on the launching stream, L2 flushed before each, barrier + synchronize on both
        sides) -> (per-step ms, max over ranks; per-stage ms of this rank)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stages = []
        barrier()
        for s_ in range(steps):
            flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            ev[s_][0].record()
            step(index_, row_shards, id_base)
            ev[s_][1].record()
            ev[s_][1].synchronize()
            stages.append(index_.last_timing())
        barrier()
        per = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(per, op=dist.ReduceOp.MAX)
        return per.cpu().numpy(), stages

    # ---- device-resident timing ("value"): inputs in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs a moment to come up: start it before the warm-up, read it after the timed region
    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        step(index, R, lo)
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
            step(index, R, lo)
        barrier()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    per_step, stage_ms = timed_steps(index, R, lo, args.steps)
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    ms_per_step = float(np.median(per_step))
    qps = B / (ms_per_step * 1e-3)
    result_ids = out_i.cpu().numpy()
    result_d = out_d.cpu().numpy()
    scan_ms = float(np.median([t["scan_ms"] for t in stage_ms]))
    peak, peak_src = measured_peak_hbm()

    # ---- k = 10 beside the config's k ("report k=10 too"), single GPU
    k10 = None
    if world == 1 and k != 10 and not args.no_k10:
        od10 = torch.empty((B, 10), dtype=torch.float32, device=dev)
        oi10 = torch.empty((B, 10), dtype=torch.int64, device=dev)
        for _ in range(3):
            index.search_dev(q_dev.data_ptr(), B, 10, 1, od10.data_ptr(), oi10.data_ptr(), 0, 0)
        t10 = []
        for _ in range(max(3, args.steps // 2)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            index.search_dev(q_dev.data_ptr(), B, 10, 1, od10.data_ptr(), oi10.data_ptr(), 0, 0)
            b.record()
            b.synchronize()
            t10.append(a.elapsed_time(b))
        same = bool(np.array_equal(oi10.cpu().numpy(), result_ids[:, :10]))
        k10 = {"value": B / (float(np.median(t10)) * 1e-3), "unit": "queries/s", "ms_per_step": float(np.median(t10)), "top10_equals_head_of_top_k": same}

    # ---- the plain row-sharded layout (row shards = N, queries replicated) beside the planned grid
    row_sharded = None
    if world > 1 and R != world:
        index_rs, lo_rs, hi_rs = make_layout(world)
        keep_d, keep_i = result_d.copy(), result_ids.copy()
        for _ in range(args.warmup):
            step(index_rs, world, lo_rs)
        per_rs, stages_rs = timed_steps(index_rs, world, lo_rs, args.steps)
        assert np.array_equal(out_i.cpu().numpy(), keep_i) and np.array_equal(out_d.cpu().numpy().view(np.uint32), keep_d.view(np.uint32)), \
            "row-sharded layout and planned grid disagree"
        # e2e of this layout: the whole batch goes to every rank, rank 0 reads the result
        def e2e_rs():
            qd = q_pinned.to(dev, non_blocking=True)
            step(index_rs, world, lo_rs, qd)
            if rank == 0:
                out_d_host.copy_(out_d, non_blocking=True)
                out_i_host.copy_(out_i, non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(2):
            e2e_rs()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_rs()
        barrier()
        e_rs = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e_rs, op=dist.ReduceOp.MAX)
        scan_rs = float(np.median([t["scan_ms"] for t in stages_rs]))
        ms_rs = float(np.median(per_rs))
        row_sharded = {"value": B / (ms_rs * 1e-3), "unit": "queries/s", "ms_per_step": ms_rs, "rows_per_gpu": hi_rs - lo_rs,
                       "scan_kernel_ms": scan_rs, "roofline_frac": float(B) * (hi_rs - lo_rs) * M / (scan_rs * 1e-3) / 1e9 / peak,
                       "e2e": {"value": B / float(e_rs.item()), "unit": "queries/s"}}
        index_rs.close()

    # ---- end-to-end through the public C-ABI call with HOST buffers (H2D + D2H inside the timed region).
    # N = 1: b200nn_pq_search, host pointers in and out.  N > 1: every rank copies in only the queries it scans (its chunk of
    # the batch; the whole batch when the rows are sharded N ways), rank 0 copies the merged result out.
    def e2e_step():
        if world == 1:
            index.search_host_ptr(q_pinned.data_ptr(), B, k, 1, out_d_host.data_ptr(), out_i_host.data_ptr())
        else:
            q_dev[q_lo:q_hi].copy_(q_pinned[q_lo:q_hi], non_blocking=True)
            step(index, R, lo)
            if rank == 0:
                out_d_host.copy_(out_d, non_blocking=True)
                out_i_host.copy_(out_i, non_blocking=True)
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
    if rank == 0:
        assert np.array_equal(out_i_host.numpy(), result_ids), "e2e path and device path disagree"

    # ---- parity on the FULL database: numpy restatement of the scan for two queries, every rank over its own rows,
    # the per-rank lists merged on rank 0 (skipped with --no-cpu-baseline)
    full_parity = None
    if not args.no_full_parity:
        from oracle import oracle as orc  # checker only
        qsel = [0, B - 1]
        _, _, codes_h = index.get_rows()
        qr = orc.opq_reorder(q[qsel], perm)
        mine = []
        for j in range(len(qsel)):
            lut = orc.opq_build_lut(qr[j], coarse[0], cb)
            mine.append(numpy_adc_topk(lut, codes_h, k, 1.0, lo))
        del codes_h
        shard_r, chunk_c = sharded.grid_coords(rank, R)
        gathered = [None] * world
        if world > 1:
            dist.all_gather_object(gathered, (shard_r, chunk_c, mine))
        else:
            gathered = [(0, 0, mine)]
        if rank == 0:
            ok_i, ok_d = True, True
            for j, qi in enumerate(qsel):
                parts = [g[2][j] for g in gathered if g[1] == 0]  # one copy of every row shard
                s_all = np.concatenate([p[0] for p in parts]); i_all = np.concatenate([p[1] for p in parts])
                order = np.lexsort((i_all, s_all))[:k]
                ok_i &= bool(np.array_equal(i_all[order], result_ids[qi]))
                ok_d &= bool(np.array_equal(s_all[order].view(np.uint32), result_d[qi].view(np.uint32)))
            full_parity = {"vs": "numpy restatement of the full scan (sequential fp32 LUT sums, clamp, (score, id) order)", "queries_checked": len(qsel),
                           "rows": n, "topk_ids_identical": ok_i, "scores_bit_identical": ok_d}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    alg_bytes = float(q_hi - q_lo) * (hi - lo) * M  # every query of this rank's chunk "reads" every code byte of its shard once
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    traffic = None  # dram bytes of ONE launch of this shape from an ncu --set full capture (profiles/scan_traffic.json), else null
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            for e in json.load(open(tp)).get("captures", []):
                if e.get("M") == M and e.get("rows") == hi - lo and e.get("queries") == q_hi - q_lo and e.get("k") == k:
                    traffic = e.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    below_clamp = float(np.mean(result_d[:, k - 1] < np.float32(1.0)))
    line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "ms_per_step_mean": float(np.mean(per_step)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (LUT sums) over u8 codes", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_rows": n, "rows_per_gpu": hi - lo, "queries_per_gpu": q_hi - q_lo, "dim": D, "M": M, "ksub": 256,
                       "batch": B, "k": k, "nprobe": 1, "clamp": 1.0,
                       "parallelism": (f"{R} row shard(s) x {Qc} query chunk(s) ({'planned' if not args.row_shards else 'forced'}), one ncclAllGather of top-k keys "
                                       "issued by the C library" if world > 1 else "single GPU"),
                       "rows": ("numpy, host" if wl["gen"] == "host" else f"generated on the device in seeded chunks of {GEN_CHUNK} rows, encoded chunk by chunk (codes only stay)"),
                       "codebooks": f"seeded Lloyd, {TRAIN_ITERS} iterations, first {min(n, TRAIN_ROWS)} rows (SURVEY.md 8(d))",
                       "timing": "per-step CUDA events on the launching stream, max over ranks per step, median over steps",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "extra_untimed_warmup_steps_for_clock_sampling": extra_warm,
                       "seeds": "SURVEY.md 8(d)", "index_build_s": build_s},
            "roofline": {"bound": "hbm", "kernel": "adc_scan_topk_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_ms, "scan_launches_per_step": 1,
                         "note": "algorithmic bytes = batch x shard rows x M; binding unit is the shared-memory gather pipe (DESIGN.md)"},
            "stage_ms": {kk: float(np.median([t[kk] for t in stage_ms])) for kk in stage_ms[0]},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int((q_hi - q_lo) * D * 4), "d2h_bytes_per_step": int(B * k * 12),
                    "note": "per rank H2D of the queries it scans; D2H of the merged [batch, k] result on rank 0" if world > 1 else "b200nn_pq_search with host buffers"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
            "sanity": {"kth_score_below_clamp_frac": below_clamp, "gate": ">0.99 (SURVEY.md 8(d))", "ok": bool(below_clamp > 0.99)}}
    if not line["sanity"]["ok"]:
        line["sanity"]["warning"] = "the k-th best score sits at the clamp for more than 1 % of the queries: id-ordered ties dominate"
    if k10 is not None:
        line["k10"] = k10
    if row_sharded is not None:
        line["row_sharded"] = row_sharded
    if full_parity is not None:
        line["parity_full_scan"] = full_parity

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on a bounded sample, + parity
    if world == 1 and not args.no_cpu_baseline:
        try:
            ns = min(B, args.cpu_sample)
            r = run_reference_sample(wl, inp, ns, repeat=1)
            cpu_qps = ns / r["query_s_mean"] * (r["rows"] / n)
            if r["rows"] == n:
                ours_d, ours_i = result_d[:ns], result_ids[:ns]
            else:  # the reference scanned a row subsample: the same subsample through our path
                sub = capi.PQIndex.create(ctx, coarse, cb, perm=perm, clamp=1.0)
                sub.add(host_head(wl, inp))
                ours_d, ours_i = sub.search(q[:ns], k)
                ours_i = ours_i.astype(np.int64)
                sub.close()
            ids_equal = bool(np.array_equal(r["topk_id"], ours_i))
            rel = np.abs(r["topk_score"] - ours_d) / np.maximum(np.abs(r["topk_score"]), 1e-30)
            recall10 = float(np.mean([len(set(r["topk_id"][i, :10]) & set(ours_i[i, :10])) / 10.0 for i in range(ns)]))
            line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": r["threads"], "kind": r["kind"], "host": host_cpu_model(),
                                    "sample": cpu_sample_text(wl, r, ns)}
            line["parity"] = {"vs": "unmodified reference (oracle/_ref/ref_opq)" if r["kind"] == "reference" else "oracle restatement (oracle/cvt_oracle.c)",
                              "queries_checked": ns, "rows": r["rows"], "topk_ids_identical": ids_equal,
                              "recall_at_10": recall10, "max_rel_dist_err": float(rel.max()),
                              "scores_bit_identical": bool(np.array_equal(r["topk_score"].view(np.uint32), ours_d.view(np.uint32)))}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    else:
        line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference",
                                "sample": "timed at N=1 only (see the N=1 line)"}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
