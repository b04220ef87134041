"""GPU tests of the row-sharded index behind the C ABI (include/b200nn.h, "several GPUs"): the sharded answers must be
IDENTICAL to a single index's -- ids, fp32 score bits, group scores, index files.

On a one-GPU box the shards are several contexts on the same device ($B200NN_ALLOW_DUPLICATE_DEVICES, peer-memory
exchange); with >= 2 GPUs the same tests also run across real devices, with the peer-memory merge and with the
ncclAllGather exchange, and the one-process-per-GPU path (b200nn_comm_* + pq_search_sharded_dev) runs under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _layouts():
    """(devices, exchange) combinations available on this box."""
    out = [([0, 0, 0], "p2p")]
    n = _n_gpus()
    if n >= 2:
        devs = list(range(min(n, 8)))
        out += [(devs, "p2p"), (devs, "nccl")]
    return out


@pytest.fixture()
def dup_env(monkeypatch):
    monkeypatch.setenv("B200NN_ALLOW_DUPLICATE_DEVICES", "1")
    yield monkeypatch


@pytest.mark.parametrize("case,nprobe,k", [("flat_m16", 1, 100), ("ivf_m8", 3, 50), ("flat_m16_clamp", 1, 50)])
def test_mpq_search_equals_single_index(dup_env, case, nprobe, k):
    from cvt_b200 import capi
    c = cases.opq_case(case)
    ctx = capi.Context(0)
    single = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"], clamp=1.0)
    single.add(c["db"])
    q = np.concatenate([c["q"], c["q"] * np.float32(3.0)])  # scaled copies: scores over the clamp -> id-ordered tails
    D0, I0 = single.search(q, k, nprobe=nprobe)
    for devices, exchange in _layouts():
        dup_env.setenv("B200NN_EXCHANGE", exchange)
        m = capi.MultiPQ.create(devices, c["coarse"], c["cb"], perm=c["reorder"], clamp=1.0)
        assert m.peer_exchange == (exchange == "p2p")
        # three add calls of different sizes: every call is dealt to the shards in contiguous blocks
        cuts = [0, 7, c["n"] // 3, c["n"]]
        for a, b in zip(cuts[:-1], cuts[1:]):
            m.add(c["db"][a:b])
        assert m.n_rows == c["n"] and int(m.shard_rows().sum()) == c["n"] and m.shard_rows().min() > 0
        for _ in range(2):  # second call reuses the exchange buffers
            D1, I1 = m.search(q, k, nprobe=nprobe)
            assert np.array_equal(I1, I0), (devices, exchange)
            assert np.array_equal(_bits(D1), _bits(D0)), (devices, exchange)
        D2, I2 = m.search(q[:5], 7, nprobe=nprobe)  # ragged: fewer queries than devices x 2, other k
        Ds, Is = single.search(q[:5], 7, nprobe=nprobe)
        assert np.array_equal(I2, Is) and np.array_equal(_bits(D2), _bits(Ds))
        m.close()
    single.close()
    ctx.close()


def test_mpq_scores_and_index_files_equal_single_index(dup_env, tmp_path):
    """IVFOPQ::QueryThrehold semantics (min per videoId over the probed lists) across shards = elementwise min of the
    shards' matrices; SaveIndex of the sharded index is byte-identical to the single index's; LoadIndex onto shards."""
    from cvt_b200 import capi
    c = cases.opq_case("ivf_m8")
    groups = (np.arange(c["n"]) // 50).astype(np.int32)
    ctx = capi.Context(0)
    single = capi.PQIndex.create(ctx, c["coarse"], c["cb"], perm=c["reorder"])
    single.add(c["db"], groups)
    S0 = single.scores(c["q"], nprobe=3)
    paths = [f"/data/video_{i}.bin" for i in range(single.n_groups)]
    os.makedirs(tmp_path / "single")
    single.save_index(str(tmp_path / "single"), paths)
    name = f"OPQ_Index_db_{single.n_groups}_dim_64_k_16_PQ_m8_k256.fvecs"
    ref_bytes = (tmp_path / "single" / name).read_bytes()
    for li, (devices, exchange) in enumerate(_layouts()):
        dup_env.setenv("B200NN_EXCHANGE", exchange)
        m = capi.MultiPQ.create(devices, c["coarse"], c["cb"], perm=c["reorder"])
        for a, b in ((0, 333), (333, c["n"])):
            m.add(c["db"][a:b], groups[a:b])
        assert m.n_groups == single.n_groups
        assert np.array_equal(_bits(m.scores(c["q"], nprobe=3)), _bits(S0)), (devices, exchange)
        d = tmp_path / f"multi{li}"
        os.makedirs(d)
        m.save_index(str(d), paths)
        assert (d / name).read_bytes() == ref_bytes
        m.close()
        m2 = capi.MultiPQ.load_index(devices, str(d / name), perm=c["reorder"])
        assert m2.n_rows == c["n"] and m2.n_groups == single.n_groups
        assert np.array_equal(_bits(m2.scores(c["q"], nprobe=3)), _bits(S0))
        m2.close()
    single.close()
    ctx.close()


def test_two_contexts_in_one_process(dup_env):
    """per-device kernel attributes must not be cached process-wide -- two contexts (here on the same GPU; on a
    multi-GPU box also on different ones) run the fused scan in one process."""
    from cvt_b200 import capi
    c = cases.opq_case("flat_m16")
    devs = [0, 1] if _n_gpus() >= 2 else [0, 0]
    res = []
    ctxs = [capi.Context(d) for d in devs]
    idxs = []
    for cx in ctxs:
        ix = capi.PQIndex.create(cx, c["coarse"], c["cb"], perm=c["reorder"])
        ix.add(c["db"])
        idxs.append(ix)
    for ix in idxs + idxs[::-1]:
        res.append(ix.search(c["q"], 100))
    for D, I in res[1:]:
        assert np.array_equal(I, res[0][1]) and np.array_equal(_bits(D), _bits(res[0][0]))
    for ix in idxs:
        ix.close()
    for cx in ctxs:
        cx.close()


def test_opq_cli_sharded_equals_single(dup_env, tmp_path):
    """The C++ IVFOPQ drop-in (tools/opq_cli: the reference's two mains) over $B200NN_DEVICES answers exactly as on one GPU:
    index file byte-identical, query result file identical."""
    exe = os.path.join(ROOT, "tools", "bin", "opq_cli")
    if not os.path.exists(exe):
        pytest.skip("tools/bin/opq_cli not built")
    G = cases.GOLDEN
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(os.path.join(G, "opq_fixture", "db", f) for f in cases.FIXTURE_DB) + "\n")
    model = os.path.join(G, "opq_shipped_k256.model")
    qs = [os.path.join(G, "opq_fixture", "query", f) for f in cases.FIXTURE_QUERY]
    outs = {}
    n = _n_gpus()
    variants = {"single": None, "dup3": "0,0,0"}
    if n >= 2:
        variants[f"gpus{min(n, 8)}"] = ",".join(str(i) for i in range(min(n, 8)))
    for tag, devs in variants.items():
        d = tmp_path / tag
        os.makedirs(d)
        env = dict(os.environ, B200NN_ALLOW_DUPLICATE_DEVICES="1")
        env.pop("B200NN_DEVICES", None)
        if devs:
            env["B200NN_DEVICES"] = devs
        r = subprocess.run([exe, "index", model, str(lst), str(d)], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr + r.stdout
        idx = d / "OPQ_Index_db_5_dim_128_k_256_PQ_m16_k256.fvecs"
        res = d / "result.txt"
        r = subprocess.run([exe, "query", model, str(idx), str(res), *qs], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr + r.stdout
        outs[tag] = (idx.read_bytes(), res.read_text())
    for tag in variants:
        assert outs[tag][0] == outs["single"][0], tag
        assert outs[tag][1] == outs["single"][1], tag
    # and the single-GPU answer is the reference's golden one (tests/test_cli_gpu.py checks the details)
    gold = np.load(os.path.join(G, "opq_shipped_k256.npz"))
    first = outs["single"][1].split("\n\n")[0].strip().split("\n")
    assert np.array_equal(np.array(first[2].split(), dtype=np.float32), gold["file_topk_score"][0])


def test_multi_process_sharded_search_under_torchrun(tmp_path):
    """One process per GPU: b200nn_comm_* (ncclCommInitRank) + b200nn_pq_search_sharded_dev, launched with torchrun as
    bench.py is; tools/sharded_check.py compares every grid (row shards x query chunks) with the single index."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded_check ok" in r.stdout
