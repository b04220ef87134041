"""Drop-in tests at the C++ level: the reference's OWN brute-force CLI source, compiled unmodified
against include/b200nn/compat and linked with libb200nn, must write the same gt.txt / index.bin as
the reference CLI built natively (golden, produced in the build container); and the C++ tools
written against the drop-in classes reproduce the reference's results on the shipped fixtures."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
G = cases.GOLDEN
BIN = os.path.join(cases.ROOT, "tools", "bin")


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not built (python -m cvt_b200.build)")
    return p


def _records(td):
    c = cases.cli_case()
    cases.write_record_file(os.path.join(td, "db.bin"), c["db_ids"], c["db"])
    cases.write_record_file(os.path.join(td, "querys.bin"), c["q_ids"], c["q"])


def test_unmodified_reference_cli_on_gpu_index(tmp_path):
    exe = _need("ref_brute_force_on_b200nn")
    # the binary must really bind the C ABI (a quoted #include resolves next to the reference source first: the build
    # pipes the source through stdin for that reason and checks the same thing)
    blob = open(exe, "rb").read()
    assert b"b200nn_flat_search" in blob and b"b200nn_flat_add" in blob
    _records(str(tmp_path))
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "gt.txt").read() == open(os.path.join(G, "cli_brute_force_gt.txt")).read()
    dig = hashlib.sha256(open(tmp_path / "index.bin", "rb").read()).hexdigest()
    assert dig == open(os.path.join(G, "cli_brute_force_index.sha256")).read().strip()


def test_brute_search_tool(tmp_path):
    exe = _need("brute_search")
    _records(str(tmp_path))
    r = subprocess.run([exe, "db.bin", "querys.bin", "index.bin", "gt.txt", "100", "128"], cwd=str(tmp_path), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "gt.txt").read() == open(os.path.join(G, "cli_brute_force_gt.txt")).read()
    dig = hashlib.sha256(open(tmp_path / "index.bin", "rb").read()).hexdigest()
    assert dig == open(os.path.join(G, "cli_brute_force_index.sha256")).read().strip()


def test_opq_cli_index_then_query(tmp_path):
    """multi_frame_index_test.cpp's two mains (index build, then query + frame-summed top-5)."""
    exe = _need("opq_cli")
    gold = np.load(os.path.join(G, "opq_shipped_k256.npz"))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(os.path.join(G, "opq_fixture", "db", f) for f in cases.FIXTURE_DB) + "\n")
    model = os.path.join(G, "opq_shipped_k256.model")
    r = subprocess.run([exe, "index", model, str(lst), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    idx = tmp_path / "OPQ_Index_db_5_dim_128_k_256_PQ_m16_k256.fvecs"  # the reference's file-name scheme
    assert idx.exists()
    qs = [os.path.join(G, "opq_fixture", "query", f) for f in cases.FIXTURE_QUERY]
    res = tmp_path / "result.txt"
    r = subprocess.run([exe, "query", model, str(idx), str(res), *qs], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    blocks = [b for b in res.read_text().split("\n\n") if b.strip()]
    assert len(blocks) == 2
    base = [f[:-4] for f in cases.FIXTURE_DB]  # get_base_name strips the directory and extension
    for bi, b in enumerate(blocks):
        lines = b.strip().split("\n")
        names = lines[1].split()
        scores = np.array(lines[2].split(), dtype=np.float32)
        assert names == [base[i] for i in gold["file_topk_id"][bi]]
        assert np.array_equal(scores, gold["file_topk_score"][bi])


def test_unmodified_reference_int8_quan_test_on_gpu(tmp_path):
    """scalar_quantization/scalar_quantization/int8_quan_test.cpp, compiled unmodified against include/b200nn/compat:
    loads model/int8_siamese_photo_embedding_8kw.bin (an IxSQ file, written here in the layout of SURVEY.md App. A-8),
    encodes its pinned 64-d vector with Int8Encode (L2 normalisation on), decodes it with Int8Decode(std::string&) and
    prints both.  Golden = the stdout of THE SAME main built from the reference's own int8_quan.cc (oracle/_ref/
    ref_int8_quan_test, stand-in faiss headers) on the same model file: the two must be byte-identical."""
    exe = _need("ref_int8_quan_test_on_b200nn")
    assert b"b200nn_sq_encode" in open(exe, "rb").read()
    vmin, vdiff = cases.int8_quan_test_model()
    os.makedirs(tmp_path / "model")
    cases.write_ixsq_file(str(tmp_path / "model" / "int8_siamese_photo_embedding_8kw.bin"), vmin, vdiff)
    r = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    gold = open(os.path.join(G, "int8_quan_test_stdout.bin"), "rb").read()
    assert r.stdout == gold
    # and the numbers in it are what the restatement says (guards the golden itself)
    from oracle import oracle as orc
    lines = r.stdout.split(b"\n")
    k_bytes = next(i for i, ln in enumerate(lines) if ln.startswith("int8压缩后表示".encode()))
    got_bytes = np.array(lines[k_bytes + 1].split(), dtype=np.int64)
    codes, _ = orc.sq_encode(cases.int8_quan_test_vector()[None, :], vmin, vdiff, l2norm=True)
    assert np.array_equal(got_bytes, codes[0].astype(np.int64))
