"""CPU tests of the drop-in boundary: the built C-ABI library loads and exports every symbol that
include/b200nn.h declares (no compute calls -- there is no GPU here), and it refuses to run
without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cvt_b200 import build, capi
    build.build_lib()
    return capi.load()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nn_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported(lib):
    from cvt_b200 import capi
    declared = _declared()
    assert len(declared) >= 40
    assert sorted(capi.SYMBOLS) == declared, "cvt_b200/capi.py SYMBOLS must list exactly what the header declares"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/b200nn.h but not exported"


def test_library_is_sm100a_native_code():
    from cvt_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(5|6|7|8|9)\d", out)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cvt_b200 import capi
    with pytest.raises(capi.B200nnError, match="no CUDA device"):
        capi.Context(0)
    assert b"sm_100a" in lib.b200nn_version()


def test_product_does_not_touch_oracle():
    # the oracle is test infrastructure: nothing under cvt_b200/, include/ or tools/ may reference it
    for base in ("cvt_b200", "include", "tools"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "/lib" in dp or "/bin" in dp:
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle" not in txt.lower() or f == "quick_scan_bench.py", f"{dp}/{f} mentions the oracle"


@pytest.mark.parametrize("M", [4, 8, 16, 32])
def test_scan_plan_tiles_the_work_exactly(lib, M):
    """Host-only: the fused scan's work plan (whole-shard CTAs + equal pieces of the last wave) covers every
    (query group, granule) exactly once, numbers a group's output slices 0..c-1 and balances the tail."""
    import numpy as np
    from cvt_b200 import capi
    qw = 128 // M
    for sm in (148, 132, 7):
        for nq, n_rows in [(4096, 1_000_000), (512, 1_000_000), (1024, 500_000), (4096, 125_000), (7, 100), (100, 0),
                           (sm * qw, 20_000), (sm * qw + 1, 20_000), (5000, 53), (801, 20_000), (3, 1_000_003)]:
            n_full, n_tail, slices, desc = capi.scan_plan(sm, M, nq, n_rows)
            groups, gran = -(-nq // qw), -(-n_rows // 64)
            assert n_full == (groups // sm) * sm and n_full + (1 if n_tail else 0) * 1 <= groups + n_tail
            rem = groups - n_full
            assert (n_tail == 0) == (rem == 0) and n_tail <= sm and (rem == 0 or n_tail >= rem)
            cover = {g: [] for g in range(n_full, groups)}
            sizes = []
            for t in range(n_tail):
                size = 0
                for s in range(2):
                    g, sl, lo, hi = (int(v) for v in desc[t, s])
                    if s == 1 and lo >= hi:
                        continue
                    assert n_full <= g < groups and 0 <= lo <= hi <= gran and 0 <= sl < slices
                    cover[g].append((sl, lo, hi))
                    size += hi - lo
                sizes.append(size)
            for g, segs in cover.items():
                segs.sort()
                assert [s[0] for s in segs] == list(range(len(segs))), "slices of a group are numbered 0..c-1"
                assert segs[0][1] == 0 and segs[-1][2] == gran and all(a[2] == b[1] for a, b in zip(segs, segs[1:]))
            if n_tail:
                assert max(sizes) - min(sizes) <= 1, "tail pieces are equal to within one granule"
                assert slices == max(len(v) for v in cover.values())
            else:
                assert slices == 1


def test_pca_yaml_model_reader(lib, tmp_path):
    """Host-only: PCAUtils::loadModel's file format (cv::PCA written by cv::FileStorage, pca_utils.cc:16-23).  The fixture
    was written AND read back by cv2 in the build container (oracle/gen_golden_frontend.py); the reader must return the
    same bits.  Where the reference tree is present, both shipped models are compared with cv2 as well."""
    import numpy as np
    from cvt_b200 import capi
    G = os.path.join(ROOT, "tests", "golden")
    gold = np.load(os.path.join(G, "pca_small_64x64.npz"))
    mean, vectors, values = capi.pca_read_model(os.path.join(G, "pca_small_64x64.yml"))
    assert vectors.shape == (64, 64) and mean.shape == (64,) and values.shape == (64,)
    for got, want in ((mean, gold["mean"]), (vectors, gold["vectors"]), (values, gold["values"])):
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want, dtype=np.float32).view(np.uint32))
    with pytest.raises(capi.B200nnError):
        capi.pca_read_model(str(tmp_path / "missing.yml"))
    bad = tmp_path / "bad.yml"
    bad.write_text("%YAML:1.0\n---\nname: PCA\nvectors: !!opencv-matrix\n   rows: 2\n   cols: 2\n   dt: f\n   data: [ 1., 2., 3. ]\n"
                   "mean: !!opencv-matrix\n   rows: 1\n   cols: 2\n   dt: f\n   data: [ 0., 0. ]\n")
    with pytest.raises(capi.B200nnError):
        capi.pca_read_model(str(bad))  # 3 numbers for a 2 x 2 matrix
    dbl = tmp_path / "dbl.yml"    # CV_64F nodes and no eigenvalues: accepted, narrowed to float, values = 0
    dbl.write_text("%YAML:1.0\n---\nvectors: !!opencv-matrix\n   rows: 1\n   cols: 2\n   dt: d\n   data: [ 0.1, -2.5e-1 ]\n"
                   "mean: !!opencv-matrix\n   rows: 1\n   cols: 2\n   dt: d\n   data: [ 1., 3. ]\n")
    m2, v2, e2 = capi.pca_read_model(str(dbl))
    assert np.array_equal(v2, np.array([[0.1, -0.25]], np.float32)) and np.array_equal(m2, [1.0, 3.0]) and not e2.any()
    ref = "/root/reference/pca_train_project/model/pca_1024_128_300w_googlenet.yml"
    if os.path.exists(ref):
        cv2 = pytest.importorskip("cv2")
        fs = cv2.FileStorage(ref, cv2.FILE_STORAGE_READ)
        mean, vectors, values = capi.pca_read_model(ref)
        assert np.array_equal(vectors.view(np.uint32), fs.getNode("vectors").mat().view(np.uint32))
        assert np.array_equal(mean, fs.getNode("mean").mat().reshape(-1)) and np.array_equal(values, fs.getNode("values").mat().reshape(-1))


@pytest.mark.parametrize("src,std,extra", [
    ("brute_force_search/src/brute_force.cpp", "c++11", ["-fno-operator-names"]),
    ("opq/src/multi_frame_index_test.cpp", "c++11", []),                      # the `#if 1` main (index build)
    ("opq/src/multi_frame_index_test.cpp", "c++11", ["-DB200NN_FLIP_MAINS"]),  # the `#if 0` main (query), see below
    ("opq/train_codebook/train_PQ.cpp", "c++11", []),
    ("scalar_quantization/scalar_quantization/int8_quan_test.cpp", "c++17", []),
])
def test_reference_callers_compile_unmodified_against_the_dropin_headers(src, std, extra, tmp_path):
    """INTEGRATION.md's claim, kept honest: the reference's own caller sources compile, unmodified, against
    include/b200nn/compat (source piped through stdin so that quoted includes cannot pick up the reference's own headers
    lying next to the file).  Only where the reference tree is present (the build container)."""
    path = os.path.join("/root/reference", src)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    text = open(path, "rb").read()
    if "-DB200NN_FLIP_MAINS" in extra:  # select the file's second main the way its author does: by flipping the two #if lines
        text = text.replace(b"#if 1", b"#if 0_", 1).replace(b"#if 0\n", b"#if 1\n", 1).replace(b"#if 0_", b"#if 0", 1)
        extra = []
    cmd = ["g++", "-std=" + std, "-fsyntax-only", "-I", os.path.join(ROOT, "include", "b200nn", "compat"), "-x", "c++", "-"] + extra
    r = subprocess.run(cmd, input=text, capture_output=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr.decode(errors="replace")[:2000]
