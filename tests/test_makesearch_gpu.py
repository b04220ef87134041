"""hnsw_sifts_retrieval/makeSearch.cpp -- the other caller north_star names -- as a drop-in: the reference's makeSearch.cpp +
siftsIndex.cpp are compiled twice, unmodified apart from the hard-coded /Users/willard/... path prefixes (-> data/):
  oracle/_ref/ref_makeSearch        with the reference's own CPU hnswlib (HierarchicalNSW, setEf(1000))
  tools/bin/ref_makeSearch_on_b200nn with include/b200nn/compat (-DB200NN_HNSW_DROP_IN: HierarchicalNSW = the exact GPU index,
                                     which reads the SAME HNSW index file, hnswalg.h:491-519)
Both run in the same data directory -- an index the reference's hnswlib built (oracle/_ref/ref_hnsw_build, as makeIdx.cpp
does), the geoInfo file of makeIdx.cpp:366-393, REAL SIFT descriptors of the shipped image (tests/golden/makesearch_china2.sift,
from cv2) -- and must print the same matches: same labels, same 1 - <a,b> distances to the printed digits, same ranking.
OpenCV itself is absent: tests/stubs/opencv2 stands in (see its header)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import cases
from cvt_b200 import synth
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_makeSearch")
REF_BUILD = os.path.join(ROOT, "oracle", "_ref", "ref_hnsw_build")
GPU_EXE = os.path.join(ROOT, "tools", "bin", "ref_makeSearch_on_b200nn")
SIFT = os.path.join(cases.GOLDEN, "makesearch_china2.sift")


def read_sift(path):
    b = open(path, "rb").read()
    n = struct.unpack_from("<i", b, 0)[0]
    kp = np.frombuffer(b, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("angle", "<f4"), ("size", "<f4"), ("resp", "<f4"), ("cls", "<i4"), ("oct", "<i4")]),
                       count=n, offset=4)
    desc = np.frombuffer(b, dtype="<f4", count=n * 128, offset=4 + n * 28).reshape(n, 128)
    return kp, desc


def make_data_dir(tmp, n_rows=3000):
    """data/: the query image's .sift, an HNSW index over n_rows rootSIFT rows (the first rows ARE the image's own
    descriptors, filed under three template names; the rest are synthetic), and the matching geoInfo file."""
    d = os.path.join(tmp, "data")
    os.makedirs(d)
    shutil.copy(SIFT, os.path.join(d, "201505310117china2.jpg.sift"))
    kp, desc = read_sift(SIFT)
    n_img = len(kp)
    rows = np.concatenate([orc.rootsift(desc), synth.sift_like(n_rows - n_img, 128, seed=0x51F7A)]).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(77))
    rows_path = os.path.join(tmp, "rows.f32")
    rows.tofile(rows_path)
    index = os.path.join(d, "sifts_125402m_ef_80_M_32_ip.bin")  # the file name makeSearch.cpp:19 hard-codes
    r = subprocess.run([REF_BUILD, rows_path, str(n_rows), "128", "32", "80", index], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    with open(os.path.join(d, "sifts_125402_geoInfo.bin"), "wb") as f:  # makeIdx.cpp:366-393 / siftsIndex.cpp:13-48
        f.write(struct.pack("<i", n_rows))
        for i in range(n_rows):
            if i < n_img:
                name = f"template_{i % 3}.jpg".encode()
                ang = float(kp["angle"][i]) + (0.0 if i % 4 else 25.0)  # every 4th match fails the 10-degree angle test
                pt = (float(kp["x"][i]), float(kp["y"][i]), ang, float(kp["size"][i]), float(kp["resp"][i]), int(kp["cls"][i]), int(kp["oct"][i]))
            else:
                name = f"other_{i % 17}.jpg".encode()
                pt = (float(rng.random() * 640), float(rng.random() * 480), float(rng.random() * 360), 3.0, 0.05, -1, 0)
            f.write(struct.pack("<i", len(name)) + name + struct.pack("<ii", i, i) + struct.pack("<fffffii", *pt))
    return rows, index


needs_ref = pytest.mark.skipif(not (os.path.exists(REF_EXE) and os.path.exists(REF_BUILD)), reason="oracle/_ref not built")


@needs_ref
def test_reference_makesearch_runs_on_cpu(tmp_path):
    """not gpu: the reference build (CPU hnswlib + the OpenCV stand-in) runs, finds the image's own descriptors and ranks the
    three template files first -- guards the fixture and the stand-in."""
    make_data_dir(str(tmp_path))
    r = subprocess.run([REF_EXE], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert sum(ln.startswith("template_") and "angle diff" in ln for ln in lines) >= 64
    ranked = [ln.split(" :: ")[0] for ln in lines if " :: " in ln]
    assert ranked and ranked[0].startswith("template_")


@needs_ref
@pytest.mark.gpu
def test_makesearch_dropin_equals_reference(tmp_path):
    if not os.path.exists(GPU_EXE):
        pytest.skip("tools/bin/ref_makeSearch_on_b200nn not built")
    assert b"b200nn_flat_load_hnsw" in open(GPU_EXE, "rb").read()
    rows, index = make_data_dir(str(tmp_path))
    ref = subprocess.run([REF_EXE], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    gpu = subprocess.run([GPU_EXE], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0 and gpu.returncode == 0, ref.stderr + gpu.stderr
    strip = lambda out: [ln for ln in out.splitlines() if not ln.startswith("Loading index")]
    assert strip(gpu.stdout) == strip(ref.stdout)
    # and the HNSW file reader hands the exact index the same rows: exact top-5 of every descriptor == brute force in numpy
    from cvt_b200 import capi
    ctx = capi.Context(0)
    idx = capi.FlatIndex.load_hnsw_file(ctx, "ip", 128, index)
    mx, n, labels = idx.info()
    assert (mx, n) == (len(rows), len(rows)) and np.array_equal(labels, np.arange(len(rows), dtype=np.uint64))
    _, desc = read_sift(SIFT)
    q = orc.rootsift(desc)
    D, L = idx.search(q, 5)
    od, ol = orc.flat_search(0, 4, rows, np.arange(len(rows), dtype=np.uint64), q, 5)
    assert np.array_equal(L, ol) and np.array_equal(D.view(np.uint32), od.view(np.uint32))
    idx.close(); ctx.close()
