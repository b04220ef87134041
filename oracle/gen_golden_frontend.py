"""Golden vectors for the front end, generated HERE with the dependency the
reference calls -- OpenCV (Python cv2 4.13; the reference builds against OpenCV 3.4 C++) -- on the
reference's shipped PCA model and on seeded inputs; aborts unless the C restatement agrees within a few ulp.

    python oracle/gen_golden_frontend.py          # writes tests/golden/frontend_pca.npz, frontend_rootsift.npz

reduceDim  (pca_train_project/pca_online/pca_utils.cc:25-35) = cv::PCA::project + per-row L2 normalisation
rootSift   (hnsw_sifts_retrieval/siftsIndex.cpp:54-71)       = abs, L1-normalise, sqrt, cv::normalize(NORM_L2)
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import oracle as orc  # noqa: E402

MODEL = "/root/reference/pca_train_project/model/pca_1024_128_300w_googlenet.yml"
N_KEEP = 64  # eigenvectors kept in the committed fixture (the full model is 2.4 MB of text)


def reduce_dim_cv2(x, mean, vectors):
    """pca_utils.cc:25-35 with cv2 calls: pca_.project, then row * row.t(), max(1e-12, sqrt), divide."""
    y = cv2.PCAProject(x, mean.reshape(1, -1), vectors)
    out = np.empty_like(y)
    for i in range(y.shape[0]):
        row = y[i:i + 1]
        norm_mat = cv2.gemm(row, row, 1.0, None, 0.0, flags=cv2.GEMM_2_T)  # row * row.t()
        denomv = np.float32(max(1e-12, float(np.sqrt(np.float32(norm_mat[0, 0])))))
        out[i] = row[0] / denomv
    return out


def root_sift_cv2(d, eps=np.float32(1e-7)):
    d = np.abs(d).astype(np.float32)
    sums = cv2.reduce(d, 1, cv2.REDUCE_SUM, dtype=cv2.CV_32F)
    out = np.empty_like(d)
    for r in range(d.shape[0]):
        row = np.sqrt(d[r] / (sums[r, 0] + eps)).astype(np.float32)
        out[r] = cv2.normalize(row.reshape(1, -1), None, 1.0, 0.0, cv2.NORM_L2)[0]
    return out


def main():
    fs = cv2.FileStorage(MODEL, cv2.FILE_STORAGE_READ)
    vectors = fs.getNode("vectors").mat().astype(np.float32)
    mean = fs.getNode("mean").mat().astype(np.float32).reshape(-1)
    assert vectors.shape == (128, 1024) and mean.shape == (1024,)
    vectors = np.ascontiguousarray(vectors[:N_KEEP])
    x = cases.frontend_pca_inputs(1024)
    y = reduce_dim_cv2(x, mean, vectors)
    yo = orc.pca_project(x, mean, vectors, True)
    err = np.abs(y - yo).max()
    print("pca: cv2 vs restatement max abs err", err, "bit-equal rows", int((y.view(np.uint32) == yo.view(np.uint32)).all(1).sum()), "/", len(y))
    assert err < 2e-7, "restatement and cv2 disagree"
    np.savez_compressed(os.path.join(cases.GOLDEN, "frontend_pca.npz"), vectors=vectors, mean=mean, y=y, input_sha=cases.sha(x))

    # a small cv::PCA model file written by cv::FileStorage itself (the format PCAUtils::loadModel reads), for the
    # model-file reader of the C ABI; the arrays it holds are committed beside it
    rng = np.random.Generator(np.random.PCG64(0x9CA))
    q, _ = np.linalg.qr(rng.standard_normal((64, 64)))
    small_vec = np.ascontiguousarray(q.astype(np.float32))
    small_mean = (rng.standard_normal((1, 64)) * 0.3).astype(np.float32)
    small_val = np.sort(rng.random((64, 1)).astype(np.float32), axis=0)[::-1].copy()
    yml = os.path.join(cases.GOLDEN, "pca_small_64x64.yml")
    fs_w = cv2.FileStorage(yml, cv2.FILE_STORAGE_WRITE)
    fs_w.write("name", "PCA")
    fs_w.write("vectors", small_vec)
    fs_w.write("values", small_val)
    fs_w.write("mean", small_mean)
    fs_w.release()
    fs_r = cv2.FileStorage(yml, cv2.FILE_STORAGE_READ)   # what OpenCV itself reads back (8 significant digits are printed)
    np.savez_compressed(os.path.join(cases.GOLDEN, "pca_small_64x64.npz"), vectors=fs_r.getNode("vectors").mat(),
                        values=fs_r.getNode("values").mat().reshape(-1), mean=fs_r.getNode("mean").mat().reshape(-1))

    d = cases.frontend_sift_inputs()
    r = root_sift_cv2(d)
    ro = orc.rootsift(d)
    err = np.abs(r - ro).max()
    print("rootsift: cv2 vs restatement max abs err", err, "bit-equal rows", int((r.view(np.uint32) == ro.view(np.uint32)).all(1).sum()), "/", len(r))
    assert err < 1e-7, "restatement and cv2 disagree"
    np.savez_compressed(os.path.join(cases.GOLDEN, "frontend_rootsift.npz"), y=r, input_sha=cases.sha(d))


if __name__ == "__main__":
    main()
