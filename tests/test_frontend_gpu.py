"""GPU parity tests of the front end through the C ABI: rootSIFT is bit-exact against the
cv2 golden vectors; the PCA projection runs as a tcgen05 split-TF32 GEMM (fp32-level accumulation, the reference's
OpenCV gemm accumulates in double) and must stay within 2e-5 absolute on unit-norm outputs -- five times inside
north_star's 1e-4 relative tolerance for fp32 results."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
G = cases.GOLDEN


@pytest.fixture(scope="module")
def ctx():
    from cvt_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_rootsift_bit_exact_vs_cv2_golden(ctx):
    from cvt_b200 import capi
    gold = np.load(os.path.join(G, "frontend_rootsift.npz"))
    d = cases.frontend_sift_inputs()
    assert str(gold["input_sha"]) == cases.sha(d)
    y = capi.rootsift(ctx, d)
    assert np.array_equal(y.view(np.uint32), gold["y"].view(np.uint32))
    # ragged shapes (rows not a multiple of the block, odd descriptor lengths) against the restatement
    rng = np.random.Generator(np.random.PCG64(5))
    for n, dd in ((1, 128), (129, 128), (1000, 64), (37, 100), (300, 513)):
        x = (rng.standard_normal((n, dd), dtype=np.float32) * np.float32(60.0)).astype(np.float32)
        assert np.array_equal(capi.rootsift(ctx, x).view(np.uint32), orc.rootsift(x).view(np.uint32)), (n, dd)


def test_pca_reduce_dim_vs_cv2_golden(ctx):
    """The reference's shipped GoogLeNet PCA model (first 64 eigenvectors) on seeded 1024-d inputs."""
    from cvt_b200 import capi
    gold = np.load(os.path.join(G, "frontend_pca.npz"))
    x = cases.frontend_pca_inputs(1024)
    assert str(gold["input_sha"]) == cases.sha(x)
    proj = capi.Projection(ctx, gold["vectors"], gold["mean"])
    y = proj.reduce_dim(x, l2norm=True)
    err = float(np.abs(y - gold["y"]).max())
    assert err <= 2e-5, err
    assert np.allclose(np.linalg.norm(y, axis=1), 1.0, atol=1e-5)
    # un-normalised projection against the double-accumulating restatement, relative to the input scale
    yp = proj.reduce_dim(x, l2norm=False)
    op = orc.pca_project(x, gold["mean"], gold["vectors"], False)
    scale = np.linalg.norm(x - gold["mean"], axis=1, keepdims=True)
    assert float((np.abs(yp - op) / scale).max()) <= 5e-6  # measured 2.4e-6 with one accumulator (truncating tensor-core adds)
    proj.close()


@pytest.mark.parametrize("K,N,n,l2norm", [(1024, 128, 700, True), (2048, 256, 260, True), (64, 64, 5, True), (96, 192, 131, False),
                                           (128, 128, 1, True), (512, 320, 77, False)])
def test_projection_shapes_vs_oracle(ctx, K, N, n, l2norm):
    """cvtk::PCAUtils::reduceDim shapes of both shipped models (1024->128, 2048->256) and ragged ones, random
    orthonormal-ish bases, against the restatement."""
    from cvt_b200 import capi
    rng = np.random.Generator(np.random.PCG64(K * 31 + N))
    V = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    mean = (rng.standard_normal(K) * 0.3).astype(np.float32)
    x = np.maximum(rng.standard_normal((n, K), dtype=np.float32), 0).astype(np.float32)
    proj = capi.Projection(ctx, V, mean)
    y = proj.reduce_dim(x, l2norm=l2norm)
    o = orc.pca_project(x, mean, V, l2norm)
    if l2norm:
        assert float(np.abs(y - o).max()) <= 2e-5
    else:
        scale = np.linalg.norm(x - mean, axis=1, keepdims=True)
        assert float((np.abs(y - o) / scale).max()) <= 5e-6
    proj.close()


def test_projection_errors(ctx):
    from cvt_b200 import capi
    with pytest.raises(capi.B200nnError):
        capi.Projection(ctx, np.zeros((100, 1024), np.float32))  # N % 64 != 0
    with pytest.raises(capi.B200nnError):
        capi.Projection(ctx, np.zeros((128, 1000), np.float32))  # K % 32 != 0
    p = capi.Projection(ctx, np.zeros((192, 64), np.float32))
    with pytest.raises(capi.B200nnError):
        p.reduce_dim(np.zeros((3, 64), np.float32), l2norm=True)  # fused normalisation needs N in {64,128,256}
    p.close()


def test_projection_from_yaml_model_file(ctx):
    """PCAUtils::loadModel + reduceDim from a cv::PCA YAML file (fixture written by cv2, tests/golden/pca_small_64x64.yml)."""
    from cvt_b200 import capi
    gold = np.load(os.path.join(G, "pca_small_64x64.npz"))
    proj = capi.Projection.load_model(ctx, os.path.join(G, "pca_small_64x64.yml"))
    assert (proj.N, proj.K) == (64, 64)
    x = cases.frontend_pca_inputs(64, 200)
    y = proj.reduce_dim(x, l2norm=True)
    o = orc.pca_project(x, gold["mean"], gold["vectors"], True)
    assert float(np.abs(y - o).max()) <= 2e-5
    proj.close()
