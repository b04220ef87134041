// capi_front.cu -- C ABI of the front end: PCA projection + L2 normalisation
// (cvtk::PCAUtils, pca_train_project/pca_online/pca_utils.h) and rootSIFT (siftsIDX::rootSift).
#include <algorithm>

#include "capi_common.cuh"
#include "front_kernels.cuh"
#include "rotate_gemm.cuh"

using namespace b200nn;

struct b200nn_proj {
    b200nn_ctx* ctx = nullptr;
    int K = 0, N = 0;
    bool has_mean = false;
    DevBuf<float> mean, planes_norm, planes_plain, ws_x, ws_y;
    bool norm_ok = false, plain_ok = false;
};

extern "C" {

int b200nn_proj_create(b200nn_ctx_t ctx, int K_in, int N_out, const float* mean, const float* vectors, b200nn_proj_t* out) {
    if (!ctx || !vectors || !out) B2_FAIL(B200NN_ERR_INVALID, "proj_create: NULL argument");
    const bool norm_ok = proj_gemm_supported(K_in, N_out, true), plain_ok = proj_gemm_supported(K_in, N_out, false);
    if (!norm_ok && !plain_ok)
        B2_FAIL(B200NN_ERR_UNSUPPORTED, "proj_create: need K % 32 == 0 and N % 64 == 0 (tcgen05 projection GEMM)");
    std::lock_guard<std::mutex> g(ctx->mu);
    Ctx* c = &ctx->c;
    B2_CUDA(cudaSetDevice(c->device));
    b200nn_proj* p = new b200nn_proj();
    p->ctx = ctx; p->K = K_in; p->N = N_out; p->norm_ok = norm_ok; p->plain_ok = plain_ok;
    int rc = 0;
    auto fail = [&](int code) { delete p; return code; };
    std::vector<float> packed;
    if (norm_ok) {
        proj_gemm_pack(vectors, N_out, K_in, proj_gemm_nblock(N_out, true), packed);
        if ((rc = p->planes_norm.ensure(packed.size()))) return fail(rc);
        cudaMemcpyAsync(p->planes_norm.p, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    if (plain_ok) {
        proj_gemm_pack(vectors, N_out, K_in, proj_gemm_nblock(N_out, false), packed);
        if ((rc = p->planes_plain.ensure(packed.size()))) return fail(rc);
        cudaMemcpyAsync(p->planes_plain.p, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    if (mean) {
        if ((rc = p->mean.ensure(K_in))) return fail(rc);
        cudaMemcpyAsync(p->mean.p, mean, sizeof(float) * K_in, cudaMemcpyHostToDevice, c->stream);
        p->has_mean = true;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        set_last_error("proj_create: device upload failed");
        return fail(B200NN_ERR_CUDA);
    }
    *out = p;
    return 0;
}

void b200nn_proj_destroy(b200nn_proj_t p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(p->ctx->mu);
        cudaSetDevice(p->ctx->c.device);
        cudaStreamSynchronize(p->ctx->c.stream);
    }
    delete p;
}

static int proj_apply_locked(b200nn_proj* p, const float* x_dev, size_t n, int l2norm, float* y_dev) {
    if (l2norm && !p->norm_ok) B2_FAIL(B200NN_ERR_UNSUPPORTED, "proj_apply: the fused L2 normalisation needs N in {64, 128, 256}");
    if (!l2norm && !p->plain_ok) B2_FAIL(B200NN_ERR_UNSUPPORTED, "proj_apply: unsupported shape");
    return launch_proj_gemm(&p->ctx->c, x_dev, (long long)n, p->K, p->N, p->has_mean ? p->mean.p : nullptr,
                            l2norm ? p->planes_norm.p : p->planes_plain.p, l2norm != 0, y_dev);
}

int b200nn_proj_apply_dev(b200nn_proj_t p, const float* x_dev, size_t n, int l2norm, float* y_dev) {
    if (!p || (n && (!x_dev || !y_dev))) B2_FAIL(B200NN_ERR_INVALID, "proj_apply_dev: NULL argument");
    std::lock_guard<std::mutex> g(p->ctx->mu);
    B2_CUDA(cudaSetDevice(p->ctx->c.device));
    return proj_apply_locked(p, x_dev, n, l2norm, y_dev);
}

int b200nn_proj_apply(b200nn_proj_t p, const float* x, size_t n, int l2norm, float* y) {
    if (!p || (n && (!x || !y))) B2_FAIL(B200NN_ERR_INVALID, "proj_apply: NULL argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> g(p->ctx->mu);
    Ctx* c = &p->ctx->c;
    B2_CUDA(cudaSetDevice(c->device));
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(n, (size_t)(256u << 20) / ((size_t)p->K * 4)));  // <= 256 MB of input staged
    int rc;
    if ((rc = p->ws_x.ensure(chunk * p->K)) || (rc = p->ws_y.ensure(chunk * p->N))) return rc;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t cn = std::min(chunk, n - off);
        B2_CUDA(cudaMemcpyAsync(p->ws_x.p, x + off * p->K, sizeof(float) * cn * p->K, cudaMemcpyHostToDevice, c->stream));
        if ((rc = proj_apply_locked(p, p->ws_x.p, cn, l2norm, p->ws_y.p))) return rc;
        B2_CUDA(cudaMemcpyAsync(y + off * p->N, p->ws_y.p, sizeof(float) * cn * p->N, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int b200nn_rootsift_dev(b200nn_ctx_t ctx, float* x_dev, size_t n, int d, float eps) {
    if (!ctx || (n && !x_dev)) B2_FAIL(B200NN_ERR_INVALID, "rootsift_dev: NULL argument");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    return launch_rootsift(&ctx->c, x_dev, (long long)n, d, eps);
}

int b200nn_rootsift(b200nn_ctx_t ctx, float* x, size_t n, int d, float eps) {
    if (!ctx || (n && !x)) B2_FAIL(B200NN_ERR_INVALID, "rootsift: NULL argument");
    if (!n) return 0;
    if (d < 1) B2_FAIL(B200NN_ERR_INVALID, "rootsift: d must be >= 1");
    std::lock_guard<std::mutex> g(ctx->mu);
    Ctx* c = &ctx->c;
    B2_CUDA(cudaSetDevice(c->device));
    DevBuf<float> buf;
    int rc;
    if ((rc = buf.ensure(n * (size_t)d))) return rc;
    B2_CUDA(cudaMemcpyAsync(buf.p, x, sizeof(float) * n * d, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_rootsift(c, buf.p, (long long)n, d, eps))) return rc;
    B2_CUDA(cudaMemcpyAsync(x, buf.p, sizeof(float) * n * d, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
