// capi_sq.cu -- C ABI of the 8-bit scalar quantizer: cvtk::quant::Int8Quan drop-in
// (scalar_quantization/scalar_quantization/int8_quan.h:17-37, int8_quan.cc:46-132).
#include <algorithm>

#include "capi_common.cuh"
#include "sq_kernels.cuh"

using namespace b200nn;

struct b200nn_sq {
    b200nn_ctx* ctx = nullptr;
    int d = 0;
    DevBuf<float> vmin, vdiff, ws_x;
    DevBuf<unsigned char> ws_codes;
};

namespace {
struct SGuard {
    std::lock_guard<std::mutex> g;
    explicit SGuard(b200nn_ctx* c) : g(c->mu) { cudaSetDevice(c->c.device); }
};
}  // namespace

extern "C" {

int b200nn_sq_create(b200nn_ctx_t ctx, int d, const float* vmin, const float* vdiff, b200nn_sq_t* out) {
    if (!ctx || !out || !vmin || !vdiff || d <= 0) B2_FAIL(B200NN_ERR_INVALID, "sq_create: bad argument");
    SGuard g(ctx);
    b200nn_sq* s = new b200nn_sq();
    s->ctx = ctx;
    s->d = d;
    int rc;
    if ((rc = s->vmin.ensure(d)) || (rc = s->vdiff.ensure(d))) { delete s; return rc; }
    cudaMemcpyAsync(s->vmin.p, vmin, sizeof(float) * d, cudaMemcpyHostToDevice, ctx->c.stream);
    cudaMemcpyAsync(s->vdiff.p, vdiff, sizeof(float) * d, cudaMemcpyHostToDevice, ctx->c.stream);
    if (cudaStreamSynchronize(ctx->c.stream) != cudaSuccess) { delete s; B2_FAIL(B200NN_ERR_CUDA, "sq_create: upload failed"); }
    *out = s;
    return 0;
}

void b200nn_sq_destroy(b200nn_sq_t sq) {
    if (!sq) return;
    {
        SGuard g(sq->ctx);
        cudaStreamSynchronize(sq->ctx->c.stream);
    }
    delete sq;
}

int b200nn_sq_train_minmax(b200nn_ctx_t ctx, int d, const float* x, size_t n, float* vmin, float* vdiff) {
    if (!ctx || !x || !vmin || !vdiff || d <= 0 || n == 0) B2_FAIL(B200NN_ERR_INVALID, "sq_train_minmax: bad argument");
    SGuard g(ctx);
    Ctx* c = &ctx->c;
    DevBuf<float> dx, dmin, ddiff;
    DevBuf<uint32_t> scratch;
    int rc;
    const size_t chunk = std::min<size_t>(n, (size_t)1 << 20);
    if ((rc = dx.ensure(chunk * d)) || (rc = dmin.ensure(d)) || (rc = ddiff.ensure(d)) || (rc = scratch.ensure(2 * (size_t)d))) return rc;
    // min/max are associative and exact: chunks accumulate into the same scratch
    if (n <= chunk) {
        B2_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(float) * n * d, cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_sq_train_minmax(c, dx.p, (long long)n, d, scratch.p, dmin.p, ddiff.p))) return rc;
    } else {
        B2_FAIL(B200NN_ERR_UNSUPPORTED, "sq_train_minmax: more than 2^20 rows per call (call on a sample, as sq_train.cpp does)");
    }
    B2_CUDA(cudaMemcpyAsync(vmin, dmin.p, sizeof(float) * d, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(vdiff, ddiff.p, sizeof(float) * d, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_sq_encode_dev(b200nn_sq_t sq, float* x_dev, size_t n, int l2norm, uint8_t* codes_dev) {
    if (!sq || (n && (!x_dev || !codes_dev))) B2_FAIL(B200NN_ERR_INVALID, "sq_encode_dev: NULL argument");
    SGuard g(sq->ctx);
    return launch_sq_encode(&sq->ctx->c, x_dev, (long long)n, sq->d, sq->vmin.p, sq->vdiff.p, l2norm, codes_dev);
}

int b200nn_sq_encode(b200nn_sq_t sq, float* x, size_t n, int l2norm, uint8_t* codes) {
    if (!sq || (n && (!x || !codes))) B2_FAIL(B200NN_ERR_INVALID, "sq_encode: NULL argument");
    if (!n) return 0;
    SGuard g(sq->ctx);
    Ctx* c = &sq->ctx->c;
    int rc;
    const size_t chunk = std::min<size_t>(n, (size_t)1 << 20);
    if ((rc = sq->ws_x.ensure(chunk * sq->d)) || (rc = sq->ws_codes.ensure(chunk * sq->d))) return rc;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t cn = std::min(chunk, n - off);
        B2_CUDA(cudaMemcpyAsync(sq->ws_x.p, x + off * sq->d, sizeof(float) * cn * sq->d, cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_sq_encode(c, sq->ws_x.p, (long long)cn, sq->d, sq->vmin.p, sq->vdiff.p, l2norm, sq->ws_codes.p))) return rc;
        if (l2norm)  // the reference normalises the caller's buffer in place (int8_quan.cc:76-78)
            B2_CUDA(cudaMemcpyAsync(x + off * sq->d, sq->ws_x.p, sizeof(float) * cn * sq->d, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaMemcpyAsync(codes + off * sq->d, sq->ws_codes.p, cn * sq->d, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int b200nn_sq_decode(b200nn_sq_t sq, const uint8_t* codes, size_t n, int faiss_float_variant, float* x) {
    if (!sq || (n && (!x || !codes))) B2_FAIL(B200NN_ERR_INVALID, "sq_decode: NULL argument");
    if (!n) return 0;
    SGuard g(sq->ctx);
    Ctx* c = &sq->ctx->c;
    int rc;
    const size_t chunk = std::min<size_t>(n, (size_t)1 << 20);
    if ((rc = sq->ws_x.ensure(chunk * sq->d)) || (rc = sq->ws_codes.ensure(chunk * sq->d))) return rc;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t cn = std::min(chunk, n - off);
        B2_CUDA(cudaMemcpyAsync(sq->ws_codes.p, codes + off * sq->d, cn * sq->d, cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_sq_decode(c, sq->ws_codes.p, (long long)cn, sq->d, sq->vmin.p, sq->vdiff.p, faiss_float_variant, sq->ws_x.p))) return rc;
        B2_CUDA(cudaMemcpyAsync(x + off * sq->d, sq->ws_x.p, sizeof(float) * cn * sq->d, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

}  // extern "C"
