// capi_front.cu -- C ABI of the front end: PCA projection + L2 normalisation
// (cvtk::PCAUtils, pca_train_project/pca_online/pca_utils.h) and rootSIFT (siftsIDX::rootSift).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

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


namespace {

// One `!!opencv-matrix` node of a cv::FileStorage YAML file: "<key>: !!opencv-matrix / rows: R / cols: C / dt: f|d /
// data: [ v, v, ... ]" (what cv::PCA::write / cv::FileStorage << Mat emit).  Numbers are parsed like OpenCV does: as
// double, then narrowed to float.  Returns false when the key or a well-formed node is missing.
bool yaml_matrix(const std::string& txt, const char* key, int* rows, int* cols, std::vector<float>* data) {
    const std::string tag = std::string("\n") + key + ": !!opencv-matrix";
    size_t p = txt.find(tag);
    if (p == std::string::npos) return false;
    p += tag.size();
    auto field = [&](const char* name, size_t from, size_t* after) -> std::string {
        const size_t q = txt.find(name, from);
        if (q == std::string::npos) return std::string();
        const size_t e = txt.find('\n', q);
        *after = e;
        return txt.substr(q + strlen(name), e - q - strlen(name));
    };
    size_t a = p, b = p, c = p;
    const std::string r = field("rows:", p, &a), cc = field("cols:", p, &b), dt = field("dt:", p, &c);
    if (r.empty() || cc.empty() || dt.empty()) return false;
    *rows = atoi(r.c_str());
    *cols = atoi(cc.c_str());
    const size_t t0 = dt.find_first_not_of(" \t");
    if (t0 == std::string::npos || (dt[t0] != 'f' && dt[t0] != 'd')) return false;  // CV_32F or CV_64F, one channel
    const size_t lb = txt.find('[', c), rb = txt.find(']', c);
    if (lb == std::string::npos || rb == std::string::npos || rb < lb || *rows < 0 || *cols < 0) return false;
    const size_t want = (size_t)*rows * (size_t)*cols;
    if (data) {
        data->clear();
        data->reserve(want);
        const char* s = txt.c_str() + lb + 1;
        const char* end = txt.c_str() + rb;
        while (s < end) {
            while (s < end && (*s == ' ' || *s == ',' || *s == '\n' || *s == '\r' || *s == '\t')) s++;
            if (s >= end) break;
            char* e = nullptr;
            const double v = strtod(s, &e);  // also reads .Inf / .NaN? no: OpenCV writes those as .Inf/.NaN -- rejected below
            if (e == s) return false;
            data->push_back((float)v);
            s = e;
        }
        if (data->size() != want) return false;
    }
    return true;
}

int read_pca_yaml(const char* path, int* K, int* N, std::vector<float>* mean, std::vector<float>* vectors, std::vector<float>* values) {
    FILE* f = fopen(path, "rb");
    if (!f) B2_FAIL(B200NN_ERR_IO, std::string("pca_read_model: cannot open ") + path);
    std::string txt = "\n";
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) txt.append(buf, got);
    fclose(f);
    int vr = 0, vc = 0, mr = 0, mc = 0, er = 0, ec = 0;
    if (!yaml_matrix(txt, "vectors", &vr, &vc, vectors) || !yaml_matrix(txt, "mean", &mr, &mc, mean))
        B2_FAIL(B200NN_ERR_IO, std::string("pca_read_model: no well-formed `vectors` / `mean` opencv-matrix in ") + path);
    if (mr * mc != vc || vr < 1 || vc < 1) B2_FAIL(B200NN_ERR_IO, "pca_read_model: mean and vectors disagree on the input dimension");
    if (values) {
        if (!yaml_matrix(txt, "values", &er, &ec, values) || er * ec != vr) values->assign(vr, 0.0f);  // eigenvalues are not used by project()
    }
    *K = vc;
    *N = vr;
    return 0;
}

}  // namespace

extern "C" {

int b200nn_pca_read_model(const char* path, int* K_in, int* N_out, float* mean, float* vectors, float* values) {
    if (!path || !K_in || !N_out) B2_FAIL(B200NN_ERR_INVALID, "pca_read_model: NULL argument");
    std::vector<float> m, v, e;
    const bool fill = mean || vectors || values;
    const int rc = read_pca_yaml(path, K_in, N_out, fill ? &m : nullptr, fill ? &v : nullptr, values ? &e : nullptr);
    if (rc) return rc;
    if (mean) memcpy(mean, m.data(), sizeof(float) * m.size());
    if (vectors) memcpy(vectors, v.data(), sizeof(float) * v.size());
    if (values) memcpy(values, e.data(), sizeof(float) * e.size());
    return 0;
}

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

int b200nn_proj_load_model(b200nn_ctx_t ctx, const char* path, b200nn_proj_t* out) {
    if (!ctx || !path || !out) B2_FAIL(B200NN_ERR_INVALID, "proj_load_model: NULL argument");
    int K = 0, N = 0;
    std::vector<float> m, v;
    const int rc = read_pca_yaml(path, &K, &N, &m, &v, nullptr);
    if (rc) return rc;
    return b200nn_proj_create(ctx, K, N, m.data(), v.data(), out);
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
