// capi_ctx.cu -- context, error reporting, stream/event plumbing of the C ABI.
#include "capi_common.cuh"
#include "topk.cuh"

namespace b200nn {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

int ensure_copy_engine(Ctx* c, size_t bytes) {
    if (!c->copy_stream) {
        B2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            B2_CUDA(cudaEventCreateWithFlags(&c->copy_done[b], cudaEventDisableTiming));
            B2_CUDA(cudaEventCreateWithFlags(&c->compute_done[b], cudaEventDisableTiming));
        }
    }
    if (bytes > c->pinned_bytes) {
        B2_CUDA(cudaStreamSynchronize(c->copy_stream));
        for (int b = 0; b < 2; b++) {
            if (c->pinned[b]) cudaFreeHost(c->pinned[b]);
            c->pinned[b] = nullptr;
        }
        c->pinned_bytes = 0;
        for (int b = 0; b < 2; b++) B2_CUDA(cudaMallocHost(&c->pinned[b], bytes));
        c->pinned_bytes = bytes;
    }
    return 0;
}
}  // namespace b200nn

using namespace b200nn;

extern "C" {

const char* b200nn_last_error(void) { return g_last_error.c_str(); }
const char* b200nn_version(void) { return "b200nn 0.1 (sm_100a)"; }

int b200nn_ctx_create(int device, b200nn_ctx_t* out) {
    if (!out) B2_FAIL(B200NN_ERR_INVALID, "ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        B2_FAIL(B200NN_ERR_CUDA, std::string("ctx_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) B2_FAIL(B200NN_ERR_INVALID, "ctx_create: bad device ordinal");
    B2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        B2_FAIL(B200NN_ERR_UNSUPPORTED, std::string("ctx_create: this library is built for sm_100a only; device is sm_") +
                                            std::to_string(prop.major) + std::to_string(prop.minor));
    b200nn_ctx* c = new b200nn_ctx();
    c->c.device = device;
    c->c.sm_count = prop.multiProcessorCount;
    c->c.smem_optin = prop.sharedMemPerBlockOptin;
    B2_CUDA(cudaStreamCreateWithFlags(&c->c.own_stream, cudaStreamNonBlocking));
    c->c.stream = c->c.own_stream;
    for (int i = 0; i < 16; i++) B2_CUDA(cudaEventCreate(&c->c.events[i]));
    B2_CUDA(cudaMalloc(&c->c.d_err, sizeof(int)));
    B2_CUDA(cudaMemset(c->c.d_err, 0, sizeof(int)));
    *out = c;
    return 0;
}

void b200nn_ctx_destroy(b200nn_ctx_t ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    for (int i = 0; i < 16; i++)
        if (ctx->c.events[i]) cudaEventDestroy(ctx->c.events[i]);
    if (ctx->c.d_err) cudaFree(ctx->c.d_err);
    if (ctx->c.dmat) cudaFree(ctx->c.dmat);
    for (int b = 0; b < 2; b++) {
        if (ctx->c.pinned[b]) cudaFreeHost(ctx->c.pinned[b]);
        if (ctx->c.copy_done[b]) cudaEventDestroy(ctx->c.copy_done[b]);
        if (ctx->c.compute_done[b]) cudaEventDestroy(ctx->c.compute_done[b]);
    }
    if (ctx->c.copy_stream) cudaStreamDestroy(ctx->c.copy_stream);
    if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.own_stream);
    delete ctx;
}

int b200nn_ctx_set_stream(b200nn_ctx_t ctx, void* cuda_stream) {
    if (!ctx) B2_FAIL(B200NN_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    B2_CUDA(cudaStreamSynchronize(ctx->c.stream));
    ctx->c.stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->c.own_stream;
    return 0;
}

int b200nn_ctx_synchronize(b200nn_ctx_t ctx) {
    if (!ctx) B2_FAIL(B200NN_ERR_INVALID, "ctx is NULL");
    B2_CUDA(cudaSetDevice(ctx->c.device));
    B2_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return 0;
}

int b200nn_ctx_launch_count(b200nn_ctx_t ctx, uint64_t* out) {
    if (!ctx || !out) B2_FAIL(B200NN_ERR_INVALID, "ctx/out is NULL");
    *out = ctx->c.launches;
    return 0;
}

int b200nn_ctx_event_record(b200nn_ctx_t ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) B2_FAIL(B200NN_ERR_INVALID, "event slot must be in [0,8)");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    B2_CUDA(cudaEventRecord(ctx->c.events[8 + slot], ctx->c.stream));
    return 0;
}

int b200nn_ctx_event_elapsed_ms(b200nn_ctx_t ctx, int a, int b, float* ms) {
    if (!ctx || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) B2_FAIL(B200NN_ERR_INVALID, "bad event slot");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    B2_CUDA(cudaEventSynchronize(ctx->c.events[8 + b]));
    B2_CUDA(cudaEventElapsedTime(ms, ctx->c.events[8 + a], ctx->c.events[8 + b]));
    return 0;
}

int b200nn_topk_merge_dev(b200nn_ctx_t ctx, const uint64_t* keys_dev, int L, size_t nq, size_t k, float* out_dist_dev,
                          uint64_t* out_id_dev) {
    if (!ctx || !keys_dev || L < 1 || k < 1) B2_FAIL(B200NN_ERR_INVALID, "topk_merge: bad arguments");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    return launch_topk_merge(&ctx->c, (const unsigned long long*)keys_dev, L, (long long)nq, (int)k, (long long)(nq * k),
                             out_dist_dev, nullptr, (unsigned long long*)out_id_dev, nullptr);
}

// The all-gathered records of a (query chunk x row shard) grid of ranks: keys [n_chunks][L][chunk_q][k],
// rank = chunk * L + shard; query q of the batch is row q % chunk_q of chunk q / chunk_q.  One launch.
int b200nn_topk_merge_grid_dev(b200nn_ctx_t ctx, const uint64_t* keys_dev, int n_chunks, int L, size_t chunk_q, size_t nq,
                               size_t k, float* out_dist_dev, uint64_t* out_id_dev) {
    if (!ctx || !keys_dev || L < 1 || k < 1 || n_chunks < 1 || chunk_q < 1) B2_FAIL(B200NN_ERR_INVALID, "topk_merge_grid: bad arguments");
    if (nq > (size_t)n_chunks * chunk_q) B2_FAIL(B200NN_ERR_INVALID, "topk_merge_grid: nq exceeds n_chunks * chunk_q");
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(ctx->c.device));
    return launch_topk_merge(&ctx->c, (const unsigned long long*)keys_dev, L, (long long)nq, (int)k, (long long)(chunk_q * k),
                             out_dist_dev, nullptr, (unsigned long long*)out_id_dev, nullptr);
}

}  // extern "C"
