// capi_pq.cu -- C ABI of the (O)PQ / IVFOPQ path: model, add (rotate+encode), LUT, scores, search,
// index persistence.  Replaces IVFOPQ's public methods (opq/src/IVFOPQ.h:31-47).
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <string>

#include "capi_common.cuh"
#include "capi_pq_internal.cuh"
#include "pq_kernels.cuh"
#include "rotate_gemm.cuh"
#include "topk.cuh"

using namespace b200nn;

struct b200nn_pq {
    b200nn_ctx* ctx = nullptr;
    int D = 0, K = 0, M = 0, ksub = 0, ds = 0;
    float clamp = 1.0f;
    bool has_perm = false, has_R = false;
    DevBuf<float> coarse, coarseT, cb, cbT, rplanes;
    DevBuf<int> perm;
    // rows in insertion order
    long long n = 0, n_groups = 0;
    DevBuf<unsigned char> codes;
    DevBuf<int> list, group;
    // scan layout (flat ADC fast path)
    DevBuf<uint32_t> codesT;
    long long codesT_rows = -1;
    // CSR by coarse list (generic IVF path)
    long long csr_rows = -1;
    DevBuf<unsigned char> codes_sorted;
    DevBuf<int> group_sorted, row_sorted;
    DevBuf<long long> list_off;
    // workspaces
    DevBuf<float> ws_x, ws_q, ws_qraw, ws_lut, ws_scores, ws_dist, ws_warm;
    DevBuf<int> ws_probes, ws_list;
    DevBuf<unsigned char> ws_codes;
    DevBuf<unsigned long long> ws_keys, ws_keys2, ws_id;
    // work plan of the fused scan, cached per (query groups, granules) shape
    ScanPlan plan;
    long long plan_qgroups = -1, plan_gran = -1;
    DevBuf<int> plan_desc;
    float timing[4] = {0, 0, 0, 0};
};

namespace {

__global__ void iota_kernel(int* p, long long n, int start) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = start + (int)i;
}
__global__ void gather_rows_kernel(const unsigned char* __restrict__ src, const int* __restrict__ rows, long long n, int M,
                                   unsigned char* __restrict__ dst) {
    const long long total = n * M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / M;
        dst[i] = src[(long long)rows[r] * M + (i - r * M)];
    }
}
__global__ void gather_int_kernel(const int* __restrict__ src, const int* __restrict__ rows, long long n, int* __restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[rows[i]];
}
__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
    const long long total = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        dst[(long long)c * rows + r] = src[i];
    }
}

inline unsigned small_grid(long long work) { return (unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, 148 * 8)); }

struct Guard {
    std::lock_guard<std::mutex> g;
    explicit Guard(b200nn_pq* p) : g(p->ctx->mu) { cudaSetDevice(p->ctx->c.device); }
};

int pq_init(b200nn_ctx_t ctx, int D, int K, int M, int ksub, const float* coarse, const float* cb, const int32_t* perm,
            const float* R, float clamp, b200nn_pq_t* out) {
    if (!ctx || !out || !coarse || !cb) B2_FAIL(B200NN_ERR_INVALID, "pq_create: NULL argument");
    if (D <= 0 || K <= 0 || M <= 0 || ksub <= 0 || ksub > 256 || D % M != 0)
        B2_FAIL(B200NN_ERR_INVALID, "pq_create: need D,K,M > 0, D % M == 0, 1 <= ksub <= 256");
    const int ds = D / M;
    if (!(ds == 1 || ds == 2 || ds == 4 || ds == 8 || ds == 16 || ds == 32))
        B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_create: D/M must be one of 1,2,4,8,16,32");
    if (R && perm) B2_FAIL(B200NN_ERR_INVALID, "pq_create: give either a permutation or a dense rotation R, not both");
    if (R && !rotate_gemm_supported(D)) B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_create: the dense rotation (tcgen05 GEMM) needs D % 64 == 0");
    if (perm) {
        std::vector<char> seen(D, 0);
        for (int i = 0; i < D; i++) {
            if (perm[i] < 0 || perm[i] >= D || seen[perm[i]]) B2_FAIL(B200NN_ERR_INVALID, "pq_create: perm is not a permutation of 0..D-1");
            seen[perm[i]] = 1;
        }
    }
    Ctx* c = &ctx->c;
    std::lock_guard<std::mutex> g(ctx->mu);
    B2_CUDA(cudaSetDevice(c->device));
    b200nn_pq* p = new b200nn_pq();
    p->ctx = ctx; p->D = D; p->K = K; p->M = M; p->ksub = ksub; p->ds = ds; p->clamp = clamp;
    int rc = 0;
    auto fail = [&](int code) { delete p; return code; };
    if ((rc = p->coarse.ensure((size_t)K * D)) || (rc = p->coarseT.ensure((size_t)K * D)) ||
        (rc = p->cb.ensure((size_t)M * ksub * ds)) || (rc = p->cbT.ensure((size_t)M * ksub * ds)))
        return fail(rc);
    cudaMemcpyAsync(p->coarse.p, coarse, sizeof(float) * K * D, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(p->cb.p, cb, sizeof(float) * M * ksub * ds, cudaMemcpyHostToDevice, c->stream);
    transpose_kernel<<<small_grid((long long)K * D), 256, 0, c->stream>>>(p->coarse.p, K, D, p->coarseT.p);
    for (int m = 0; m < M; m++)  // cb[m] is [ksub][ds] -> cbT[m] is [ds][ksub]
        transpose_kernel<<<small_grid((long long)ksub * ds), 256, 0, c->stream>>>(p->cb.p + (size_t)m * ksub * ds, ksub, ds,
                                                                                  p->cbT.p + (size_t)m * ksub * ds);
    c->launches += 1 + M;
    if (perm) {
        if ((rc = p->perm.ensure(D))) return fail(rc);
        cudaMemcpyAsync(p->perm.p, perm, sizeof(int) * D, cudaMemcpyHostToDevice, c->stream);
        p->has_perm = true;
    }
    std::vector<float> packed;
    if (R) {
        rotate_gemm_pack_R(R, D, packed);
        if ((rc = p->rplanes.ensure(packed.size()))) return fail(rc);
        cudaMemcpyAsync(p->rplanes.p, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice, c->stream);
        p->has_R = true;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        set_last_error("pq_create: device upload failed");
        return fail(B200NN_ERR_CUDA);
    }
    *out = p;
    return 0;
}

// rotate n device rows into dst (or alias src when there is no rotation)
int rotate_dev(b200nn_pq* p, const float* src, long long n, float* dst, const float** out) {
    if (p->has_R) {
        int rc = launch_rotate_gemm(&p->ctx->c, src, n, p->D, p->rplanes.p, dst);
        if (rc) return rc;
        *out = dst;
    } else if (p->has_perm) {
        int rc = launch_rotate_perm(&p->ctx->c, src, n, p->D, p->perm.p, dst);
        if (rc) return rc;
        *out = dst;
    } else {
        *out = src;
    }
    return 0;
}

int encode_dev(b200nn_pq* p, const float* x_rot, long long n, int* list, unsigned char* codes) {
    Ctx* c = &p->ctx->c;
    int rc;
    if (p->K == 1) B2_CUDA(cudaMemsetAsync(list, 0, sizeof(int) * n, c->stream));
    else if ((rc = launch_coarse_assign(c, x_rot, n, p->D, p->coarseT.p, p->K, list))) return rc;
    return launch_pq_encode(c, x_rot, n, p->D, p->coarse.p, list, p->cbT.p, p->M, p->ksub, codes);
}

int add_dev_locked(b200nn_pq* p, const float* x_dev, long long nn, const int* group_dev, const int* group_host,
                   bool already_rotated = false) {
    Ctx* c = &p->ctx->c;
    if (nn <= 0) return 0;
    if (p->n + nn > 0x7fffffffLL) B2_FAIL(B200NN_ERR_STATE, "pq_add: more than 2^31-1 rows per shard");
    int rc;
    if ((rc = p->codes.reserve((size_t)(p->n + nn) * p->M, (size_t)p->n * p->M, c->stream))) return rc;
    if ((rc = p->list.reserve((size_t)(p->n + nn), (size_t)p->n, c->stream))) return rc;
    if ((rc = p->group.reserve((size_t)(p->n + nn), (size_t)p->n, c->stream))) return rc;
    const long long chunk = 1 << 18;
    if ((p->has_perm || p->has_R) && !already_rotated && (rc = p->ws_x.ensure((size_t)std::min(chunk, nn) * p->D))) return rc;
    for (long long off = 0; off < nn; off += chunk) {
        const long long cn = std::min(chunk, nn - off);
        const float* xr = x_dev + off * p->D;
        if (!already_rotated && (rc = rotate_dev(p, x_dev + off * p->D, cn, p->ws_x.p, &xr))) return rc;
        if ((rc = encode_dev(p, xr, cn, p->list.p + p->n + off, p->codes.p + (p->n + off) * p->M))) return rc;
    }
    if (group_dev) {
        B2_CUDA(cudaMemcpyAsync(p->group.p + p->n, group_dev, sizeof(int) * nn, cudaMemcpyDeviceToDevice, c->stream));
        std::vector<int> h;
        const int* gh = group_host;
        if (!gh) {
            h.resize(nn);
            B2_CUDA(cudaMemcpyAsync(h.data(), group_dev, sizeof(int) * nn, cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA(cudaStreamSynchronize(c->stream));
            gh = h.data();
        }
        for (long long i = 0; i < nn; i++) {
            if (gh[i] < 0) B2_FAIL(B200NN_ERR_INVALID, "pq_add: negative group id");
            p->n_groups = std::max<long long>(p->n_groups, (long long)gh[i] + 1);
        }
    } else {
        iota_kernel<<<small_grid(nn), 256, 0, c->stream>>>(p->group.p + p->n, nn, (int)p->n);
        c->launches++;
        p->n_groups = std::max(p->n_groups, p->n + nn);
    }
    p->n += nn;
    p->codesT_rows = -1;
    p->csr_rows = -1;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int ensure_scan_layout(b200nn_pq* p) {
    if (p->codesT_rows == p->n) return 0;
    const long long n_pad = ((p->n + 63) / 64) * 64;
    int rc;
    if ((rc = p->codesT.ensure((size_t)std::max<long long>(n_pad, 64) * (p->M / 4)))) return rc;
    if ((rc = launch_codes_to_scan_layout(&p->ctx->c, p->codes.p, p->n, p->M, p->codesT.p, n_pad))) return rc;
    p->codesT_rows = p->n;
    return 0;
}

int ensure_csr(b200nn_pq* p) {
    if (p->csr_rows == p->n) return 0;
    Ctx* c = &p->ctx->c;
    const long long n = p->n;
    std::vector<int> hl(n);
    B2_CUDA(cudaMemcpyAsync(hl.data(), p->list.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<long long> off(p->K + 1, 0);
    for (long long i = 0; i < n; i++) {
        if (hl[i] < 0 || hl[i] >= p->K) B2_FAIL(B200NN_ERR_STATE, "index holds a row with an invalid coarse list id");
        off[hl[i] + 1]++;
    }
    for (int k = 0; k < p->K; k++) off[k + 1] += off[k];
    std::vector<int> rows(n);
    {
        std::vector<long long> cur(off.begin(), off.end() - 1);
        for (long long i = 0; i < n; i++) rows[cur[hl[i]]++] = (int)i;  // stable: insertion order inside a list
    }
    int rc;
    if ((rc = p->row_sorted.ensure(std::max<long long>(n, 1))) || (rc = p->group_sorted.ensure(std::max<long long>(n, 1))) ||
        (rc = p->codes_sorted.ensure(std::max<long long>(n, 1) * p->M)) || (rc = p->list_off.ensure(p->K + 1)))
        return rc;
    B2_CUDA(cudaMemcpyAsync(p->row_sorted.p, rows.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(p->list_off.p, off.data(), sizeof(long long) * (p->K + 1), cudaMemcpyHostToDevice, c->stream));
    if (n) {
        gather_rows_kernel<<<small_grid(n * p->M), 256, 0, c->stream>>>(p->codes.p, p->row_sorted.p, n, p->M, p->codes_sorted.p);
        gather_int_kernel<<<small_grid(n), 256, 0, c->stream>>>(p->group.p, p->row_sorted.p, n, p->group_sorted.p);
        c->launches += 2;
    }
    B2_CUDA(cudaStreamSynchronize(c->stream));  // rows/off are host temporaries
    B2_CUDA(cudaGetLastError());
    p->csr_rows = n;
    return 0;
}

// probes + standard-layout LUTs for nq rotated device queries
int probes_and_luts(b200nn_pq* p, const float* q_rot, long long nq, int nprobe, int* probes, float* lut) {
    Ctx* c = &p->ctx->c;
    int rc;
    if (p->K == 1) B2_CUDA(cudaMemsetAsync(probes, 0, sizeof(int) * nq * nprobe, c->stream));
    else if ((rc = launch_coarse_probe(c, q_rot, nq, p->D, p->coarseT.p, p->K, nprobe, probes))) return rc;
    return launch_lut_build_std(c, q_rot, nq, p->D, probes, nprobe, p->coarse.p, p->cb.p, p->M, p->ksub, lut);
}

bool fast_path_ok(const b200nn_pq* p, int nprobe, size_t k) {
    return p->K == 1 && nprobe == 1 && p->ksub == 256 && (p->M == 4 || p->M == 8 || p->M == 16 || p->M == 32) && k <= (size_t)KP;
}

int search_dev_locked(b200nn_pq* p, const float* q_raw_dev, long long nq, int nprobe, int k, float* out_dist,
                      unsigned long long* out_id, unsigned long long* out_key, unsigned long long id_base) {
    Ctx* c = &p->ctx->c;
    if (nq <= 0) return 0;
    if (nprobe < 1 || nprobe > p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_search: nprobe must be in [1, K]");
    if (k < 1 || k > KP) B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_search: k must be in [1, 128]");
    if (id_base + (unsigned long long)p->n > 0x100000000ULL) B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_search: id_base + rows must fit 32 bits");
    int rc;
    cudaEvent_t* ev = c->events;
    B2_CUDA(cudaEventRecord(ev[0], c->stream));
    const float* qr = nullptr;
    if ((p->has_perm || p->has_R) && (rc = p->ws_q.ensure((size_t)nq * p->D))) return rc;
    if ((rc = rotate_dev(p, q_raw_dev, nq, p->ws_q.p, &qr))) return rc;
    B2_CUDA(cudaEventRecord(ev[1], c->stream));
    if (fast_path_ok(p, nprobe, (size_t)k)) {
        const int QW = scan_queries_per_cta(p->M);
        const long long qgroups = (nq + QW - 1) / QW;
        if ((rc = ensure_scan_layout(p))) return rc;
        if ((rc = p->ws_lut.ensure((size_t)qgroups * 32768))) return rc;
        if ((rc = launch_lut_build_scan(c, p->M, qr, nq, p->D, p->coarse.p, p->cb.p, p->ws_lut.p))) return rc;
        B2_CUDA(cudaEventRecord(ev[2], c->stream));
        const long long n_gran = (p->n + 63) / 64;
        if (p->plan_qgroups != qgroups || p->plan_gran != n_gran) {  // plan + tail descriptors, cached per (batch, rows) shape
            scan_plan(c->sm_count, qgroups, n_gran, &p->plan);
            if ((rc = p->plan_desc.ensure(std::max<size_t>(8, p->plan.desc.size())))) return rc;
            if (!p->plan.desc.empty())
                B2_CUDA(cudaMemcpyAsync(p->plan_desc.p, p->plan.desc.data(), sizeof(int) * p->plan.desc.size(), cudaMemcpyHostToDevice, c->stream));
            p->plan_qgroups = qgroups;
            p->plan_gran = n_gran;
        }
        const int n_full = p->plan.n_full, n_tail = p->plan.n_tail, S = p->plan.slices;
        if ((rc = p->ws_keys.ensure((size_t)S * qgroups * QW * k))) return rc;
        if (S > 1)  // whole-shard CTAs write slice 0 only: the other slices of those queries stay empty (KEY_MAX)
            B2_CUDA(cudaMemsetAsync(p->ws_keys.p, 0xFF, (size_t)S * qgroups * QW * k * sizeof(unsigned long long), c->stream));
        if ((rc = p->ws_warm.ensure(scan_warm_scratch_floats(p->M, n_full, n_tail)))) return rc;
        if ((rc = launch_adc_scan_topk(c, p->M, p->codesT.p, p->ws_lut.p, p->n, qgroups, n_full, n_tail, p->plan_desc.p, k, p->clamp,
                                       (uint32_t)id_base, p->ws_keys.p, p->ws_warm.p)))
            return rc;
        B2_CUDA(cudaEventRecord(ev[3], c->stream));
        if ((rc = launch_topk_merge(c, p->ws_keys.p, S, nq, k, qgroups * QW * k, out_dist, nullptr, out_id, out_key))) return rc;
        B2_CUDA(cudaEventRecord(ev[4], c->stream));
    } else {
        // IVF path (any K): coarse probes, then one fused CTA per query: residual LUT + list scan + top-k
        if ((rc = ensure_csr(p))) return rc;
        if ((rc = p->ws_probes.ensure((size_t)nq * nprobe)) || (rc = p->ws_keys.ensure((size_t)nq * k))) return rc;
        if (p->K == 1) B2_CUDA(cudaMemsetAsync(p->ws_probes.p, 0, sizeof(int) * nq * nprobe, c->stream));
        else if ((rc = launch_coarse_probe(c, qr, nq, p->D, p->coarseT.p, p->K, nprobe, p->ws_probes.p))) return rc;
        B2_CUDA(cudaEventRecord(ev[2], c->stream));
        if ((rc = launch_ivf_search_topk(c, qr, nq, p->D, p->ws_probes.p, nprobe, p->coarse.p, p->cb.p, p->M, p->ksub, p->list_off.p,
                                         p->codes_sorted.p, p->row_sorted.p, p->n, k, p->clamp, (uint32_t)id_base, p->ws_keys.p)))
            return rc;
        B2_CUDA(cudaEventRecord(ev[3], c->stream));
        if ((rc = launch_topk_merge(c, p->ws_keys.p, 1, nq, k, nq * k, out_dist, nullptr, out_id, out_key))) return rc;
        B2_CUDA(cudaEventRecord(ev[4], c->stream));
    }
    return 0;
}

}  // namespace

// ---- host-side pieces shared with the multi-GPU layer (capi_pq_internal.cuh) -----------------------------------
std::string pq_index_file_name(const char* dir_or_path, long long n_groups, int D, int K, int M, int ksub) {
    std::string path = dir_or_path;
    if (path.size() < 6 || path.substr(path.size() - 6) != ".fvecs")
        path += "/OPQ_Index_db_" + std::to_string(n_groups) + "_dim_" + std::to_string(D) + "_k_" + std::to_string(K) + "_PQ_m" +
                std::to_string(M) + "_k" + std::to_string(ksub) + ".fvecs";
    return path;
}

int pq_model_to_host(b200nn_pq_t p, std::vector<float>* coarse, std::vector<float>* cb, std::vector<int>* perm) {
    Guard g(p);
    Ctx* c = &p->ctx->c;
    if (coarse) {
        coarse->resize((size_t)p->K * p->D);
        B2_CUDA(cudaMemcpyAsync(coarse->data(), p->coarse.p, sizeof(float) * coarse->size(), cudaMemcpyDeviceToHost, c->stream));
    }
    if (cb) {
        cb->resize((size_t)p->M * p->ksub * p->ds);
        B2_CUDA(cudaMemcpyAsync(cb->data(), p->cb.p, sizeof(float) * cb->size(), cudaMemcpyDeviceToHost, c->stream));
    }
    if (perm) {
        perm->clear();
        if (p->has_perm) {
            perm->resize(p->D);
            B2_CUDA(cudaMemcpyAsync(perm->data(), p->perm.p, sizeof(int) * p->D, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int pq_rows_to_host(b200nn_pq_t p, PQHostRows* rows) {
    Guard g(p);
    Ctx* c = &p->ctx->c;
    const size_t n = (size_t)p->n;
    rows->lists.resize(n);
    rows->groups.resize(n);
    rows->codes.resize(n * p->M);
    if (n) {
        B2_CUDA(cudaMemcpyAsync(rows->lists.data(), p->list.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaMemcpyAsync(rows->groups.data(), p->group.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaMemcpyAsync(rows->codes.data(), p->codes.p, n * p->M, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// rows in insertion order -> the reference's file: per coarse list, elements in insertion order (IVFOPQ.cpp:541-580)
int pq_write_index_file(const std::string& path, int D, int K, int M, int ksub, long long n_groups, const float* coarse, const float* cb,
                        long long n, const int* lists, const int* groups, const unsigned char* codes, const char* const* group_paths,
                        size_t n_paths) {
    std::vector<long long> off((size_t)K + 1, 0);
    for (long long i = 0; i < n; i++) {
        if (lists[i] < 0 || lists[i] >= K) B2_FAIL(B200NN_ERR_STATE, "index holds a row with an invalid coarse list id");
        off[lists[i] + 1]++;
    }
    for (int k = 0; k < K; k++) off[k + 1] += off[k];
    std::vector<long long> order((size_t)n), cur(off.begin(), off.end() - 1);
    for (long long i = 0; i < n; i++) order[cur[lists[i]]++] = i;  // stable
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) B2_FAIL(B200NN_ERR_IO, "pq_save_index: cannot open " + path);
    const int32_t h[5] = {D, K, M, ksub, (int32_t)n_groups};
    fwrite(h, 4, 5, f);
    fwrite(coarse, 4, (size_t)K * D, f);
    fwrite(cb, 4, (size_t)M * ksub * (D / M), f);
    std::vector<unsigned char> blk;
    for (int k = 0; k < K; k++) {
        const int32_t cnt = (int32_t)(off[k + 1] - off[k]);
        fwrite(&cnt, 4, 1, f);
        blk.resize((size_t)cnt * (4 + M));
        for (long long r = off[k]; r < off[k + 1]; r++) {
            unsigned char* d = blk.data() + (size_t)(r - off[k]) * (4 + M);
            memcpy(d, &groups[order[r]], 4);
            memcpy(d + 4, codes + (size_t)order[r] * M, M);
        }
        if (cnt) fwrite(blk.data(), 1, blk.size(), f);
    }
    char name[260];
    for (long long i = 0; i < n_groups; i++) {
        memset(name, 0, sizeof name);
        if (group_paths && (size_t)i < n_paths && group_paths[i]) strncpy(name, group_paths[i], 259);
        fwrite(name, 1, 260, f);
    }
    const bool ok = !ferror(f);
    fclose(f);
    if (!ok) B2_FAIL(B200NN_ERR_IO, "pq_save_index: write failed");
    return 0;
}

// The file is untrusted input: every id and code byte is checked before it can index device memory.
int pq_parse_index_file(const char* path, int* D_, int* K_, int* M_, int* ksub_, long long* ng_, std::vector<float>* coarse,
                        std::vector<float>* cb, PQHostRows* rows) {
    FILE* f = fopen(path, "rb");
    if (!f) B2_FAIL(B200NN_ERR_IO, "Can not open the index file.");  // IVFOPQ.cpp:470
    auto bad = [&](const char* msg) { fclose(f); B2_FAIL(B200NN_ERR_IO, msg); };
    int32_t h[5];
    if (fread(h, 4, 5, f) != 5) return bad("pq_load_index: truncated header");
    const int D = h[0], K = h[1], M = h[2], ksub = h[3];
    const long long ng = h[4];
    if (D <= 0 || K <= 0 || M <= 0 || ksub <= 0 || ksub > 256 || D % M || ng < 0 || (long long)K * D > (1LL << 32))
        return bad("pq_load_index: implausible header");
    coarse->resize((size_t)K * D);
    cb->resize((size_t)M * ksub * (D / M));
    if (fread(coarse->data(), 4, coarse->size(), f) != coarse->size() || fread(cb->data(), 4, cb->size(), f) != cb->size())
        return bad("pq_load_index: truncated model section");
    std::vector<unsigned char> blk;
    const size_t rec = 4 + (size_t)M;  // int32 videoId + M code bytes per element (IVFOPQ.cpp:569-573)
    rows->lists.clear(); rows->groups.clear(); rows->codes.clear();
    for (int k = 0; k < K; k++) {
        int32_t cnt = 0;
        if (fread(&cnt, 4, 1, f) != 1 || cnt < 0) return bad("pq_load_index: truncated list section");
        if (rows->lists.size() + (size_t)cnt > 0x7fffffffull) return bad("pq_load_index: more than 2^31-1 rows");
        blk.resize((size_t)cnt * rec);
        if (cnt && fread(blk.data(), rec, (size_t)cnt, f) != (size_t)cnt) return bad("pq_load_index: truncated element");
        const size_t base = rows->lists.size();
        rows->lists.resize(base + cnt, k);
        rows->groups.resize(base + cnt);
        rows->codes.resize((base + cnt) * (size_t)M);
        for (size_t j = 0; j < (size_t)cnt; j++) {
            int32_t gid;
            memcpy(&gid, blk.data() + j * rec, 4);
            if (gid < 0) return bad("pq_load_index: negative group id in the index file");
            rows->groups[base + j] = gid;
            const unsigned char* cj = blk.data() + j * rec + 4;
            if (ksub < 256)
                for (int m = 0; m < M; m++)
                    if (cj[m] >= ksub) return bad("pq_load_index: code byte >= ksub in the index file");
            memcpy(rows->codes.data() + (base + j) * M, cj, M);
        }
    }
    fclose(f);
    *D_ = D; *K_ = K; *M_ = M; *ksub_ = ksub; *ng_ = ng;
    return 0;
}

extern "C" {

int b200nn_pq_create(b200nn_ctx_t ctx, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                     const int32_t* perm, const float* R, float clamp_threshold, b200nn_pq_t* out) {
    return pq_init(ctx, D, K, M, ksub, coarse, codebooks, perm, R, clamp_threshold, out);
}

// IVFOPQ::LoadModel byte format, opq/src/IVFOPQ.cpp:75-95
int b200nn_pq_load_model(b200nn_ctx_t ctx, const char* model_path, b200nn_pq_t* out) {
    if (!model_path) B2_FAIL(B200NN_ERR_INVALID, "pq_load_model: path is NULL");
    FILE* f = fopen(model_path, "rb");
    if (!f) B2_FAIL(B200NN_ERR_IO, "Can not open the model file!");  // the reference's message, IVFOPQ.cpp:71
    int32_t h[4];
    if (fread(h, 4, 4, f) != 4) { fclose(f); B2_FAIL(B200NN_ERR_IO, "pq_load_model: truncated header"); }
    const int D = h[0], K = h[1], M = h[2], ksub = h[3];
    if (D <= 0 || K <= 0 || M <= 0 || ksub <= 0 || D % M || (long long)K * D > (1LL << 32)) {
        fclose(f);
        B2_FAIL(B200NN_ERR_IO, "pq_load_model: implausible header");
    }
    std::vector<float> coarse((size_t)K * D), cb((size_t)M * ksub * (D / M));
    std::vector<int32_t> perm(D);
    bool ok = fread(coarse.data(), 4, coarse.size(), f) == coarse.size() && fread(cb.data(), 4, cb.size(), f) == cb.size() &&
              fread(perm.data(), 4, D, f) == (size_t)D;
    fclose(f);
    if (!ok) B2_FAIL(B200NN_ERR_IO, "pq_load_model: truncated model file");
    return pq_init(ctx, D, K, M, ksub, coarse.data(), cb.data(), perm.data(), nullptr, 1.0f, out);
}

void b200nn_pq_destroy(b200nn_pq_t idx) {
    if (!idx) return;
    {
        Guard g(idx);
        cudaStreamSynchronize(idx->ctx->c.stream);
    }
    delete idx;
}

int b200nn_pq_set_clamp(b200nn_pq_t idx, float clamp) {
    if (!idx) B2_FAIL(B200NN_ERR_INVALID, "pq is NULL");
    if (!(clamp >= 0.0f)) B2_FAIL(B200NN_ERR_INVALID, "clamp must be >= 0 (INFINITY disables it)");
    idx->clamp = clamp;
    return 0;
}

int b200nn_pq_info(b200nn_pq_t idx, int* D, int* K, int* M, int* ksub, uint64_t* n_rows, uint64_t* n_groups) {
    if (!idx) B2_FAIL(B200NN_ERR_INVALID, "pq is NULL");
    if (D) *D = idx->D;
    if (K) *K = idx->K;
    if (M) *M = idx->M;
    if (ksub) *ksub = idx->ksub;
    if (n_rows) *n_rows = (uint64_t)idx->n;
    if (n_groups) *n_groups = (uint64_t)idx->n_groups;
    return 0;
}

int b200nn_pq_rotate(b200nn_pq_t p, const float* x, size_t n, float* y) {
    if (!p || (n && (!x || !y))) B2_FAIL(B200NN_ERR_INVALID, "pq_rotate: NULL argument");
    if (!n) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = p->ws_qraw.ensure(n * p->D)) || (rc = p->ws_q.ensure(n * p->D))) return rc;
    B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, x, sizeof(float) * n * p->D, cudaMemcpyHostToDevice, c->stream));
    const float* r = nullptr;
    if ((rc = rotate_dev(p, p->ws_qraw.p, (long long)n, p->ws_q.p, &r))) return rc;
    B2_CUDA(cudaMemcpyAsync(y, r, sizeof(float) * n * p->D, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_pq_encode(b200nn_pq_t p, const float* x_rot, size_t n, int32_t* out_list, uint8_t* out_codes) {
    if (!p || (n && (!x_rot || !out_codes))) B2_FAIL(B200NN_ERR_INVALID, "pq_encode: NULL argument");
    if (!n) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = p->ws_qraw.ensure(n * p->D)) || (rc = p->ws_list.ensure(n)) || (rc = p->ws_codes.ensure(n * p->M))) return rc;
    B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, x_rot, sizeof(float) * n * p->D, cudaMemcpyHostToDevice, c->stream));
    if ((rc = encode_dev(p, p->ws_qraw.p, (long long)n, p->ws_list.p, p->ws_codes.p))) return rc;
    if (out_list) B2_CUDA(cudaMemcpyAsync(out_list, p->ws_list.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(out_codes, p->ws_codes.p, n * p->M, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int pq_add_host(b200nn_pq_t p, const float* x_raw, size_t n, const int32_t* group_ids, bool already_rotated);

int b200nn_pq_add(b200nn_pq_t p, const float* x_raw, size_t n, const int32_t* group_ids) {
    return pq_add_host(p, x_raw, n, group_ids, false);
}

int b200nn_pq_add_rotated(b200nn_pq_t p, const float* x_rotated, size_t n, const int32_t* group_ids) {
    return pq_add_host(p, x_rotated, n, group_ids, true);
}

// IVFOPQ::Add from host memory.  The rows go through two pinned staging buffers (kept by the context) and two device
// buffers: while the kernels of one block (rotate, coarse assign, encode) run on the context stream, the next block is
// copied into its pinned buffer by the host and sent ahead on a second stream -- the copies hide behind the encode instead
// of queueing in front of it (pageable, single-buffered, synchronous before: 0.75 s per 10^6 rows at K = 8192, 0.15 s of
// it kernels).
static int pq_add_host(b200nn_pq_t p, const float* x_raw, size_t n, const int32_t* group_ids, bool already_rotated) {
    if (!p || (n && !x_raw)) B2_FAIL(B200NN_ERR_INVALID, "pq_add: NULL argument");
    if (!n) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    const size_t blk = std::max<size_t>(1024, std::min<size_t>(n, (size_t)(32u << 20) / (sizeof(float) * p->D)));  // <= 32 MB per block
    int rc;
    if ((rc = ensure_copy_engine(c, blk * p->D * sizeof(float)))) return rc;
    DevBuf<float> stage[2];
    DevBuf<int> gstage[2];
    for (int b = 0; b < 2; b++) {
        if ((rc = stage[b].ensure(blk * p->D))) return rc;
        if (group_ids && (rc = gstage[b].ensure(blk))) return rc;
    }
    // grow the row store once, not per block
    if ((rc = p->codes.reserve((size_t)(p->n + n) * p->M, (size_t)p->n * p->M, c->stream)) ||
        (rc = p->list.reserve((size_t)(p->n + n), (size_t)p->n, c->stream)) || (rc = p->group.reserve((size_t)(p->n + n), (size_t)p->n, c->stream)))
        return rc;
    int b = 0;
    for (size_t off = 0; off < n; off += blk, b ^= 1) {
        const size_t cn = std::min(blk, n - off);
        // pinned[b] / stage[b] were last used by block (off - 2 blk): its copy and its kernels must be done
        B2_CUDA(cudaEventSynchronize(c->copy_done[b]));
        B2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->compute_done[b], 0));
        memcpy(c->pinned[b], x_raw + off * p->D, cn * p->D * sizeof(float));
        B2_CUDA(cudaMemcpyAsync(stage[b].p, c->pinned[b], cn * p->D * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
        if (group_ids) B2_CUDA(cudaMemcpyAsync(gstage[b].p, group_ids + off, sizeof(int) * cn, cudaMemcpyHostToDevice, c->copy_stream));
        B2_CUDA(cudaEventRecord(c->copy_done[b], c->copy_stream));
        B2_CUDA(cudaStreamWaitEvent(c->stream, c->copy_done[b], 0));
        if ((rc = add_dev_locked(p, stage[b].p, (long long)cn, group_ids ? gstage[b].p : nullptr, group_ids ? group_ids + off : nullptr,
                                 already_rotated)))
            return rc;
        B2_CUDA(cudaEventRecord(c->compute_done[b], c->stream));
    }
    B2_CUDA(cudaStreamSynchronize(c->copy_stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_pq_add_dev(b200nn_pq_t p, const float* x_raw_dev, size_t n, const int32_t* group_ids_dev) {
    if (!p || (n && !x_raw_dev)) B2_FAIL(B200NN_ERR_INVALID, "pq_add_dev: NULL argument");
    Guard g(p);
    return add_dev_locked(p, x_raw_dev, (long long)n, group_ids_dev, nullptr);
}

int b200nn_pq_get_rows(b200nn_pq_t p, uint64_t start, size_t n, int32_t* out_list, int32_t* out_group, uint8_t* out_codes) {
    if (!p) B2_FAIL(B200NN_ERR_INVALID, "pq is NULL");
    if (start + n > (uint64_t)p->n) B2_FAIL(B200NN_ERR_INVALID, "pq_get_rows: range exceeds the number of rows");
    if (!n) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    if (out_list) B2_CUDA(cudaMemcpyAsync(out_list, p->list.p + start, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    if (out_group) B2_CUDA(cudaMemcpyAsync(out_group, p->group.p + start, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    if (out_codes) B2_CUDA(cudaMemcpyAsync(out_codes, p->codes.p + start * p->M, n * p->M, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_pq_build_lut(b200nn_pq_t p, const float* q_rot, size_t nq, int nprobe, int32_t* out_lists, float* out_lut) {
    if (!p || (nq && (!q_rot || !out_lut))) B2_FAIL(B200NN_ERR_INVALID, "pq_build_lut: NULL argument");
    if (nprobe < 1 || nprobe > p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_build_lut: nprobe must be in [1, K]");
    if (!nq) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    const size_t lut_elems = nq * nprobe * p->M * p->ksub;
    if ((rc = p->ws_qraw.ensure(nq * p->D)) || (rc = p->ws_probes.ensure(nq * nprobe)) || (rc = p->ws_lut.ensure(lut_elems))) return rc;
    B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, q_rot, sizeof(float) * nq * p->D, cudaMemcpyHostToDevice, c->stream));
    if ((rc = probes_and_luts(p, p->ws_qraw.p, (long long)nq, nprobe, p->ws_probes.p, p->ws_lut.p))) return rc;
    if (out_lists) B2_CUDA(cudaMemcpyAsync(out_lists, p->ws_probes.p, sizeof(int) * nq * nprobe, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(out_lut, p->ws_lut.p, sizeof(float) * lut_elems, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// IVFOPQ::QueryThrehold (IVFOPQ.cpp:322-422) for nq raw DEVICE query rows: out[q][g] (row stride out_stride >= n_groups)
// is initialised to the clamp value and min-aggregated per group over the probed lists.  Workspaces must hold nq queries.
static int scores_dev_locked(b200nn_pq* p, const float* q_raw_dev, long long nq, int nprobe, long long out_stride, float* out_dev) {
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = ensure_csr(p))) return rc;
    if ((rc = p->ws_q.ensure((size_t)nq * p->D)) || (rc = p->ws_probes.ensure((size_t)nq * nprobe)) ||
        (rc = p->ws_lut.ensure((size_t)nq * nprobe * p->M * p->ksub)))
        return rc;
    const float* qr = nullptr;
    if ((rc = rotate_dev(p, q_raw_dev, nq, p->ws_q.p, &qr))) return rc;
    if ((rc = probes_and_luts(p, qr, nq, nprobe, p->ws_probes.p, p->ws_lut.p))) return rc;
    if ((rc = launch_fill_f32(c, out_dev, nq * out_stride, p->clamp))) return rc;
    if (p->n == 0) return 0;
    return launch_ivf_scan(c, p->ws_lut.p, p->ws_probes.p, p->list_off.p, p->codes_sorted.p, p->group_sorted.p, p->M, p->ksub, nq, nprobe,
                           out_stride, out_dev);
}

int b200nn_pq_scores_dev(b200nn_pq_t p, const float* q_raw_dev, size_t nq, int nprobe, size_t out_stride, float* out_scores_dev) {
    if (!p || (nq && (!q_raw_dev || !out_scores_dev))) B2_FAIL(B200NN_ERR_INVALID, "pq_scores_dev: NULL argument");
    if (nprobe < 1 || nprobe > p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_scores_dev: nprobe must be in [1, K]");
    if (out_stride < (size_t)p->n_groups) B2_FAIL(B200NN_ERR_INVALID, "pq_scores_dev: out_stride is smaller than the number of groups");
    if (!nq || !out_stride) return 0;
    Guard g(p);
    return scores_dev_locked(p, q_raw_dev, (long long)nq, nprobe, (long long)out_stride, out_scores_dev);
}

int b200nn_pq_scores(b200nn_pq_t p, const float* q_raw, size_t nq, int nprobe, float* out_scores) {
    if (!p || (nq && (!q_raw || !out_scores))) B2_FAIL(B200NN_ERR_INVALID, "pq_scores: NULL argument");
    if (nprobe < 1 || nprobe > p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_scores: nprobe must be in [1, K]");
    if (!nq || !p->n_groups) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    const long long ng = p->n_groups;
    const long long qc = std::max<long long>(1, std::min<long long>((long long)nq, std::min<long long>(1024, (1LL << 28) / ng)));
    if ((rc = p->ws_qraw.ensure((size_t)qc * p->D)) || (rc = p->ws_scores.ensure((size_t)qc * ng))) return rc;
    for (long long q0 = 0; q0 < (long long)nq; q0 += qc) {
        const long long cq = std::min<long long>(qc, (long long)nq - q0);
        B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, q_raw + q0 * p->D, sizeof(float) * cq * p->D, cudaMemcpyHostToDevice, c->stream));
        if ((rc = scores_dev_locked(p, p->ws_qraw.p, cq, nprobe, ng, p->ws_scores.p))) return rc;
        B2_CUDA(cudaMemcpyAsync(out_scores + q0 * ng, p->ws_scores.p, sizeof(float) * cq * ng, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// f-5: what the reference's query main does with Query's output (multi_frame_index_test.cpp:54-68): the frames
// of one query video are scored against every indexed video (QueryThrehold: clamp-initialised, min per videoId),
// the per-frame scores are summed per video and get_sort_results keeps the k smallest (score, videoId).
// Everything stays on the device; only [n_videos, k] results come back.
int b200nn_pq_query_groups(b200nn_pq_t p, const float* q_raw, const int64_t* frame_off, size_t n_videos, int nprobe, size_t k,
                           float* out_score, uint64_t* out_group) {
    if (!p || !frame_off || (n_videos && (!out_score || !out_group))) B2_FAIL(B200NN_ERR_INVALID, "pq_query_groups: NULL argument");
    if (nprobe < 1 || nprobe > p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_query_groups: nprobe must be in [1, K]");
    if (k < 1 || k > (size_t)KP) B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_query_groups: k must be in [1, 128]");
    for (size_t v = 0; v < n_videos; v++)
        if (frame_off[v + 1] < frame_off[v] || frame_off[0] != 0) B2_FAIL(B200NN_ERR_INVALID, "pq_query_groups: frame_off must ascend from 0");
    if (!n_videos) return 0;
    if (frame_off[n_videos] > 0 && !q_raw) B2_FAIL(B200NN_ERR_INVALID, "pq_query_groups: NULL queries");
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = ensure_csr(p))) return rc;
    const long long ng = p->n_groups;
    if (ng > 0xFFFFFFFFLL) B2_FAIL(B200NN_ERR_UNSUPPORTED, "pq_query_groups: more than 2^32 groups");
    // frames of one video are scored in chunks of fc frames; the running per-video total is folded in frame order
    const long long fc = std::max<long long>(1, std::min<long long>(256, (1LL << 28) / std::max<long long>(1, ng)));
    if ((rc = p->ws_qraw.ensure((size_t)fc * p->D)) || (rc = p->ws_q.ensure((size_t)fc * p->D)) ||
        (rc = p->ws_probes.ensure((size_t)fc * nprobe)) || (rc = p->ws_lut.ensure((size_t)fc * nprobe * p->M * p->ksub)) ||
        (rc = p->ws_scores.ensure((size_t)(fc + 1) * std::max<long long>(1, ng))) || (rc = p->ws_keys.ensure(n_videos * k)) ||
        (rc = p->ws_dist.ensure(n_videos * k)) || (rc = p->ws_id.ensure(n_videos * k)))
        return rc;
    float* total = p->ws_scores.p;            // row 0: running total (also frame_sum's "frame -1")
    float* frames = p->ws_scores.p + ng;      // rows 1..fc: this chunk's per-frame scores
    for (size_t v = 0; v < n_videos; v++) {
        const long long f0 = frame_off[v], f1 = frame_off[v + 1];
        if ((rc = launch_fill_f32(c, total, ng, 0.0f))) return rc;  // vector<float> score_total(img_num, 0.0f)
        for (long long fa = f0; fa < f1; fa += fc) {
            const long long cf = std::min<long long>(fc, f1 - fa);
            B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, q_raw + fa * p->D, sizeof(float) * cf * p->D, cudaMemcpyHostToDevice, c->stream));
            const float* qr = nullptr;
            if ((rc = rotate_dev(p, p->ws_qraw.p, cf, p->ws_q.p, &qr))) return rc;
            if ((rc = probes_and_luts(p, qr, cf, nprobe, p->ws_probes.p, p->ws_lut.p))) return rc;
            if ((rc = launch_fill_f32(c, frames, cf * ng, p->clamp))) return rc;
            if (ng && (rc = launch_ivf_scan(c, p->ws_lut.p, p->ws_probes.p, p->list_off.p, p->codes_sorted.p, p->group_sorted.p, p->M,
                                            p->ksub, cf, nprobe, ng, frames)))
                return rc;
            // total = (((total + s[0]) + s[1]) + ...): row 0 of the buffer is the running total, 0.0f + x == x exactly
            if ((rc = launch_frame_sum(c, total, (int)cf + 1, ng, total))) return rc;
        }
        if ((rc = launch_dense_topk(c, total, 1, ng, (int)k, p->ws_keys.p + v * k))) return rc;
    }
    if ((rc = launch_topk_merge(c, p->ws_keys.p, 1, (long long)n_videos, (int)k, (long long)(n_videos * k), p->ws_dist.p, nullptr,
                                p->ws_id.p, nullptr)))
        return rc;
    B2_CUDA(cudaMemcpyAsync(out_score, p->ws_dist.p, sizeof(float) * n_videos * k, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(out_group, p->ws_id.p, sizeof(uint64_t) * n_videos * k, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_pq_search_dev(b200nn_pq_t p, const float* q_raw_dev, size_t nq, int nprobe, size_t k, float* out_dist_dev,
                         uint64_t* out_id_dev, uint64_t* out_key_dev, uint64_t id_base) {
    if (!p || (nq && !q_raw_dev)) B2_FAIL(B200NN_ERR_INVALID, "pq_search_dev: NULL argument");
    Guard g(p);
    int rc = search_dev_locked(p, q_raw_dev, (long long)nq, nprobe, (int)k, out_dist_dev, (unsigned long long*)out_id_dev,
                               (unsigned long long*)out_key_dev, id_base);
    return rc;
}

int b200nn_pq_search(b200nn_pq_t p, const float* q_raw, size_t nq, int nprobe, size_t k, float* out_dist, uint64_t* out_id) {
    if (!p || (nq && (!q_raw || !out_dist || !out_id))) B2_FAIL(B200NN_ERR_INVALID, "pq_search: NULL argument");
    if (!nq) return 0;
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = p->ws_qraw.ensure(nq * p->D)) || (rc = p->ws_dist.ensure(nq * k)) || (rc = p->ws_id.ensure(nq * k))) return rc;
    B2_CUDA(cudaMemcpyAsync(p->ws_qraw.p, q_raw, sizeof(float) * nq * p->D, cudaMemcpyHostToDevice, c->stream));
    if ((rc = search_dev_locked(p, p->ws_qraw.p, (long long)nq, nprobe, (int)k, p->ws_dist.p, p->ws_id.p, nullptr, 0))) return rc;
    B2_CUDA(cudaMemcpyAsync(out_dist, p->ws_dist.p, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(out_id, p->ws_id.p, sizeof(uint64_t) * nq * k, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return check_dev_err(c, "pq_search");
}

int b200nn_pq_last_timing(b200nn_pq_t p, float* ms4) {
    if (!p || !ms4) B2_FAIL(B200NN_ERR_INVALID, "pq_last_timing: NULL argument");
    Guard g(p);
    Ctx* c = &p->ctx->c;
    B2_CUDA(cudaEventSynchronize(c->events[4]));
    for (int i = 0; i < 4; i++) B2_CUDA(cudaEventElapsedTime(&ms4[i], c->events[i], c->events[i + 1]));
    return 0;
}

int b200nn_pq_scan_bytes(b200nn_pq_t p, uint64_t* code_bytes) {
    if (!p || !code_bytes) B2_FAIL(B200NN_ERR_INVALID, "NULL argument");
    *code_bytes = (uint64_t)p->n * p->M;
    return 0;
}

// Host-only view of the fused scan's work plan (no device work): how a batch of `nq` queries over
// `n_rows` rows of M-byte codes is cut into CTAs on a GPU with `sm_count` SMs.
int b200nn_pq_scan_plan(int sm_count, int M, size_t nq, size_t n_rows, int* n_full, int* n_tail, int* slices, int32_t* desc,
                        size_t desc_capacity) {
    const int QW = scan_queries_per_cta(M);
    if (sm_count < 1 || QW < 1 || 128 % M) B2_FAIL(B200NN_ERR_INVALID, "pq_scan_plan: need sm_count >= 1 and M in {4, 8, 16, 32}");
    ScanPlan plan;
    scan_plan(sm_count, ((long long)nq + QW - 1) / QW, ((long long)n_rows + 63) / 64, &plan);
    if (n_full) *n_full = plan.n_full;
    if (n_tail) *n_tail = plan.n_tail;
    if (slices) *slices = plan.slices;
    if (desc) {
        if (desc_capacity < plan.desc.size()) B2_FAIL(B200NN_ERR_INVALID, "pq_scan_plan: desc buffer too small (8 ints per tail CTA)");
        std::copy(plan.desc.begin(), plan.desc.end(), desc);
    }
    return 0;
}

// IVFOPQ::SaveIndex byte format (IVFOPQ.cpp:541-580).  `dir_or_path` ending in
// ".fvecs" is used verbatim, otherwise the reference's file name is composed inside that directory.
int b200nn_pq_save_index(b200nn_pq_t p, const char* dir_or_path, const char* const* group_paths) {
    return b200nn_pq_save_index_n(p, dir_or_path, group_paths, group_paths ? (size_t)-1 : 0);
}

int b200nn_pq_save_index_n(b200nn_pq_t p, const char* dir_or_path, const char* const* group_paths, size_t n_paths) {
    if (!p || !dir_or_path) B2_FAIL(B200NN_ERR_INVALID, "pq_save_index: NULL argument");
    std::vector<float> coarse, cb;
    PQHostRows rows;
    int rc;
    if ((rc = pq_model_to_host(p, &coarse, &cb, nullptr)) || (rc = pq_rows_to_host(p, &rows))) return rc;
    return pq_write_index_file(pq_index_file_name(dir_or_path, p->n_groups, p->D, p->K, p->M, p->ksub), p->D, p->K, p->M, p->ksub,
                               p->n_groups, coarse.data(), cb.data(), (long long)rows.lists.size(), rows.lists.data(), rows.groups.data(),
                               rows.codes.data(), group_paths, n_paths);
}

// Append rows that are already coded (host pointers): coarse list id, videoId and M code bytes per row -- what an index
// file holds.  Every value is validated before it can index device memory.
int b200nn_pq_append_coded(b200nn_pq_t p, const int32_t* lists, const int32_t* group_ids, const uint8_t* codes, size_t n) {
    if (!p || (n && (!lists || !codes))) B2_FAIL(B200NN_ERR_INVALID, "pq_append_coded: NULL argument");
    if (!n) return 0;
    if ((unsigned long long)p->n + n > 0x7fffffffull) B2_FAIL(B200NN_ERR_STATE, "pq_append_coded: more than 2^31-1 rows per shard");
    long long ng = p->n_groups;
    for (size_t i = 0; i < n; i++) {
        if (lists[i] < 0 || lists[i] >= p->K) B2_FAIL(B200NN_ERR_INVALID, "pq_append_coded: coarse list id out of range");
        if (group_ids && group_ids[i] < 0) B2_FAIL(B200NN_ERR_INVALID, "pq_append_coded: negative group id");
        if (group_ids) ng = std::max<long long>(ng, (long long)group_ids[i] + 1);
        if (p->ksub < 256)
            for (int m = 0; m < p->M; m++)
                if (codes[i * p->M + m] >= p->ksub) B2_FAIL(B200NN_ERR_INVALID, "pq_append_coded: code byte >= ksub");
    }
    Guard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    const size_t n0 = (size_t)p->n;
    if ((rc = p->codes.reserve((n0 + n) * p->M, n0 * p->M, c->stream)) || (rc = p->list.reserve(n0 + n, n0, c->stream)) ||
        (rc = p->group.reserve(n0 + n, n0, c->stream)))
        return rc;
    B2_CUDA(cudaMemcpyAsync(p->codes.p + n0 * p->M, codes, n * p->M, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(p->list.p + n0, lists, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    if (group_ids) B2_CUDA(cudaMemcpyAsync(p->group.p + n0, group_ids, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    else {
        iota_kernel<<<small_grid((long long)n), 256, 0, c->stream>>>(p->group.p + n0, (long long)n, (int)n0);
        c->launches++;
        ng = std::max<long long>(ng, (long long)(n0 + n));
    }
    B2_CUDA(cudaStreamSynchronize(c->stream));
    p->n += (long long)n;
    p->n_groups = ng;
    p->codesT_rows = -1;
    p->csr_rows = -1;
    return 0;
}

// Symmetric reader of what SaveIndex writes (the reference's own LoadIndex cannot parse it,
// SURVEY.md App. D-2).  The file has no reorder tail, so perm comes from the caller (may be NULL).
int b200nn_pq_load_index(b200nn_ctx_t ctx, const char* path, const int32_t* perm, float clamp, b200nn_pq_t* out) {
    if (!ctx || !path || !out) B2_FAIL(B200NN_ERR_INVALID, "pq_load_index: NULL argument");
    *out = nullptr;
    int D, K, M, ksub;
    long long ng;
    std::vector<float> coarse, cb;
    PQHostRows rows;
    int rc;
    if ((rc = pq_parse_index_file(path, &D, &K, &M, &ksub, &ng, &coarse, &cb, &rows))) return rc;
    b200nn_pq* p = nullptr;
    if ((rc = pq_init(ctx, D, K, M, ksub, coarse.data(), cb.data(), perm, nullptr, clamp, &p))) return rc;
    if ((rc = b200nn_pq_append_coded(p, rows.lists.data(), rows.groups.data(), rows.codes.data(), rows.lists.size()))) {
        b200nn_pq_destroy(p);  // no half-built handle escapes
        return rc;
    }
    p->n_groups = std::max<long long>(p->n_groups, ng);
    *out = p;
    return 0;
}

}  // extern "C"
