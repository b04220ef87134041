// capi_flat.cu -- C ABI of the exact flat index: hnswlib::BruteforceSearch<float|int> drop-in
// (brute_force_search/src/brutoforce.hpp:8-136; identical copy hnsw_sifts_retrieval/hnswlib/brutoforce.h).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <unordered_map>

#include "capi_common.cuh"
#include "flat_kernels.cuh"
#include "topk.cuh"
#include "u8_scan_tc.cuh"

using namespace b200nn;

struct b200nn_flat {
    b200nn_ctx* ctx = nullptr;
    int metric = 0, order = 4;
    size_t dim = 0, max_elements = 0, row_bytes = 0;
    size_t n = 0;
    DevBuf<unsigned char> data;                 // [max_elements][row_bytes]
    std::vector<uint64_t> labels;               // host mirror, row order
    std::unordered_map<uint64_t, size_t> l2r;   // dict_external_to_internal, brutoforce.hpp:41
    bool ranks_valid = false;
    DevBuf<uint32_t> rank;                      // rank of each row's label
    DevBuf<unsigned long long> label_sorted;    // rank -> label
    // tensor-core scan layout of u8 rows (metric 2): canonical 256-row tiles + |x|^2
    long long tc_rows = -1;
    DevBuf<unsigned char> xcan;
    DevBuf<int> xnorm;   // per-tile row meta: |x|^2 and label rank
    DevBuf<int> ws_gmin;  // [nq][32] list minima of the one-pass shared-bound scan
    DevBuf<int> ws_thr;  // [nq][k] distances of the sample pass (column k-1 = the bound the full pass starts from)
    DevBuf<unsigned char> ws_q;
    DevBuf<unsigned long long> ws_keys, ws_id, ws_best;  // ws_best: the best k of the passes so far (sorted keys)
    DevBuf<float> ws_dist;
};

namespace {

struct FGuard {
    std::lock_guard<std::mutex> g;
    explicit FGuard(b200nn_flat* p) : g(p->ctx->mu) { cudaSetDevice(p->ctx->c.device); }
};

int flat_new(b200nn_ctx_t ctx, int metric, int order, size_t dim, size_t max_elements, b200nn_flat_t* out) {
    if (!ctx || !out) B2_FAIL(B200NN_ERR_INVALID, "flat_create: NULL argument");
    if (metric < 0 || metric > 2 || dim == 0) B2_FAIL(B200NN_ERR_INVALID, "flat_create: bad metric or dim");
    if (metric != 2) {
        if (order != 1 && order != 4 && order != 8) B2_FAIL(B200NN_ERR_INVALID, "flat_create: order must be 1, 4 or 8");
        // the reference picks the kernel from the dimension (space_ip.hpp:217-225, space_l2.h:159-164)
        if (order == 8 && dim % 16 != 0) B2_FAIL(B200NN_ERR_UNSUPPORTED, "flat_create: AVX order (8) needs dim % 16 == 0");
        if (order == 4 && dim % 4 != 0) B2_FAIL(B200NN_ERR_UNSUPPORTED, "flat_create: SSE order (4) needs dim % 4 == 0");
    }
    if (max_elements >= 0xFFFFFFFFull) B2_FAIL(B200NN_ERR_UNSUPPORTED, "flat_create: at most 2^32-2 rows per index");
    b200nn_flat* p = new b200nn_flat();
    p->ctx = ctx; p->metric = metric; p->order = metric == 2 ? 0 : order; p->dim = dim; p->max_elements = max_elements;
    p->row_bytes = metric == 2 ? dim : dim * sizeof(float);
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->c.device);
    int rc = p->data.ensure(std::max<size_t>(1, max_elements) * p->row_bytes);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

int ensure_ranks(b200nn_flat* p) {
    if (p->ranks_valid) return 0;
    Ctx* c = &p->ctx->c;
    const size_t n = p->n;
    std::vector<uint32_t> order(n), rank(n);
    std::iota(order.begin(), order.end(), 0u);
    bool sorted = true;
    for (size_t i = 1; i < n && sorted; i++) sorted = p->labels[i - 1] < p->labels[i];
    if (!sorted) std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return p->labels[a] < p->labels[b]; });
    std::vector<unsigned long long> ls(n);
    for (size_t r = 0; r < n; r++) { rank[order[r]] = (uint32_t)r; ls[r] = p->labels[order[r]]; }
    int rc;
    if ((rc = p->rank.ensure(std::max<size_t>(1, n))) || (rc = p->label_sorted.ensure(std::max<size_t>(1, n)))) return rc;
    if (n) {
        B2_CUDA(cudaMemcpyAsync(p->rank.p, rank.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaMemcpyAsync(p->label_sorted.p, ls.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    p->ranks_valid = true;
    return 0;
}

int search_dev_locked(b200nn_flat* p, const void* q_dev, size_t nq, size_t k, void* out_dist, unsigned long long* out_label) {
    Ctx* c = &p->ctx->c;
    if (!nq) return 0;
    if (k < 1 || k > (size_t)KP) B2_FAIL(B200NN_ERR_UNSUPPORTED, "flat_search: k must be in [1, 128]");
    int rc;
    if ((rc = ensure_ranks(p))) return rc;
    if (p->metric == 2 && u8_scan_tc_supported((int)p->dim, (int)k) && !getenv("B200NN_NO_TC_U8")) {
        // u8 x u8 -> s32 contraction on the tensor cores (tcgen05 kind::i8) with the fused top-k epilogue
        if (p->tc_rows != (long long)p->n) {
            const long long n_pad = std::max<long long>(256, ((long long)p->n + 255) / 256 * 256);
            if ((rc = p->xcan.ensure((size_t)n_pad * p->dim)) || (rc = p->xnorm.ensure((size_t)n_pad * 2))) return rc;
            if ((rc = launch_u8_rows_to_canonical(c, p->data.p, p->rank.p, (long long)p->n, (int)p->dim, p->xcan.p, p->xnorm.p, n_pad))) return rc;
            p->tc_rows = (long long)p->n;
        }
        const int groups = u8_scan_tc_lists_per_slice((int)p->dim, (int)k);
        const long long tiles = ((long long)p->n + 255) / 256;
        // Every (slice, group) list pays its own top-k warm-up, and a chunk of accumulators that holds a survivor costs the
        // epilogue far more than a clean one does (a dependent chain of shared-memory loads per survivor).  A large index is
        // therefore scanned in passes of growing size and each pass starts from the exact k-th best distance of everything
        // before it (an upper bound on the final k-th best; ties are kept, so the result is exact): by the last pass, which
        // holds 7/8 of the rows, the bound sits at a selectivity of k / 131072.
        long long bounds[5] = {0, tiles, tiles, tiles, tiles};
        int n_pass = 1;
        // One pass with a shared bound (u8_scan_tc.cu, the bound warp) when a query has at least k lists to take minima from:
        // no pass structure, no intermediate merges, the lists live through the whole scan.
        const int S_all = u8_scan_tc_slices(c->sm_count, (long long)nq, tiles, 4);
        const bool shared_bound = tiles >= 1024 && (int)k <= std::min(u8_scan_tc_bound_lists(), S_all * groups) && !getenv("B200NN_U8_NO_SHARED_BOUND");
        if (shared_bound) {
            if ((rc = p->ws_gmin.ensure(nq * (size_t)u8_scan_tc_bound_lists()))) return rc;
            B2_CUDA(cudaMemsetAsync(p->ws_gmin.p, 0x7f, nq * (size_t)u8_scan_tc_bound_lists() * sizeof(int), c->stream));
        } else if (tiles >= 1024 && !getenv("B200NN_NO_U8_SAMPLE")) {
            // 64 tiles (16 k rows), 8 x as many, then the rest.  (A finer schedule -- 16 / 100 / 625 tiles / rest, the first pass
            // one tile per CTA -- was measured: 0.403 against 0.396 ms for cfg2; every pass has ~15 us of fixed cost.)
            bounds[1] = 64; bounds[2] = 512; bounds[3] = tiles;
            n_pass = 3;
        }
        auto pass_slices = [&](int ps) { return u8_scan_tc_slices(c->sm_count, (long long)nq, bounds[ps + 1] - bounds[ps], 4); };
        int Lmax = 0;
        for (int ps = 0; ps < n_pass; ps++) Lmax = std::max(Lmax, pass_slices(ps) * groups + 1);
        if ((rc = p->ws_keys.ensure((size_t)Lmax * nq * k))) return rc;
        if (n_pass > 1 && ((rc = p->ws_thr.ensure(nq * k)) || (rc = p->ws_best.ensure(nq * k)))) return rc;
        for (int ps = 0; ps < n_pass; ps++) {
            const long long t0 = bounds[ps], nt = bounds[ps + 1] - bounds[ps];
            const int S = pass_slices(ps);
            int L = S * groups;
            if (nt <= 0) B2_CUDA(cudaMemsetAsync(p->ws_keys.p, 0xFF, (size_t)L * nq * k * sizeof(unsigned long long), c->stream));  // empty index
            if ((rc = launch_u8_scan_tc(c, p->xcan.p, p->xnorm.p, (long long)p->n, t0, nt, (int)p->dim, (const unsigned char*)q_dev, (long long)nq, S,
                                        (int)k, ps ? p->ws_thr.p + (k - 1) : nullptr, (int)k, shared_bound ? p->ws_gmin.p : nullptr, p->ws_keys.p)))
                return rc;
            if (ps) {  // the best k of the earlier passes ride along as one more list
                B2_CUDA(cudaMemcpyAsync(p->ws_keys.p + (size_t)L * nq * k, p->ws_best.p, nq * k * sizeof(unsigned long long), cudaMemcpyDeviceToDevice,
                                        c->stream));
                L++;
            }
            const bool last = ps == n_pass - 1;
            if ((rc = launch_topk_merge(c, p->ws_keys.p, L, (long long)nq, (int)k, (long long)(nq * k), nullptr, last ? (int*)out_dist : p->ws_thr.p,
                                        last ? out_label : nullptr, last ? nullptr : p->ws_best.p)))
                return rc;
        }
        return launch_rank_to_label(c, out_label, (long long)(nq * k), p->label_sorted.p);
    }
    if (p->metric != 2) {  // fp32: register-tiled all-pairs distances in the reference's accumulation order + selection
        long long cr;
        int nc, S;
        flat_f32_plan(c->sm_count, (long long)nq, (long long)p->n, &cr, &nc, &S);
        const int L = nc * S;
        if ((rc = p->ws_keys.ensure((size_t)L * nq * k))) return rc;
        if ((rc = launch_flat_scan_f32(c, p->metric, p->order, (const float*)p->data.p, p->rank.p, (long long)p->n, (int)p->dim, (const float*)q_dev,
                                       (long long)nq, (int)k, p->ws_keys.p)))
            return rc;
        if ((rc = launch_topk_merge(c, p->ws_keys.p, L, (long long)nq, (int)k, (long long)(nq * k), (float*)out_dist, nullptr, out_label, nullptr)))
            return rc;
        return launch_rank_to_label(c, out_label, (long long)(nq * k), p->label_sorted.p);
    }
    const int S = flat_pick_slices(c->sm_count, (long long)nq, (long long)p->n);
    if ((rc = p->ws_keys.ensure((size_t)S * nq * k))) return rc;
    if ((rc = launch_flat_scan(c, p->metric, p->order, p->data.p, p->rank.p, (long long)p->n, (int)p->dim, q_dev, (long long)nq, S,
                               (int)k, p->ws_keys.p)))
        return rc;
    if ((rc = launch_topk_merge(c, p->ws_keys.p, S, (long long)nq, (int)k, (long long)(nq * k), nullptr, (int*)out_dist, out_label, nullptr)))
        return rc;
    return launch_rank_to_label(c, out_label, (long long)(nq * k), p->label_sorted.p);
}

}  // namespace

extern "C" {

int b200nn_flat_create(b200nn_ctx_t ctx, int metric, int order, size_t dim, size_t max_elements, b200nn_flat_t* out) {
    return flat_new(ctx, metric, order, dim, max_elements, out);
}

void b200nn_flat_destroy(b200nn_flat_t p) {
    if (!p) return;
    {
        FGuard g(p);
        cudaStreamSynchronize(p->ctx->c.stream);
    }
    delete p;
}

int b200nn_flat_add(b200nn_flat_t p, const void* vectors, const uint64_t* labels, size_t n) {
    if (!p || (n && (!vectors || !labels))) B2_FAIL(B200NN_ERR_INVALID, "flat_add: NULL argument");
    if (!n) return 0;
    FGuard g(p);
    // validate the whole batch first, in addPoint's own order of checks per element
    // (duplicate label, then capacity; brutoforce.hpp:44-50); nothing is added if any row fails
    {
        std::unordered_map<uint64_t, size_t> seen;
        for (size_t i = 0; i < n; i++) {
            if (p->l2r.count(labels[i]) || !seen.emplace(labels[i], i).second) B2_FAIL(B200NN_ERR_STATE, "Ids have to be unique");
            if (p->n + i >= p->max_elements) B2_FAIL(B200NN_ERR_STATE, "The number of elements exceeds the specified limit\n");
        }
    }
    Ctx* c = &p->ctx->c;
    B2_CUDA(cudaMemcpyAsync(p->data.p + p->n * p->row_bytes, vectors, n * p->row_bytes, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) {
        p->l2r[labels[i]] = p->n + i;
        p->labels.push_back(labels[i]);
    }
    p->n += n;
    p->ranks_valid = false;
    p->tc_rows = -1;
    return 0;
}

// removePoint, brutoforce.hpp:58-70: the last element moves into the hole.
int b200nn_flat_remove(b200nn_flat_t p, uint64_t label) {
    if (!p) B2_FAIL(B200NN_ERR_INVALID, "flat is NULL");
    FGuard g(p);
    auto it = p->l2r.find(label);
    if (it == p->l2r.end()) B2_FAIL(B200NN_ERR_STATE, "flat_remove: label not found");
    const size_t cur = it->second, last = p->n - 1;
    Ctx* c = &p->ctx->c;
    p->l2r.erase(it);
    if (cur != last) {
        B2_CUDA(cudaMemcpyAsync(p->data.p + cur * p->row_bytes, p->data.p + last * p->row_bytes, p->row_bytes, cudaMemcpyDeviceToDevice, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
        p->labels[cur] = p->labels[last];
        p->l2r[p->labels[cur]] = cur;
    }
    p->labels.pop_back();
    p->n--;
    p->ranks_valid = false;
    p->tc_rows = -1;
    return 0;
}

int b200nn_flat_size(b200nn_flat_t p, size_t* out) {
    if (!p || !out) B2_FAIL(B200NN_ERR_INVALID, "NULL argument");
    *out = p->n;
    return 0;
}

int b200nn_flat_search_dev(b200nn_flat_t p, const void* q_dev, size_t nq, size_t k, void* out_dist_dev, uint64_t* out_label_dev) {
    if (!p || (nq && (!q_dev || !out_dist_dev || !out_label_dev))) B2_FAIL(B200NN_ERR_INVALID, "flat_search_dev: NULL argument");
    FGuard g(p);
    return search_dev_locked(p, q_dev, nq, k, out_dist_dev, (unsigned long long*)out_label_dev);
}

int b200nn_flat_search(b200nn_flat_t p, const void* queries, size_t nq, size_t k, void* out_dist, uint64_t* out_label) {
    if (!p || (nq && (!queries || !out_dist || !out_label))) B2_FAIL(B200NN_ERR_INVALID, "flat_search: NULL argument");
    if (!nq) return 0;
    FGuard g(p);
    Ctx* c = &p->ctx->c;
    int rc;
    if ((rc = p->ws_q.ensure(nq * p->row_bytes)) || (rc = p->ws_dist.ensure(nq * k)) || (rc = p->ws_id.ensure(nq * k))) return rc;
    B2_CUDA(cudaMemcpyAsync(p->ws_q.p, queries, nq * p->row_bytes, cudaMemcpyHostToDevice, c->stream));
    if ((rc = search_dev_locked(p, p->ws_q.p, nq, k, p->ws_dist.p, p->ws_id.p))) return rc;
    B2_CUDA(cudaMemcpyAsync(out_dist, p->ws_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(out_label, p->ws_id.p, nq * k * 8, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// saveIndex byte format, brutoforce.hpp:95-106: size_t maxelements_, size_per_element_, cur_element_count,
// then maxelements_ x [vector bytes | size_t label].  Unused slots are zero (the reference leaves
// them uninitialised).
int b200nn_flat_save(b200nn_flat_t p, const char* path) {
    if (!p || !path) B2_FAIL(B200NN_ERR_INVALID, "flat_save: NULL argument");
    FGuard g(p);
    Ctx* c = &p->ctx->c;
    std::vector<unsigned char> rows(p->n * p->row_bytes);
    if (p->n) {
        B2_CUDA(cudaMemcpyAsync(rows.data(), p->data.p, rows.size(), cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    FILE* f = fopen(path, "wb");
    if (!f) B2_FAIL(B200NN_ERR_IO, std::string("flat_save: cannot open ") + path);
    const size_t spe = p->row_bytes + sizeof(size_t);
    const size_t hdr[3] = {p->max_elements, spe, p->n};
    fwrite(hdr, sizeof(size_t), 3, f);
    std::vector<unsigned char> el(spe, 0);
    for (size_t i = 0; i < p->max_elements; i++) {
        if (i < p->n) {
            memcpy(el.data(), rows.data() + i * p->row_bytes, p->row_bytes);
            const size_t lab = (size_t)p->labels[i];
            memcpy(el.data() + p->row_bytes, &lab, sizeof(size_t));
        } else {
            memset(el.data(), 0, spe);
        }
        fwrite(el.data(), 1, spe, f);
    }
    const bool ok = !ferror(f);
    fclose(f);
    if (!ok) B2_FAIL(B200NN_ERR_IO, "flat_save: write failed");
    return 0;
}

// loadIndex, brutoforce.hpp:108-134: sizes are re-derived from the space (metric, dim), as the reference does.
int b200nn_flat_load(b200nn_ctx_t ctx, int metric, int order, size_t dim, const char* path, b200nn_flat_t* out) {
    if (!ctx || !path || !out) B2_FAIL(B200NN_ERR_INVALID, "flat_load: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) B2_FAIL(B200NN_ERR_IO, std::string("flat_load: cannot open ") + path);
    size_t hdr[3];
    if (fread(hdr, sizeof(size_t), 3, f) != 3) { fclose(f); B2_FAIL(B200NN_ERR_IO, "flat_load: truncated header"); }
    const size_t row_bytes = metric == 2 ? dim : dim * 4, spe = row_bytes + sizeof(size_t);
    if (hdr[1] != spe || hdr[2] > hdr[0]) { fclose(f); B2_FAIL(B200NN_ERR_IO, "flat_load: element size does not match metric/dim"); }
    int rc = flat_new(ctx, metric, order, dim, hdr[0], out);
    if (rc) { fclose(f); return rc; }
    const size_t n = hdr[2];
    std::vector<unsigned char> rows(n * row_bytes), el(spe);
    std::vector<uint64_t> labels(n);
    for (size_t i = 0; i < n; i++) {
        if (fread(el.data(), 1, spe, f) != spe) { fclose(f); b200nn_flat_destroy(*out); *out = nullptr; B2_FAIL(B200NN_ERR_IO, "flat_load: truncated data"); }
        memcpy(rows.data() + i * row_bytes, el.data(), row_bytes);
        size_t lab;
        memcpy(&lab, el.data() + row_bytes, sizeof(size_t));
        labels[i] = lab;
    }
    fclose(f);
    rc = b200nn_flat_add(*out, rows.data(), labels.data(), n);
    if (rc) { b200nn_flat_destroy(*out); *out = nullptr; }
    return rc;
}

// The index file of hnswlib::HierarchicalNSW<dist_t>::saveIndex (hnsw_sifts_retrieval/hnswlib/hnswalg.h:491-519), the file
// hnsw_sifts_retrieval/makeSearch.cpp:19-22 / siftsIndex.cpp:5-7 open: header {size_t offsetLevel0, max_elements,
// cur_element_count, size_data_per_element, label_offset, offsetData; int maxlevel; unsigned enterpoint; size_t maxM, maxM0,
// M; double mult; size_t ef_construction}, then max_elements x size_data_per_element bytes of level-0 memory where element i
// is [links | vector at offsetData | size_t label at label_offset]; the upper-level link lists follow and are skipped --
// the exact scan needs the vectors and labels only.  The graph is NOT rebuilt: searchKnn over this index is exact.
int b200nn_flat_load_hnsw(b200nn_ctx_t ctx, int metric, int order, size_t dim, const char* path, b200nn_flat_t* out) {
    if (!ctx || !path || !out) B2_FAIL(B200NN_ERR_INVALID, "flat_load_hnsw: NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) B2_FAIL(B200NN_ERR_IO, std::string("flat_load_hnsw: cannot open ") + path);
    auto bad = [&](const char* msg) { fclose(f); B2_FAIL(B200NN_ERR_IO, msg); };
    size_t h6[6];
    int maxlevel;
    unsigned enter;
    size_t m3[3], efc;
    double mult;
    if (fread(h6, sizeof(size_t), 6, f) != 6 || fread(&maxlevel, 4, 1, f) != 1 || fread(&enter, 4, 1, f) != 1 ||
        fread(m3, sizeof(size_t), 3, f) != 3 || fread(&mult, 8, 1, f) != 1 || fread(&efc, sizeof(size_t), 1, f) != 1)
        return bad("flat_load_hnsw: truncated header");
    const size_t off0 = h6[0], max_el = h6[1], cur = h6[2], spe = h6[3], label_off = h6[4], off_data = h6[5];
    const size_t row_bytes = metric == 2 ? dim : dim * 4;
    // hnswalg.h:39-44: size_links_level0 = maxM0*4 + 4; size_data_per_element = links + data + label; offsetData = links
    if (off0 != 0 || cur > max_el || label_off != off_data + row_bytes || spe != label_off + sizeof(size_t) ||
        off_data != m3[1] * sizeof(unsigned) + sizeof(unsigned) || max_el >= 0xFFFFFFFFull)
        return bad("flat_load_hnsw: not a HierarchicalNSW index of this metric/dimension");
    int rc = flat_new(ctx, metric, order, dim, max_el, out);
    if (rc) { fclose(f); return rc; }
    const size_t blk = 16384;
    std::vector<unsigned char> raw(blk * spe), rows(blk * row_bytes);
    std::vector<uint64_t> labels(blk);
    for (size_t i0 = 0; i0 < cur; i0 += blk) {
        const size_t cn = std::min(blk, cur - i0);
        if (fread(raw.data(), spe, cn, f) != cn) { b200nn_flat_destroy(*out); *out = nullptr; return bad("flat_load_hnsw: truncated level-0 memory"); }
        for (size_t i = 0; i < cn; i++) {
            memcpy(rows.data() + i * row_bytes, raw.data() + i * spe + off_data, row_bytes);
            size_t lab;
            memcpy(&lab, raw.data() + i * spe + label_off, sizeof(size_t));
            labels[i] = lab;
        }
        if ((rc = b200nn_flat_add(*out, rows.data(), labels.data(), cn))) {
            fclose(f);
            b200nn_flat_destroy(*out);
            *out = nullptr;
            return rc;
        }
    }
    fclose(f);
    return 0;
}

/* capacity and labels of an index (row order) -- what a loaded index has to tell its host-side wrapper */
int b200nn_flat_info(b200nn_flat_t p, size_t* max_elements, size_t* n, uint64_t* labels_out, size_t labels_capacity) {
    if (!p) B2_FAIL(B200NN_ERR_INVALID, "flat is NULL");
    FGuard g(p);
    if (max_elements) *max_elements = p->max_elements;
    if (n) *n = p->n;
    if (labels_out) {
        if (labels_capacity < p->n) B2_FAIL(B200NN_ERR_INVALID, "flat_info: labels buffer too small");
        std::copy(p->labels.begin(), p->labels.end(), labels_out);
    }
    return 0;
}

}  // extern "C"
