// capi_train.cu -- C ABI of training: k-means for the coarse quantizer and the PQ codebooks
// (TrainPQ::CoarseQuan / ProdQuan, opq/train_codebook/train_PQ_codebook.cpp:150-244) and the model writer
// (TrainPQ::SaveCodebook, :247-288).
//
// The reference delegates to yael's kmeans (un-vendored, random initialisation => numerically unpinned).  The
// replacement is a deterministic Lloyd iteration, defined completely here so that a CPU checker can restate it bit for bit:
//   init     : a splitmix64 stream seeded with `seed` drives a partial Fisher-Yates shuffle of the row indices;
//              centroid j starts as row idx[j]                                              (needs n >= k)
//   assign   : argmin_j of the sequential fp32 squared distance, first minimum wins          (IVFOPQ.cpp:107-129)
//   stop     : when the assignment pass reproduces the partition the centroids were formed from, when every distance
//              is zero, or after max_iter updates (0 = 10000, like yael's "niter = 0 for convergence",
//              train_PQ_codebook.cpp:158); an assignment pass always comes last, so the returned assignment/distances
//              belong to the returned centroids
//   empties  : before the means are formed, empty clusters (ascending) each take the row with the largest distance to its
//              own centroid (ties: lowest row first) among the rows whose cluster keeps at least one other row
//   update   : rows of a cluster in ascending row order, summed in double in blocks of 512 rows, block sums added in
//              block order, centroid = (float)(sum / count)
// Host: the O(n) integer bookkeeping.  Device: every floating-point operation (train_kernels.cu).
#include <stdio.h>

#include <algorithm>
#include <numeric>
#include <string>

#include "capi_common.cuh"
#include "pq_kernels.cuh"
#include "train_kernels.cuh"

using namespace b200nn;

namespace {

inline uint64_t splitmix64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct KmWork {
    DevBuf<int> assign, row_sorted, count, donors, empties;
    DevBuf<float> dist, cT;
    DevBuf<long long> blk_lo, blk_hi, cl_off;
    DevBuf<double> partial;
};

// ---- host bookkeeping (also exported host-only so that CPU tests can drive it without a device) -------------------
// rows the k initial centroids are copied from
void km_init_rows(long long n, int k, uint64_t seed, std::vector<int>& out) {
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    out.resize(k);
    uint64_t s = seed;
    for (int j = 0; j < k; j++) {
        const long long r = j + (long long)(splitmix64(s) % (uint64_t)(n - j));
        std::swap(idx[j], idx[r]);
        out[j] = idx[j];
    }
}

// One update's integer work: cluster sizes, donors for empty clusters (assign is modified), stable counting sort.
// off[j]..off[j+1] = positions of cluster j in rows.  Returns 0, or a negative error code.
int km_plan_update(long long n, int k, int* assign, const float* dist, std::vector<int>& count, std::vector<int>& rows,
                   std::vector<long long>& off) {
    count.assign(k, 0);
    rows.resize(n);
    off.assign(k + 1, 0);
    for (long long i = 0; i < n; i++) {
        const int a = assign[i];
        if (a < 0 || a >= k) B2_FAIL(B200NN_ERR_INVALID, "kmeans: a row has no nearest centroid (NaN or Inf in the training data?)");
        count[a]++;
    }
    // empty clusters, ascending: each takes the farthest row (ties: lowest row) among the rows not taken yet whose cluster
    // keeps at least one other row; the row MOVES to the empty cluster before the means are formed.  Rows are visited in
    // (dist desc, row asc) order; a row skipped because it is alone in its cluster stays alone (sizes of non-empty
    // clusters only fall), so one pass over the k best candidates serves every empty cluster.
    int ne = 0;
    for (int j = 0; j < k; j++) ne += (count[j] == 0);
    if (ne) {
        const long long T = std::min<long long>(n, k);
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::partial_sort(order.begin(), order.begin() + T, order.end(),
                          [&](int a, int b) { return dist[a] > dist[b] || (dist[a] == dist[b] && a < b); });
        long long p = 0;
        for (int j = 0; j < k; j++) {
            if (count[j] != 0) continue;
            while (p < T && count[assign[order[p]]] < 2) p++;
            if (p >= T) B2_FAIL(B200NN_ERR_STATE, "kmeans: no donor row for an empty cluster");
            const int r = order[p++];
            count[assign[r]]--;
            assign[r] = j;
            count[j] = 1;
        }
    }
    for (int j = 0; j < k; j++) off[j + 1] = off[j] + count[j];
    std::vector<long long> pos(off.begin(), off.end() - 1);
    for (long long i = 0; i < n; i++) rows[pos[assign[i]]++] = (int)i;
    return 0;
}

// k-means over columns [col0, col0+d) of the device matrix x[n][ld]; centroids -> c_dev [k][d] (device, row-major).
// h_assign / h_dist receive the final assignment (host).
int kmeans_device(Ctx* c, KmWork& w, const float* x, long long ld, int col0, long long n, int d, int k, int max_iter, uint64_t seed,
                  float* c_dev, std::vector<int>& h_assign, std::vector<float>& h_dist, int* iters_out, double* mse_out) {
    if (n < k) B2_FAIL(B200NN_ERR_INVALID, "kmeans: need at least as many rows as centroids");
    if (n > 0x7fffffffLL) B2_FAIL(B200NN_ERR_UNSUPPORTED, "kmeans: at most 2^31-1 training rows");
    int rc;
    const long long max_blocks = n / KM_SUM_BLOCK + k + 1;
    if ((rc = w.assign.ensure(n)) || (rc = w.dist.ensure(n)) || (rc = w.row_sorted.ensure(n)) || (rc = w.count.ensure(k)) ||
        (rc = w.donors.ensure(k)) || (rc = w.empties.ensure(k)) || (rc = w.cT.ensure((size_t)k * d)) || (rc = w.blk_lo.ensure(max_blocks)) ||
        (rc = w.blk_hi.ensure(max_blocks)) || (rc = w.cl_off.ensure(k + 1)) || (rc = w.partial.ensure((size_t)max_blocks * d)))
        return rc;
    // ---- init
    std::vector<int> donors, empties(k);
    km_init_rows(n, k, seed, donors);
    std::iota(empties.begin(), empties.end(), 0);
    B2_CUDA(cudaMemcpyAsync(w.donors.p, donors.data(), sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(w.empties.p, empties.data(), sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_kmeans_reseed(c, x, ld, col0, d, w.donors.p, w.empties.p, k, k, c_dev, w.cT.p))) return rc;
    B2_CUDA(cudaStreamSynchronize(c->stream));  // donors/empties are reused below

    const int cap = max_iter > 0 ? max_iter : 10000;
    std::vector<int> prev, count, rows;
    std::vector<long long> off, cl_off(k + 1), blk_lo, blk_hi;
    h_assign.assign(n, -1);
    h_dist.assign(n, 0.0f);
    int iters = 0;
    for (;;) {
        if ((rc = launch_kmeans_assign(c, x, ld, col0, n, d, w.cT.p, k, w.assign.p, w.dist.p))) return rc;
        B2_CUDA(cudaMemcpyAsync(h_assign.data(), w.assign.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaMemcpyAsync(h_dist.data(), w.dist.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
        bool zero = true;
        for (long long i = 0; i < n && zero; i++) zero = (h_dist[i] == 0.0f);
        if ((iters > 0 && h_assign == prev) || zero) break;  // fixed point: the centroids are the means of this assignment
        if (iters == cap) break;
        if ((rc = km_plan_update(n, k, h_assign.data(), h_dist.data(), count, rows, off))) return rc;
        blk_lo.clear(); blk_hi.clear();
        for (int j = 0; j < k; j++) {
            cl_off[j] = (long long)blk_lo.size();
            for (long long q = off[j]; q < off[j + 1]; q += KM_SUM_BLOCK) {
                blk_lo.push_back(q);
                blk_hi.push_back(std::min<long long>(q + KM_SUM_BLOCK, off[j + 1]));
            }
        }
        cl_off[k] = (long long)blk_lo.size();
        const long long nb = (long long)blk_lo.size();
        B2_CUDA(cudaMemcpyAsync(w.row_sorted.p, rows.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaMemcpyAsync(w.count.p, count.data(), sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaMemcpyAsync(w.cl_off.p, cl_off.data(), sizeof(long long) * (k + 1), cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaMemcpyAsync(w.blk_lo.p, blk_lo.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaMemcpyAsync(w.blk_hi.p, blk_hi.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_kmeans_partial(c, x, ld, col0, d, w.row_sorted.p, w.blk_lo.p, w.blk_hi.p, nb, w.partial.p))) return rc;
        if ((rc = launch_kmeans_finalize(c, w.partial.p, w.cl_off.p, w.count.p, d, k, c_dev, w.cT.p))) return rc;
        B2_CUDA(cudaStreamSynchronize(c->stream));  // the host vectors above are rewritten in the next pass
        prev = h_assign;
        iters++;
    }
    double sum = 0.0;
    for (long long i = 0; i < n; i++) sum += (double)h_dist[i];
    if (iters_out) *iters_out = iters;
    if (mse_out) *mse_out = n ? sum / (double)n : 0.0;
    return 0;
}

}  // namespace

extern "C" {

int b200nn_kmeans(b200nn_ctx_t ctx, const float* x, size_t n, int d, int k, int max_iter, uint64_t seed, float* centroids,
                  int32_t* assign, float* dist, int* iters_done, double* mse) {
    if (!ctx || !x || !centroids) B2_FAIL(B200NN_ERR_INVALID, "kmeans: NULL argument");
    if (d < 1 || k < 1 || max_iter < 0) B2_FAIL(B200NN_ERR_INVALID, "kmeans: need d >= 1, k >= 1, max_iter >= 0");
    std::lock_guard<std::mutex> g(ctx->mu);
    Ctx* c = &ctx->c;
    B2_CUDA(cudaSetDevice(c->device));
    DevBuf<float> xd, cd;
    KmWork w;
    int rc;
    if ((rc = xd.ensure(std::max<size_t>(1, n * (size_t)d))) || (rc = cd.ensure((size_t)k * d))) return rc;
    B2_CUDA(cudaMemcpyAsync(xd.p, x, sizeof(float) * n * d, cudaMemcpyHostToDevice, c->stream));
    std::vector<int> ha;
    std::vector<float> hd;
    if ((rc = kmeans_device(c, w, xd.p, d, 0, (long long)n, d, k, max_iter, seed, cd.p, ha, hd, iters_done, mse))) return rc;
    B2_CUDA(cudaMemcpyAsync(centroids, cd.p, sizeof(float) * k * d, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    if (assign) memcpy(assign, ha.data(), sizeof(int32_t) * n);
    if (dist) memcpy(dist, hd.data(), sizeof(float) * n);
    return 0;
}

int b200nn_pq_train(b200nn_ctx_t ctx, const float* x_raw, size_t n, int D, int K, int M, int ksub, const int32_t* perm, int max_iter,
                    uint64_t seed, float* coarse, float* codebooks, double* mse_out) {
    if (!ctx || !x_raw || !coarse || !codebooks) B2_FAIL(B200NN_ERR_INVALID, "pq_train: NULL argument");
    if (D <= 0 || K < 0 || M <= 0 || ksub <= 0 || ksub > 256 || D % M != 0 || max_iter < 0)
        B2_FAIL(B200NN_ERR_INVALID, "pq_train: need D, M > 0, K >= 0, D % M == 0, 1 <= ksub <= 256, max_iter >= 0");
    if (perm) {
        std::vector<char> seen(D, 0);
        for (int i = 0; i < D; i++) {
            if (perm[i] < 0 || perm[i] >= D || seen[perm[i]]) B2_FAIL(B200NN_ERR_INVALID, "pq_train: perm is not a permutation of 0..D-1");
            seen[perm[i]] = 1;
        }
    }
    const int ds = D / M, Kc = std::max(K, 1);
    std::lock_guard<std::mutex> g(ctx->mu);
    Ctx* c = &ctx->c;
    B2_CUDA(cudaSetDevice(c->device));
    DevBuf<float> xraw, xr, res, coarse_d, cb_d;
    DevBuf<int> perm_d;
    KmWork w;
    int rc;
    const size_t elems = std::max<size_t>(1, n * (size_t)D);
    if ((rc = xraw.ensure(elems)) || (rc = coarse_d.ensure((size_t)Kc * D)) || (rc = cb_d.ensure((size_t)M * ksub * ds))) return rc;
    B2_CUDA(cudaMemcpyAsync(xraw.p, x_raw, sizeof(float) * n * D, cudaMemcpyHostToDevice, c->stream));
    const float* x = xraw.p;
    if (perm) {  // LoadFeatureSample applies reorder_ to every training row (train_PQ_codebook.cpp:80,98,112)
        if ((rc = xr.ensure(elems)) || (rc = perm_d.ensure(D))) return rc;
        B2_CUDA(cudaMemcpyAsync(perm_d.p, perm, sizeof(int) * D, cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_rotate_perm(c, xraw.p, (long long)n, D, perm_d.p, xr.p))) return rc;
        x = xr.p;
    }
    std::vector<int> ha;
    std::vector<float> hd;
    double mse = 0.0;
    const float* resid = x;
    if (K >= 1) {  // CoarseQuan (:150-199): k-means over the full vectors, then the residue of every row
        if ((rc = kmeans_device(c, w, x, D, 0, (long long)n, D, K, max_iter, seed, coarse_d.p, ha, hd, nullptr, &mse))) return rc;
        if ((rc = res.ensure(elems))) return rc;
        if ((rc = launch_residual(c, x, (long long)n, D, coarse_d.p, w.assign.p, res.p))) return rc;
        resid = res.p;
    } else {  // K == 0: no coarse quantizer -- the flat-ADC model (one all-zero centroid)
        B2_CUDA(cudaMemsetAsync(coarse_d.p, 0, sizeof(float) * D, c->stream));
    }
    if (mse_out) mse_out[0] = mse;
    for (int m = 0; m < M; m++) {  // ProdQuan (:201-244): one k-means per sub-space of the residue
        if ((rc = kmeans_device(c, w, resid, D, m * ds, (long long)n, ds, ksub, max_iter, seed + 1 + (uint64_t)m,
                                cb_d.p + (size_t)m * ksub * ds, ha, hd, nullptr, &mse)))
            return rc;
        if (mse_out) mse_out[1 + m] = mse;
    }
    B2_CUDA(cudaMemcpyAsync(coarse, coarse_d.p, sizeof(float) * Kc * D, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(codebooks, cb_d.p, sizeof(float) * M * ksub * ds, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int b200nn_kmeans_init_rows(size_t n, int k, uint64_t seed, int32_t* rows_out) {
    if (!rows_out || k < 1 || n < (size_t)k || n > 0x7fffffffull) B2_FAIL(B200NN_ERR_INVALID, "kmeans_init_rows: need 1 <= k <= n < 2^31");
    std::vector<int> r;
    km_init_rows((long long)n, k, seed, r);
    memcpy(rows_out, r.data(), sizeof(int32_t) * k);
    return 0;
}

int b200nn_kmeans_plan_update(size_t n, int k, int32_t* assign, const float* dist, int32_t* count, int32_t* row_sorted, int64_t* cluster_off) {
    if (!assign || !dist || !count || !row_sorted || !cluster_off || k < 1 || n < (size_t)k || n > 0x7fffffffull)
        B2_FAIL(B200NN_ERR_INVALID, "kmeans_plan_update: NULL argument or not 1 <= k <= n < 2^31");
    std::vector<int> cnt, rows;
    std::vector<long long> off;
    const int rc = km_plan_update((long long)n, k, assign, dist, cnt, rows, off);
    if (rc) return rc;
    memcpy(count, cnt.data(), sizeof(int32_t) * k);
    memcpy(row_sorted, rows.data(), sizeof(int32_t) * n);
    for (int j = 0; j <= k; j++) cluster_off[j] = off[j];
    return 0;
}

int b200nn_pq_write_model(const char* path, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                          const int32_t* perm) {
    if (!path || !coarse || !codebooks) B2_FAIL(B200NN_ERR_INVALID, "pq_write_model: NULL argument");
    if (D <= 0 || K <= 0 || M <= 0 || ksub <= 0 || D % M != 0) B2_FAIL(B200NN_ERR_INVALID, "pq_write_model: bad shape");
    FILE* f = fopen(path, "wb");
    if (!f) B2_FAIL(B200NN_ERR_IO, std::string("pq_write_model: cannot open ") + path);
    const int32_t hdr[4] = {D, K, M, ksub};
    std::vector<int32_t> p(D);
    for (int i = 0; i < D; i++) p[i] = perm ? perm[i] : i;
    bool ok = fwrite(hdr, sizeof(int32_t), 4, f) == 4 && fwrite(coarse, sizeof(float), (size_t)K * D, f) == (size_t)K * D &&
              fwrite(codebooks, sizeof(float), (size_t)M * ksub * (D / M), f) == (size_t)M * ksub * (D / M) &&
              fwrite(p.data(), sizeof(int32_t), D, f) == (size_t)D;
    ok = (fclose(f) == 0) && ok;
    if (!ok) B2_FAIL(B200NN_ERR_IO, std::string("pq_write_model: short write to ") + path);
    return 0;
}

}  // extern "C"
