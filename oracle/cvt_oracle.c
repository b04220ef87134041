/*
 * cvt_oracle.c -- CPU restatement of the reference's quantized nearest-neighbour path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cvt_b200/, include/, tools/) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker.
 *
 * Every function restates one reference routine, scalar, in the reference's own operation
 * order, and cites the file:line it follows (paths relative to /root/reference).  It is
 * compiled with `gcc -O2 -ffp-contract=off` and NO -march/-mfma/-ffast-math, i.e. the same
 * arithmetic as the canonical `g++ -O2 -std=c++11` build of the reference on x86-64: every
 * fp32 add / sub / mul / div is a separately rounded IEEE operation.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - opq_* functions: pinned against the UNMODIFIED reference compiled in place
 *     (oracle/_ref/ref_opq, built from /root/reference/opq/src/IVFOPQ.cpp) on the shipped model +
 *     fixtures and on synthetic data, and against tests/golden/.
 *   - flat_* functions: pinned against the reference's hnswlib headers compiled in place
 *     (oracle/_ref/ref_flat) on synthetic data, and against tests/golden/.
 *   - sq_l2normalize / sq_encode / sq_decode: pinned against the UNMODIFIED reference class
 *     (oracle/_ref/ref_int8_quan = scalar_quantization/scalar_quantization/int8_quan.cc compiled in
 *     place against stand-ins for its two un-vendored dependencies, oracle/stubs/): the arithmetic of
 *     L2NormalizeVector, Int8Encode and Int8Decode(std::string&) is the reference's own and reads
 *     only sq.trained / sq.code_size from faiss.  Goldens tests/golden/sq_d*.npz are reference-run.
 *   - sq_decode_faiss, sq_train_minmax: PARITY UNPINNED -- they restate faiss 1.5.3 itself
 *     (ScalarQuantizer::decode, RS_minmax training), an un-vendored dependency; no faiss build and
 *     no trained model exist here, so the published algorithm is all there is to follow.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * opq/  (IVFOPQ)
 * ---------------------------------------------------------------------------------------- */

/* IVFOPQ::reorder, opq/src/IVFOPQ.cpp:424-439 (applied to every row by LoadSingleFeatFile
 * :459-461):  y[i] = x[reorder_[i]]. */
ORC_API void orc_opq_reorder(const float* x, int64_t n, int D, const int32_t* perm, float* y) {
    for (int64_t r = 0; r < n; r++)
        for (int i = 0; i < D; i++) y[r * D + i] = x[r * D + perm[i]];
}

/* Dense-rotation extension (the reference's "rotation" is always a permutation): y = R x with a
 * sequential fp32 dot product per output, R row-major [D][D].  Used to bound the tcgen05 GEMM. */
ORC_API void orc_opq_rotate_dense(const float* x, int64_t n, int D, const float* R, float* y) {
    for (int64_t r = 0; r < n; r++)
        for (int i = 0; i < D; i++) {
            float acc = 0.0f;
            for (int j = 0; j < D; j++) acc += R[(int64_t)i * D + j] * x[r * D + j];
            y[r * D + i] = acc;
        }
}

/* Squared distance exactly as written at IVFOPQ.cpp:117-122 / :147-154 / :283-288:
 * acc = 0.0f; for k ascending { tmp = a[k]-b[k]; acc += tmp*tmp; }  (two roundings per term). */
static inline float sqdist_seq(const float* a, const float* b, int d) {
    float acc = 0.0f;
    for (int k = 0; k < d; k++) {
        float tmp = a[k] - b[k];
        acc += tmp * tmp;
    }
    return acc;
}

/* IVFOPQ::Add coarse assignment, IVFOPQ.cpp:107-129.  dismin starts at (float)UINT_MAX
 * (= 4294967296.0f), strict '<' so the lowest index wins ties, vw stays -1 if nothing wins. */
ORC_API void orc_opq_coarse_assign(const float* x, int64_t n, int D, const float* coarse, int K,
                                   int32_t* out_list) {
    for (int64_t f = 0; f < n; f++) {
        int vw = -1;
        float dismin = (float)4294967295u;
        for (int i = 0; i < K; i++) {
            float distmp = sqdist_seq(x + f * D, coarse + (int64_t)i * D, D);
            if (distmp < dismin) {
                dismin = distmp;
                vw = i;
            }
        }
        out_list[f] = vw;
    }
}

/* IVFOPQ::Add residual + PQ argmin encode, IVFOPQ.cpp:135-163.  cb is [M][ksub][D/M]
 * (m_prodQuantizer as loaded at :88-93).  `list` gives the coarse centroid per row. */
ORC_API void orc_opq_pq_encode(const float* x, int64_t n, int D, const float* coarse,
                               const int32_t* list, const float* cb, int M, int ksub,
                               uint8_t* codes) {
    int step = D / M;
    float* res = (float*)malloc(sizeof(float) * (size_t)D);
    for (int64_t f = 0; f < n; f++) {
        const float* c = coarse + (int64_t)list[f] * D;
        for (int i = 0; i < D; i++) res[i] = x[f * D + i] - c[i];
        for (int i = 0; i < M; i++) {
            float dismin1 = (float)4294967295u;
            int vw1 = -1;
            for (int j = 0; j < ksub; j++) {
                float d = sqdist_seq(res + i * step, cb + ((int64_t)i * ksub + j) * step, step);
                if (d < dismin1) {
                    dismin1 = d;
                    vw1 = j;
                }
            }
            codes[f * M + i] = (uint8_t)vw1; /* elem.PQindex[i] = vw1 (uchar), :161 */
        }
    }
    free(res);
}

/* std::priority_queue<std::pair<float,int>> restated: binary max-heap ordered by
 * std::pair's operator< (lexicographic: first, then second). */
typedef struct { float d; int64_t i; } orc_pair;
static inline int pair_less(orc_pair a, orc_pair b) {
    if (a.d < b.d) return 1;
    if (b.d < a.d) return 0;
    return a.i < b.i;
}
static void heap_push(orc_pair* h, int* n, orc_pair v) {
    int i = (*n)++;
    h[i] = v;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (pair_less(h[p], h[i])) { orc_pair t = h[p]; h[p] = h[i]; h[i] = t; i = p; }
        else break;
    }
}
static void heap_pop(orc_pair* h, int* n) {
    h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && pair_less(h[m], h[l])) m = l;
        if (r < *n && pair_less(h[m], h[r])) m = r;
        if (m == i) break;
        orc_pair t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
}

/* IVFOPQ::Query coarse probe selection, IVFOPQ.cpp:238-260 (= QueryThrehold :346-368):
 * first nk pushed unconditionally, then replace top iff dis < top.first (strict).
 * out_lists receives the nk list ids in the order the reference pops them (:266-267,
 * largest (dist,id) first). */
ORC_API void orc_opq_coarse_probe(const float* q, int D, const float* coarse, int K, int nk,
                                  int32_t* out_lists) {
    orc_pair* h = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(nk + 1));
    int hn = 0;
    for (int i = 0; i < K; i++) {
        float dis = sqdist_seq(q, coarse + (int64_t)i * D, D);
        orc_pair v = {dis, i};
        if (i < nk) heap_push(h, &hn, v);
        else if (dis < h[0].d) { heap_pop(h, &hn); heap_push(h, &hn, v); }
    }
    for (int k = 0; k < nk; k++) { out_lists[k] = (int32_t)h[0].i; heap_pop(h, &hn); }
    free(h);
}

/* IVFOPQ::Query residual + LUT build, IVFOPQ.cpp:273-291 (= :380-398).
 * lut is [M][ksub]:  lut[m][j] = sum_k (res[m*step+k] - cb[m][j][k])^2, sequential fp32. */
ORC_API void orc_opq_build_lut(const float* q, int D, const float* centroid, const float* cb,
                               int M, int ksub, float* lut) {
    int step = D / M;
    float* res = (float*)malloc(sizeof(float) * (size_t)D);
    for (int i = 0; i < D; i++) res[i] = q[i] - centroid[i];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < ksub; j++)
            lut[i * ksub + j] = sqdist_seq(res + i * step, cb + ((int64_t)i * ksub + j) * step, step);
    free(res);
}

/* ADC score of one stored element, IVFOPQ.cpp:302-306 (= :405-409):
 * score = 0.0f; for k < M: score += PQ_table[k][code[k]]  (sequential fp32). */
static inline float adc_score(const float* lut, int M, int ksub, const uint8_t* code) {
    float score = 0.0f;
    for (int k = 0; k < M; k++) score += lut[k * ksub + code[k]];
    return score;
}

/* Flat ADC scan of n codes against one LUT (the inner loop of IVFOPQ.cpp:300-309 without the
 * min-aggregate): scores[j] = adc_score(code_j). */
ORC_API void orc_opq_adc_scan(const float* lut, int M, int ksub, const uint8_t* codes, int64_t n,
                              float* scores) {
    for (int64_t j = 0; j < n; j++) scores[j] = adc_score(lut, M, ksub, codes + j * M);
}

/* Whole IVFOPQ::QueryThrehold (IVFOPQ.cpp:322-422) for nq already-reordered query rows over an
 * index given as: per stored row its coarse list id, group id (videoId) and M-byte code.
 * match is [nq][n_groups], pre-filled with `clamp` (threhold = 1.0, :5,:369) and min-updated
 * (:410).  Elements of a list are visited in insertion order (irrelevant for min). */
ORC_API void orc_opq_query_scores(const float* q, int64_t nq, int D, const float* coarse, int K,
                                  const float* cb, int M, int ksub, int nk,
                                  const int32_t* row_list, const int32_t* row_group,
                                  const uint8_t* codes, int64_t n, int64_t n_groups, float clamp,
                                  float* match) {
    float* lut = (float*)malloc(sizeof(float) * (size_t)M * ksub);
    int32_t* probes = (int32_t*)malloc(sizeof(int32_t) * (size_t)nk);
    for (int64_t f = 0; f < nq; f++) {
        float* ms = match + f * n_groups;
        for (int64_t g = 0; g < n_groups; g++) ms[g] = clamp;
        orc_opq_coarse_probe(q + f * D, D, coarse, K, nk, probes);
        for (int p = 0; p < nk; p++) {
            int vw = probes[p];
            orc_opq_build_lut(q + f * D, D, coarse + (int64_t)vw * D, cb, M, ksub, lut);
            for (int64_t j = 0; j < n; j++) {
                if (row_list[j] != vw) continue;
                float score = adc_score(lut, M, ksub, codes + j * M);
                float prev = ms[row_group[j]];
                ms[row_group[j]] = score < prev ? score : prev; /* std::min(score, prev) */
            }
        }
    }
    free(lut);
    free(probes);
}

/* get_sort_results, opq/src/common.h:25-37: std::partial_sort_copy of pair<float,uint>
 * ascending => the k smallest under lexicographic (score, id), ascending.  O(n*k) insertion
 * is fine for an oracle. */
ORC_API void orc_topk_pairs(const float* score, int64_t n, int k, float* out_score,
                            int64_t* out_id) {
    int cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        orc_pair v = {score[i], i};
        if (cnt == k) {
            orc_pair last = {out_score[k - 1], out_id[k - 1]};
            if (!pair_less(v, last)) continue;
        }
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0) {
            orc_pair pv = {out_score[pos - 1], out_id[pos - 1]};
            if (pair_less(v, pv)) { out_score[pos] = pv.d; out_id[pos] = pv.i; pos--; }
            else break;
        }
        out_score[pos] = v.d;
        out_id[pos] = v.i;
        if (cnt < k) cnt++;
    }
    for (int j = cnt; j < k; j++) { out_score[j] = INFINITY; out_id[j] = -1; }
}

/* ------------------------------------------------------------------------------------------
 * brute_force_search/ + hnsw_sifts_retrieval/hnswlib/  (exact scan)
 * ---------------------------------------------------------------------------------------- */

/* The reference's SIMD distance kernels keep L lane accumulators; lane l sums elements
 * l, l+L, l+2L ... in order (mul then add, no FMA) and the lanes are summed left to right:
 *   L=1: InnerProduct / L2Sqr scalar loops        space_ip.hpp:25-34, space_l2.h:26-37
 *   L=4: SSE paths of *SIMD4Ext / *SIMD16Ext      space_ip.hpp:82-130,168-206, space_l2.h:82-119,123-151
 *   L=8: AVX path of *SIMD16Ext (dim%16==0)       space_ip.hpp:140-166, space_l2.h:46-70
 * The brute-force CLI's own flags (brute_force_search/src/CMakeLists.txt:4: no -mavx) select L=4. */
ORC_API float orc_flat_ip(const float* a, const float* b, int64_t d, int L) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < d; i++) acc[i % L] += a[i] * b[i];
    float sum = acc[0];
    for (int l = 1; l < L; l++) sum = sum + acc[l];
    return 1.0f - sum;
}
ORC_API float orc_flat_l2(const float* a, const float* b, int64_t d, int L) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < d; i++) {
        float t = a[i] - b[i];
        acc[i % L] += t * t;
    }
    float sum = acc[0];
    for (int l = 1; l < L; l++) sum = sum + acc[l];
    return sum;
}
/* L2SqrI, hnsw_sifts_retrieval/hnswlib/space_l2.h:186-219: exact int32, processes (d>>2)*4 elements. */
ORC_API int32_t orc_flat_l2_u8(const uint8_t* a, const uint8_t* b, int64_t d) {
    int32_t res = 0;
    int64_t n4 = d >> 2;
    for (int64_t i = 0; i < n4 * 4; i++) {
        int t = (int)a[i] - (int)b[i];
        res += t * t;
    }
    return res;
}

/* BruteforceSearch<dist_t>::searchKnn, brute_force_search/src/brutoforce.hpp:73-93, literally:
 * first k rows pushed, then push iff dist <= lastdist, pop when size > k.  Output is the heap
 * drained and reversed (ascending), as brute_force.cpp:89-101 does.
 * metric: 0 = 1-IP, 1 = L2 (both float, lanes L), 2 = L2SqrI over uint8 (dist returned as float-
 * exact int32 in out_dist_i).  Requires n >= k (the reference reads garbage otherwise). */
ORC_API void orc_flat_search(int metric, int L, const void* data, const uint64_t* labels,
                             int64_t n, int64_t d, const void* query, int k, float* out_dist_f,
                             int32_t* out_dist_i, uint64_t* out_label) {
    orc_pair* h = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(k + 2));
    int hn = 0;
    /* int distances are carried in the float field only for ordering when metric==2 would lose
     * precision above 2^24, so keep a parallel exact copy. */
    typedef struct { int64_t di; int64_t lab; } ipair;
    ipair* hi = NULL;
    int hin = 0;
    if (metric == 2) hi = (ipair*)malloc(sizeof(ipair) * (size_t)(k + 2));
#define IPLESS(a, b) ((a).di < (b).di || ((a).di == (b).di && (a).lab < (b).lab))
    if (metric != 2) {
        const float* X = (const float*)data;
        const float* Q = (const float*)query;
        for (int64_t i = 0; i < n; i++) {
            float dist = metric == 0 ? orc_flat_ip(Q, X + i * d, d, L) : orc_flat_l2(Q, X + i * d, d, L);
            orc_pair v = {dist, (int64_t)labels[i]};
            if (i < k) { heap_push(h, &hn, v); continue; }
            float lastdist = h[0].d;
            if (dist <= lastdist) {
                heap_push(h, &hn, v);
                if (hn > k) heap_pop(h, &hn);
            }
        }
        for (int j = hn - 1; j >= 0; j--) {
            out_dist_f[j] = h[0].d;
            out_label[j] = (uint64_t)h[0].i;
            heap_pop(h, &hn);
        }
    } else {
        const uint8_t* X = (const uint8_t*)data;
        const uint8_t* Q = (const uint8_t*)query;
        for (int64_t i = 0; i < n; i++) {
            ipair v = {orc_flat_l2_u8(Q, X + i * d, d), (int64_t)labels[i]};
            int take = 0;
            if (i < k) take = 1;
            else {
                /* top = max element */
                int mx = 0;
                for (int t = 1; t < hin; t++) if (IPLESS(hi[mx], hi[t])) mx = t;
                if (v.di <= hi[mx].di) take = 1;
            }
            if (take) {
                hi[hin++] = v;
                if (hin > k) {
                    int mx = 0;
                    for (int t = 1; t < hin; t++) if (IPLESS(hi[mx], hi[t])) mx = t;
                    hi[mx] = hi[--hin];
                }
            }
        }
        /* ascending selection sort */
        for (int a = 0; a < hin; a++) {
            int mn = a;
            for (int t = a + 1; t < hin; t++) if (IPLESS(hi[t], hi[mn])) mn = t;
            ipair tmp = hi[a]; hi[a] = hi[mn]; hi[mn] = tmp;
            out_dist_i[a] = (int32_t)hi[a].di;
            out_label[a] = (uint64_t)hi[a].lab;
        }
        free(hi);
    }
#undef IPLESS
    free(h);
}

/* ------------------------------------------------------------------------------------------
 * scalar_quantization/  (Int8Quan)  -- reference-run except the two faiss-internal routines (see the header).
 * ---------------------------------------------------------------------------------------- */

/* Int8Quan::L2NormalizeVector, scalar_quantization/scalar_quantization/int8_quan.cc:46-56:
 * product in float, accumulated in double, sqrt in double, denominator narrowed to float. */
ORC_API void orc_sq_l2normalize(float* v, int d) {
    double accum = 0.0;
    for (int i = 0; i < d; ++i) accum += v[i] * v[i];
    accum = sqrt(accum);
    float denorm_v = (float)(1e-12 < accum ? accum : 1e-12); /* std::max((double)1e-12, (double)accum) */
    for (int i = 0; i < d; ++i) v[i] = v[i] / denorm_v;
}

/* Int8Quan::Int8Encode, int8_quan.cc:72-94 (the reference's own restatement of faiss 1.5.3
 * QT_8bit non-uniform encode).  Encodes ONE d-dim vector; x is normalised in place when l2norm. */
ORC_API void orc_sq_encode(float* x, uint8_t* bytes, int d, const float* vmin, const float* vdiff,
                           int l2norm) {
    if (l2norm) orc_sq_l2normalize(x, d);
    for (int i = 0; i < d; i++) {
        float xi = 0;
        if (vdiff[i] != 0) xi = (x[i] - vmin[i]) / vdiff[i];
        if (xi < 0) xi = 0;
        if (xi > 1.0) xi = 1.0;
        bytes[i] = (uint8_t)(int)(255 * xi);
    }
}

/* Int8Quan::Int8Decode(std::string&, float*), int8_quan.cc:117-132:
 * x = vmin + vdiff * (byte + 0.5) / 255.0 evaluated in double, rounded once to float. */
ORC_API void orc_sq_decode(const uint8_t* bytes, float* x, int d, const float* vmin,
                           const float* vdiff) {
    for (int i = 0; i < d; ++i) x[i] = vmin[i] + vdiff[i] * (bytes[i] + 0.5) / 255.0;
}

/* faiss 1.5.3 ScalarQuantizer QT_8bit non-uniform decode (reached via Int8Decode(uint8_t*) /
 * Int8DecodeFaiss, int8_quan.cc:96-115): Codec8bit::decode_component = (code + 0.5f) / 255.0f,
 * reconstruct = vmin + xi * vdiff, all fp32. */
ORC_API void orc_sq_decode_faiss(const uint8_t* bytes, float* x, int d, const float* vmin,
                                 const float* vdiff) {
    for (int i = 0; i < d; ++i) {
        float xi = (bytes[i] + 0.5f) / 255.0f;
        x[i] = vmin[i] + xi * vdiff[i];
    }
}

/* faiss 1.5.3 RS_minmax training with rangestat_arg = 0 (sq_train.cpp:100-101; semantics
 * confirmed by the reference's own print check :105-132): vmin = per-dim min, vdiff = max - min. */
ORC_API void orc_sq_train_minmax(const float* x, int64_t n, int d, float* vmin, float* vdiff) {
    for (int j = 0; j < d; j++) {
        float lo = HUGE_VALF, hi = -HUGE_VALF;
        for (int64_t i = 0; i < n; i++) {
            float v = x[i * d + j];
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
        vmin[j] = lo;
        vdiff[j] = hi - lo;
    }
}

/* ------------------------------------------------------------------------------------------
 * front end: the steps that PRODUCE the vectors fed to the path.
 * Both call into OpenCV (cv::PCA::project, cv::reduce, cv::normalize), an un-vendored dependency:
 * pinned against cv2 4.13 (Python, this container) outputs under tests/golden/ -- OpenCV's gemm
 * accumulates float products in double (GEMMSingleMul<float,double>), which is restated here; its SIMD
 * reduction orders are not, hence a tolerance (a few ulp) instead of bit equality at this boundary.
 * ---------------------------------------------------------------------------------------- */

/* cvtk::PCAUtils::reduceDim, pca_train_project/pca_online/pca_utils.cc:25-35:
 *   pca_.project(mat, reduceMat)      == (x - mean) * eigenvectors^T  (cv::PCA::project: subtract, then gemm)
 *   per row: normMat = row * row.t(); denomv = max(1e-12, (double)sqrt(normMat.at<float>(0,0))); row /= denomv
 * vectors = cv::PCA::eigenvectors, row-major [N][K]; mean [K]. */
ORC_API void orc_pca_project(const float* x, int64_t n, int K, const float* mean, const float* vectors, int N, int l2norm, float* y) {
    float* t = (float*)malloc(sizeof(float) * (size_t)K);
    for (int64_t r = 0; r < n; r++) {
        for (int k = 0; k < K; k++) t[k] = mean ? x[r * K + k] - mean[k] : x[r * K + k];
        float* yr = y + r * N;
        for (int j = 0; j < N; j++) {
            double acc = 0.0;
            for (int k = 0; k < K; k++) acc += (double)t[k] * (double)vectors[(int64_t)j * K + k];
            yr[j] = (float)acc;
        }
        if (l2norm) {
            double s = 0.0;
            for (int j = 0; j < N; j++) s += (double)yr[j] * (double)yr[j];
            const float nrm2 = (float)s;
            const double d = (double)sqrtf(nrm2);
            const float denomv = (float)(d > 1e-12 ? d : 1e-12);
            for (int j = 0; j < N; j++) yr[j] = yr[j] / denomv;
        }
    }
    free(t);
}

/* siftsIDX::rootSift, hnsw_sifts_retrieval/siftsIndex.cpp:54-71 (same code makeSIFTs.cpp:79-95), eps = 1e-7:
 *   d = abs(d); sums = reduce(d, SUM over columns, CV_32F); d = sqrt(d / (sums + eps)); normalize(row, NORM_L2)
 * cv::normalize(NORM_L2, alpha = 1): scale = 1 / norm(row) with the norm accumulated in double, row *= scale. */
ORC_API void orc_rootsift(float* x, int64_t n, int d, float eps) {
    for (int64_t r = 0; r < n; r++) {
        float* v = x + r * d;
        double sum = 0.0;
        for (int j = 0; j < d; j++) { v[j] = fabsf(v[j]); sum += (double)v[j]; }
        const float sums = (float)sum;
        double s2 = 0.0;
        for (int j = 0; j < d; j++) { v[j] = sqrtf(v[j] / (sums + eps)); s2 += (double)v[j] * (double)v[j]; }
        const double nrm = sqrt(s2);
        const double scale = nrm > 2.220446049250313e-16 ? 1.0 / nrm : 0.0;  /* cv::normalize: DBL_EPSILON guard */
        for (int j = 0; j < d; j++) v[j] = (float)((double)v[j] * scale);
    }
}

/* ------------------------------------------------------------------------------------------
 * training.  The reference trains with yael's kmeans
 * (opq/train_codebook/train_PQ_codebook.cpp:164,229), an un-vendored dependency with random initialisation:
 * PARITY UNPINNED at that boundary.  What is restated here is the product's own deterministic Lloyd iteration
 * (cvt_b200/csrc/capi_train.cu states it), operation by operation, so that the device result can be checked bit
 * for bit; the reference-shaped part is the arithmetic of the assignment (IVFOPQ.cpp:107-129, sqdist_seq above)
 * and the structure CoarseQuan -> residue -> ProdQuan (train_PQ_codebook.cpp:150-244).
 * ---------------------------------------------------------------------------------------- */
#define ORC_KM_SUM_BLOCK 512

static inline uint64_t orc_splitmix64(uint64_t* s) {
    *s += 0x9E3779B97F4A7C15ull;
    uint64_t z = *s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* k-means over columns [col0, col0+d) of x[n][ld].  centroids [k][d]; assign [n], dist [n] (both required).
 * On return assign/dist belong to the returned centroids.  Returns the number of updates performed, or -1 on bad input. */
ORC_API int orc_kmeans(const float* x, int64_t n, int64_t ld, int col0, int d, int k, int max_iter, uint64_t seed,
                       float* centroids, int32_t* assign, float* dist, double* mse) {
    if (n < k || k < 1 || d < 1) return -1;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    double* sum = (double*)malloc(sizeof(double) * (size_t)k * d);
    double* part = (double*)malloc(sizeof(double) * (size_t)k * d);
    int64_t* inblk = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    char* taken = (char*)malloc((size_t)n);
    float* row = (float*)malloc(sizeof(float) * (size_t)d);
    for (int64_t i = 0; i < n; i++) idx[i] = (int32_t)i;
    uint64_t s = seed;
    for (int j = 0; j < k; j++) { /* partial Fisher-Yates */
        const int64_t r = j + (int64_t)(orc_splitmix64(&s) % (uint64_t)(n - j));
        const int32_t t = idx[j]; idx[j] = idx[r]; idx[r] = t;
        for (int t2 = 0; t2 < d; t2++) centroids[(int64_t)j * d + t2] = x[(int64_t)idx[j] * ld + col0 + t2];
    }
    const int cap = max_iter > 0 ? max_iter : 10000;
    int iters = 0;
    for (;;) {
        for (int64_t i = 0; i < n; i++) { /* assignment: IVFOPQ.cpp:107-129 arithmetic */
            for (int t = 0; t < d; t++) row[t] = x[i * ld + col0 + t];
            int best = -1;
            float dismin = (float)4294967295u;
            for (int j = 0; j < k; j++) {
                const float dd = sqdist_seq(row, centroids + (int64_t)j * d, d);
                if (dd < dismin) { dismin = dd; best = j; }
            }
            assign[i] = best;
            dist[i] = dismin;
        }
        int same = iters > 0, zero = 1;
        for (int64_t i = 0; i < n; i++)
            if (dist[i] != 0.0f) { zero = 0; break; }
        if (same)
            for (int64_t i = 0; i < n; i++)
                if (assign[i] != prev[i]) { same = 0; break; }
        if (same || zero || iters == cap) break;
        for (int j = 0; j < k; j++) { cnt[j] = 0; inblk[j] = 0; }
        for (int64_t i = 0; i < n; i++) {
            if (assign[i] < 0) { iters = -1; goto done; }
            cnt[assign[i]]++;
        }
        /* empty clusters, ascending: each takes the farthest row (ties: lowest row) among the rows not taken yet whose
         * cluster keeps at least one other row; the row MOVES to the empty cluster before the means are formed */
        memset(taken, 0, (size_t)n);
        for (int j = 0; j < k; j++) {
            if (cnt[j] != 0) continue;
            int64_t far = -1;
            for (int64_t i = 0; i < n; i++)
                if (!taken[i] && cnt[assign[i]] >= 2 && (far < 0 || dist[i] > dist[far])) far = i;
            taken[far] = 1;
            cnt[assign[far]]--;
            assign[far] = j;
            cnt[j] = 1;
        }
        /* update: per cluster, rows in ascending order, blocks of 512 rows summed in double, block sums added in order */
        for (int64_t e = 0; e < (int64_t)k * d; e++) { sum[e] = 0.0; part[e] = 0.0; }
        for (int64_t i = 0; i < n; i++) {
            const int a = assign[i];
            for (int t = 0; t < d; t++) part[(int64_t)a * d + t] += (double)x[i * ld + col0 + t];
            if (++inblk[a] == ORC_KM_SUM_BLOCK) {
                for (int t = 0; t < d; t++) { sum[(int64_t)a * d + t] += part[(int64_t)a * d + t]; part[(int64_t)a * d + t] = 0.0; }
                inblk[a] = 0;
            }
        }
        for (int j = 0; j < k; j++)
            for (int t = 0; t < d; t++) {
                double sj = sum[(int64_t)j * d + t];
                if (inblk[j]) sj += part[(int64_t)j * d + t];
                centroids[(int64_t)j * d + t] = (float)(sj / (double)cnt[j]);
            }
        memcpy(prev, assign, sizeof(int32_t) * (size_t)n);
        iters++;
    }
    if (mse) {
        double tot = 0.0;
        for (int64_t i = 0; i < n; i++) tot += (double)dist[i];
        *mse = n ? tot / (double)n : 0.0;
    }
done:
    free(idx); free(prev); free(cnt); free(sum); free(part); free(inblk); free(taken); free(row);
    return iters;
}

/* TrainPQ::IFVPQ: LoadFeatureSample's reorder (:80,98,112), CoarseQuan (:150-199), ProdQuan (:201-244).
 * K == 0: no coarse quantizer (one zero centroid).  mse_out: 1+M doubles or NULL. */
ORC_API int orc_pq_train(const float* x_raw, int64_t n, int D, int K, int M, int ksub, const int32_t* perm, int max_iter,
                         uint64_t seed, float* coarse, float* codebooks, double* mse_out) {
    const int ds = D / M;
    float* x = (float*)malloc(sizeof(float) * (size_t)n * D);
    float* res = (float*)malloc(sizeof(float) * (size_t)n * D);
    int32_t* assign = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* dist = (float*)malloc(sizeof(float) * (size_t)n);
    int rc = 0;
    double mse = 0.0;
    for (int64_t r = 0; r < n; r++)
        for (int i = 0; i < D; i++) x[r * D + i] = x_raw[r * D + (perm ? perm[i] : i)];
    if (K >= 1) {
        if (orc_kmeans(x, n, D, 0, D, K, max_iter, seed, coarse, assign, dist, &mse) < 0) { rc = -1; goto out; }
        for (int64_t r = 0; r < n; r++)
            for (int i = 0; i < D; i++) res[r * D + i] = x[r * D + i] - coarse[(int64_t)assign[r] * D + i];
    } else {
        for (int i = 0; i < D; i++) coarse[i] = 0.0f;
        memcpy(res, x, sizeof(float) * (size_t)n * D);
    }
    if (mse_out) mse_out[0] = mse;
    for (int m = 0; m < M; m++) {
        if (orc_kmeans(res, n, D, m * ds, ds, ksub, max_iter, seed + 1 + (uint64_t)m, codebooks + (int64_t)m * ksub * ds, assign, dist, &mse) < 0) {
            rc = -1;
            goto out;
        }
        if (mse_out) mse_out[1 + m] = mse;
    }
out:
    free(x); free(res); free(assign); free(dist);
    return rc;
}
