// topk_merge.cu -- cross-slice / cross-shard merge of sorted top-k key lists.
#include "topk.cuh"

namespace b200nn {

// ---------------------------------------------------------------------------------------------
// Merge of L sorted key lists per query (slices of one GPU, or the all-gathered shard results).
// One warp per query; rank = own index + sum over the other lists of lower_bound(key).
// Layout: keys[chunk][l][cq][k] with cq = list_stride / k queries per chunk; query q lives in chunk
// q / cq.  One chunk (q < cq) is the plain [L][cq][k] layout of slices / row shards; several chunks
// are the all-gathered records of a (query chunk x row shard) grid of ranks.  Missing results (fewer
// than k real records) are written as (+inf | INT32_MAX, UINT64_MAX).
// ---------------------------------------------------------------------------------------------
// ptrs != nullptr: list l of query q starts at ptrs[l] + q*k instead -- the lists may then live in the memory of OTHER
// GPUs (peer access over NVLink): the gather of the shards' candidates and their merge are this one kernel.
__global__ void topk_merge_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* const* __restrict__ ptrs, int L,
                                  long long nq, int k, long long list_stride, float* __restrict__ out_dist_f, int* __restrict__ out_dist_i,
                                  unsigned long long* __restrict__ out_id, unsigned long long* __restrict__ out_key) {
    extern __shared__ unsigned long long s_keys[];  // [warps][L*k]: the query's lists, staged once
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + w;
    if (q >= nq) return;
    unsigned long long* sk = s_keys + (size_t)w * L * k;
    if (ptrs) {
        for (int e = lane; e < L * k; e += 32) {
            const int l = e / k, j = e - l * k;
            sk[e] = ptrs[l][q * k + j];
        }
    } else {
        const long long cq = list_stride / k, chunk = q / cq;
        const unsigned long long* base = keys + (chunk * (L - 1) * cq + q) * k;  // = ((chunk*L)*cq + (q - chunk*cq)) * k
        for (int e = lane; e < L * k; e += 32) {
            const int l = e / k, j = e - l * k;
            sk[e] = base[(long long)l * list_stride + j];
        }
    }
    __syncwarp();
    int real = 0;
    for (int e = lane; e < L * k; e += 32) {
        const int l = e / k, j = e - l * k;
        const unsigned long long key = sk[e];
        if (key == KEY_MAX) continue;
        real++;
        int rank = j;
        for (int l2 = 0; l2 < L && rank < k; l2++) {  // a key already ranked past k is out whatever the other lists hold
            if (l2 == l) continue;
            const unsigned long long* o = sk + l2 * k;
            int lo = 0, hi = k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (o[mid] < key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            const long long o = q * k + rank;
            const uint32_t ord = (uint32_t)(key >> 32);
            if (out_dist_f) out_dist_f[o] = f32_from_orderable(ord);
            if (out_dist_i) out_dist_i[o] = s32_from_orderable(ord);
            if (out_id) out_id[o] = key & 0xFFFFFFFFull;
            if (out_key) out_key[o] = key;
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) real += __shfl_xor_sync(0xffffffffu, real, s);
    for (int r = real + lane; r < k; r += 32) {
        const long long o = q * k + r;
        if (out_dist_f) out_dist_f[o] = __int_as_float(0x7f800000);
        if (out_dist_i) out_dist_i[o] = 0x7fffffff;
        if (out_id) out_id[o] = 0xFFFFFFFFFFFFFFFFull;
        if (out_key) out_key[o] = KEY_MAX;
    }
}

// ---------------------------------------------------------------------------------------------
// Few queries, many lists (a small batch cut into up to one piece per SM: L ~ 19 at 64 queries x 20 k rows): one warp per
// query leaves most of the machine idle while that warp walks L*k*(L-1) binary searches.  Here a whole CTA works on one
// query: same rank rule, the keys are spread over all threads.  Same layout and outputs as topk_merge_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void topk_merge_cta_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* const* __restrict__ ptrs, int L,
                                      long long nq, int k, long long list_stride, float* __restrict__ out_dist_f, int* __restrict__ out_dist_i,
                                      unsigned long long* __restrict__ out_id, unsigned long long* __restrict__ out_key) {
    extern __shared__ unsigned long long sk[];  // [L*k]
    __shared__ int s_real;
    const long long q = blockIdx.x;
    if (threadIdx.x == 0) s_real = 0;
    if (ptrs) {
        for (int e = threadIdx.x; e < L * k; e += blockDim.x) {
            const int l = e / k, j = e - l * k;
            sk[e] = ptrs[l][q * k + j];
        }
    } else {
        const long long cq = list_stride / k, chunk = q / cq;
        const unsigned long long* base = keys + (chunk * (L - 1) * cq + q) * k;
        for (int e = threadIdx.x; e < L * k; e += blockDim.x) {
            const int l = e / k, j = e - l * k;
            sk[e] = base[(long long)l * list_stride + j];
        }
    }
    __syncthreads();
    int real = 0;
    for (int e = threadIdx.x; e < L * k; e += blockDim.x) {
        const int l = e / k, j = e - l * k;
        const unsigned long long key = sk[e];
        if (key == KEY_MAX) continue;
        real++;
        int rank = j;
        for (int l2 = 0; l2 < L && rank < k; l2++) {  // a key already ranked past k is out whatever the other lists hold
            if (l2 == l) continue;
            const unsigned long long* o = sk + l2 * k;
            int lo = 0, hi = k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (o[mid] < key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            const long long o = q * k + rank;
            const uint32_t ord = (uint32_t)(key >> 32);
            if (out_dist_f) out_dist_f[o] = f32_from_orderable(ord);
            if (out_dist_i) out_dist_i[o] = s32_from_orderable(ord);
            if (out_id) out_id[o] = key & 0xFFFFFFFFull;
            if (out_key) out_key[o] = key;
        }
    }
    if (real) atomicAdd(&s_real, real);
    __syncthreads();
    for (int r = s_real + threadIdx.x; r < k; r += blockDim.x) {
        const long long o = q * k + r;
        if (out_dist_f) out_dist_f[o] = __int_as_float(0x7f800000);
        if (out_dist_i) out_dist_i[o] = 0x7fffffff;
        if (out_id) out_id[o] = 0xFFFFFFFFFFFFFFFFull;
        if (out_key) out_key[o] = KEY_MAX;
    }
}

// ---------------------------------------------------------------------------------------------
// Many short lists (L * k <= 512 keys, e.g. the 36 x 10 lists of the u8 tensor-core scan): the rank merge above
// would spend (L - 1) binary searches per key.  Here every lane keeps <= 16 of the query's keys in registers and
// the warp extracts the minimum k times (lane-local minimum + shuffle minimum); order inside the lists is
// irrelevant.  Keys are distinct apart from the KEY_MAX padding.  Same layout and outputs as topk_merge_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void topk_merge_extract_kernel(const unsigned long long* __restrict__ keys, int L, long long nq, int k,
                                          long long list_stride, float* __restrict__ out_dist_f, int* __restrict__ out_dist_i,
                                          unsigned long long* __restrict__ out_id, unsigned long long* __restrict__ out_key) {
    constexpr int E = 16;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + w;
    if (q >= nq) return;
    const long long cq = list_stride / k, chunk = q / cq;
    const unsigned long long* base = keys + (chunk * (L - 1) * cq + q) * k;
    unsigned long long x[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        x[e] = KEY_MAX;
        if (idx < L * k) {
            const int l = idx / k, j = idx - l * k;
            x[e] = __ldg(base + (long long)l * list_stride + j);
        }
    }
    for (int r = 0; r < k; r++) {
        unsigned long long m = x[0];
#pragma unroll
        for (int e = 1; e < E; e++) m = x[e] < m ? x[e] : m;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, s);
            m = o < m ? o : m;
        }
        if (m != KEY_MAX) {
#pragma unroll
            for (int e = 0; e < E; e++) x[e] = (x[e] == m) ? KEY_MAX : x[e];
        }
        if (lane == (r & 31)) {
            const long long o = q * k + r;
            const uint32_t ord = (uint32_t)(m >> 32);
            const bool real = m != KEY_MAX;
            if (out_dist_f) out_dist_f[o] = real ? f32_from_orderable(ord) : __int_as_float(0x7f800000);
            if (out_dist_i) out_dist_i[o] = real ? s32_from_orderable(ord) : 0x7fffffff;
            if (out_id) out_id[o] = real ? (m & 0xFFFFFFFFull) : 0xFFFFFFFFFFFFFFFFull;
            if (out_key) out_key[o] = m;
        }
    }
}

int launch_topk_merge(Ctx* ctx, const unsigned long long* keys, int L, long long nq, int k, long long list_stride,
                      float* out_dist_f, int* out_dist_i, unsigned long long* out_id, unsigned long long* out_key) {
    if (nq <= 0) return 0;
    if (L >= 8 && (long long)L * k <= 512) {  // many short lists: minimum extraction from registers
        const int wq = 4;
        topk_merge_extract_kernel<<<(unsigned)((nq + wq - 1) / wq), wq * 32, 0, ctx->stream>>>(keys, L, nq, k, list_stride, out_dist_f,
                                                                                             out_dist_i, out_id, out_key);
        ctx->launches++;
        B2_CUDA(cudaGetLastError());
        return 0;
    }
    if (L >= 3 && nq <= 4LL * ctx->sm_count && (size_t)L * k * 8 <= 200 * 1024) {  // small batch: a CTA per query
        const size_t smem = (size_t)L * k * 8;
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(topk_merge_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_merge_cta_kernel<<<(unsigned)nq, 256, smem, ctx->stream>>>(keys, nullptr, L, nq, k, list_stride, out_dist_f, out_dist_i, out_id, out_key);
        ctx->launches++;
        B2_CUDA(cudaGetLastError());
        return 0;
    }
    int warps = 4;
    while (warps > 1 && (size_t)warps * L * k * 8 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * L * k * 8;
    if (smem > 200 * 1024) B2_FAIL(-4, "topk_merge: too many lists x k for one warp's shared memory");
    if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)((nq + warps - 1) / warps), warps * 32, smem, ctx->stream>>>(
        keys, nullptr, L, nq, k, list_stride, out_dist_f, out_dist_i, out_id, out_key);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_topk_merge_ptrs(Ctx* ctx, const unsigned long long* const* list_ptrs_dev, int L, long long nq, int k, float* out_dist_f,
                           int* out_dist_i, unsigned long long* out_id, unsigned long long* out_key) {
    if (nq <= 0) return 0;
    int warps = 4;
    while (warps > 1 && (size_t)warps * L * k * 8 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * L * k * 8;
    if (smem > 200 * 1024) B2_FAIL(-4, "topk_merge: too many lists x k for one warp's shared memory");
    if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)((nq + warps - 1) / warps), warps * 32, smem, ctx->stream>>>(nullptr, list_ptrs_dev, L, nq, k, 0,
                                                                                               out_dist_f, out_dist_i, out_id, out_key);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
