// flat_kernels.cu -- exact scan behind hnswlib's BruteforceSearch (brute_force_search/src/
// brutoforce.hpp:73-93) with the reference's own distance arithmetic:
//   metric 0: 1 - <q,x>     InnerProduct / InnerProductSIMD4Ext / SIMD16Ext  (space_ip.hpp:25-207)
//   metric 1: sum (q-x)^2   L2Sqr / L2SqrSIMD4Ext / SIMD16Ext                (space_l2.h:26-151)
//   metric 2: sum (q-x)^2   L2SqrI over unsigned char, exact int32           (space_l2.h:186-219)
// fp32 distances keep the reference's accumulation order: L lane accumulators (L = 1 scalar loop,
// 4 = SSE, 8 = AVX), lane l sums elements l, l+L, ... (mul then add, never fused), lanes added left
// to right -- so distances are bit-identical and the (dist, label) order matches the max-heap rule.
//
// The fp32 metrics run on the register-tiled kernel of flat_tile.cu; this file keeps the generic uint8 kernel (metric 2
// for shapes the tensor-core kernel of u8_scan_tc.cu does not take): one CTA = QT queries x a slice of rows, a lane owns
// a row, queries are broadcast from shared memory.  Selection: ballot-compacted staging + sorted CTA lists (topk.cuh);
// keys carry the RANK of the row's label so that ties resolve on the label exactly as std::pair<dist_t,labeltype> does.
#include <algorithm>

#include "flat_kernels.cuh"
#include "topk.cuh"

namespace b200nn {

constexpr int FLAT_QT = 8;      // queries per CTA
constexpr int FLAT_WARPS = 8;
constexpr int FLAT_SBW = 64;    // staging records per (warp, query)

struct FlatShared {
    unsigned long long list[FLAT_QT][KP];
    unsigned long long stage[FLAT_WARPS][FLAT_QT][FLAT_SBW];
    unsigned long long tau[FLAT_QT];
    int lock[FLAT_QT];
};

// L2SqrI: res += (a-b)*(a-b) over the first (d>>2)*4 bytes (space_l2.h:199-213), exact int32.
// |a-b| fits a byte, so 4 squared differences = one __vabsdiffu4 + one dp4a.
__global__ void __launch_bounds__(FLAT_WARPS * 32)
flat_scan_u8_kernel(const unsigned char* __restrict__ data, const uint32_t* __restrict__ rank, long long n, int d,
                    const unsigned char* __restrict__ queries, long long nq, int n_slices, int k,
                    unsigned long long* __restrict__ out_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FlatShared& S = *reinterpret_cast<FlatShared*>(smem_raw);
    uint32_t* s_q = reinterpret_cast<uint32_t*>(smem_raw + sizeof(FlatShared));  // [QT][d/4] words
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long q0 = (long long)blockIdx.x * FLAT_QT;
    const int slice = blockIdx.y;
    const int dw = d >> 2;  // 32-bit words actually used
    for (int i = threadIdx.x; i < FLAT_QT * KP; i += blockDim.x) (&S.list[0][0])[i] = KEY_MAX;
    if (threadIdx.x < FLAT_QT) { S.tau[threadIdx.x] = KEY_MAX; S.lock[threadIdx.x] = 0; }
    for (int i = threadIdx.x; i < FLAT_QT * dw; i += blockDim.x) {
        const int qq = i / dw, wd = i - qq * dw;
        uint32_t v = 0;
        if (q0 + qq < nq) {
            const unsigned char* qp = queries + (q0 + qq) * d + wd * 4;
            v = (uint32_t)qp[0] | ((uint32_t)qp[1] << 8) | ((uint32_t)qp[2] << 16) | ((uint32_t)qp[3] << 24);
        }
        s_q[i] = v;
    }
    __syncthreads();
    const long long r_lo = (n * slice) / n_slices, r_hi = (n * (slice + 1)) / n_slices;
    int cnt[FLAT_QT];
#pragma unroll
    for (int t = 0; t < FLAT_QT; t++) cnt[t] = 0;
    volatile unsigned long long* tau = S.tau;
    const bool aligned4 = (d & 3) == 0;
    for (long long base = r_lo + (long long)w * 32; base < r_hi; base += FLAT_WARPS * 32) {
        const long long row = base + lane;
        const bool valid = row < r_hi;
        int acc[FLAT_QT];
#pragma unroll
        for (int t = 0; t < FLAT_QT; t++) acc[t] = 0;
        if (valid) {
            const unsigned char* x = data + row * d;
            for (int wd = 0; wd < dw; wd++) {
                uint32_t xv;
                if (aligned4) xv = __ldg(reinterpret_cast<const uint32_t*>(x) + wd);
                else xv = (uint32_t)x[4 * wd] | ((uint32_t)x[4 * wd + 1] << 8) | ((uint32_t)x[4 * wd + 2] << 16) | ((uint32_t)x[4 * wd + 3] << 24);
#pragma unroll
                for (int t = 0; t < FLAT_QT; t++) {
                    const uint32_t ad = __vabsdiffu4(s_q[t * dw + wd], xv);
                    acc[t] = (int)__dp4a(ad, ad, (unsigned)acc[t]);
                }
            }
        }
        const uint32_t rk = valid ? rank[row] : 0u;
#pragma unroll
        for (int t = 0; t < FLAT_QT; t++) {
            const unsigned long long key = make_key(s32_orderable(acc[t]), rk);
            const bool pass = valid && (q0 + t < nq) && key < tau[t];
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                const uint32_t st = smem_u32(&S.stage[w][t][0]);
                if (pass) sts64(st + (uint32_t)(cnt[t] + __popc(m & ((1u << lane) - 1))) * 8u, key);
                cnt[t] += __popc(m);
                __syncwarp();
                if (cnt[t] > FLAT_SBW - 32) {
                    for (int off = 0; off < cnt[t]; off += 32)
                        warp_flush(smem_u32(&S.list[t][0]), &S.lock[t], tau + t, st + off * 8, min(32, cnt[t] - off), k);
                    cnt[t] = 0;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < FLAT_QT; t++) {
        const uint32_t st = smem_u32(&S.stage[w][t][0]);
        for (int off = 0; off < cnt[t]; off += 32)
            warp_flush(smem_u32(&S.list[t][0]), &S.lock[t], tau + t, st + off * 8, min(32, cnt[t] - off), k);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FLAT_QT * k; i += blockDim.x) {
        const int t = i / k, j = i - t * k;
        if (q0 + t < nq) out_keys[((long long)slice * nq + q0 + t) * k + j] = S.list[t][j];
    }
}

// rank -> label after the merge (ids in keys are label ranks)
__global__ void rank_to_label_kernel(unsigned long long* __restrict__ ids, long long count,
                                     const unsigned long long* __restrict__ label_sorted) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long r = ids[i];
        ids[i] = (r == 0xFFFFFFFFFFFFFFFFull) ? r : label_sorted[r];
    }
}

int flat_pick_slices(int sm_count, long long nq, long long n) {
    const long long qt = (nq + FLAT_QT - 1) / FLAT_QT;
    long long s = (2LL * sm_count + qt - 1) / qt;
    const long long max_s = std::max<long long>(1, n / 2048);
    if (s > max_s) s = max_s;
    if (s > 1024) s = 1024;
    if (s < 1) s = 1;
    return (int)s;
}

int launch_flat_scan(Ctx* ctx, int metric, int order, const void* data, const uint32_t* rank, long long n, int d,
                     const void* queries, long long nq, int n_slices, int k, unsigned long long* out_keys) {
    if (nq <= 0) return 0;
    if (k < 1 || k > KP) B2_FAIL(-4, "flat search supports 1 <= k <= 128");
    if ((size_t)d * FLAT_QT * 4 + sizeof(FlatShared) > 200 * 1024) B2_FAIL(-4, "flat search: dimension too large");
    if (metric != 2) B2_FAIL(-1, "launch_flat_scan: the fp32 metrics run on the tiled kernel (launch_flat_scan_f32, flat_tile.cu)");
    const size_t smem = sizeof(FlatShared) + (size_t)FLAT_QT * (d >> 2) * 4 + 16;
    B2_CUDA(cudaFuncSetAttribute(flat_scan_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((nq + FLAT_QT - 1) / FLAT_QT), (unsigned)n_slices);
    flat_scan_u8_kernel<<<grid, FLAT_WARPS * 32, smem, ctx->stream>>>((const unsigned char*)data, rank, n, d,
                                                                    (const unsigned char*)queries, nq, n_slices, k, out_keys);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_rank_to_label(Ctx* ctx, unsigned long long* ids, long long count, const unsigned long long* label_sorted) {
    if (count <= 0) return 0;
    rank_to_label_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(ids, count, label_sorted);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
