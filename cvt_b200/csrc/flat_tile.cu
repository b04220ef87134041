// flat_tile.cu -- the exact fp32 scan behind hnswlib's BruteforceSearch<float> (brute_force_search/src/brutoforce.hpp:73-93;
// the kernel that replaces HNSW in hnsw_sifts_retrieval/makeSearch.cpp:52) as a register-tiled all-pairs kernel in the
// reference's own distance arithmetic:
//   metric 0: 1 - <q,x>     InnerProduct / InnerProductSIMD4Ext / SIMD16Ext  (space_ip.hpp:25-207)
//   metric 1: sum (q-x)^2   L2Sqr / L2SqrSIMD4Ext / SIMD16Ext                (hnswlib/space_l2.h:26-151)
// The distances keep the reference's accumulation order: L lane accumulators (L = 1 scalar loop, 4 = SSE, 8 = AVX), lane l
// sums elements l, l+L, ... in order (mul then add, never fused), the lanes are added left to right -- bit-identical
// distances, so the (dist, label) selection is the reference's max-heap rule exactly.  That order is a per-pair property:
// a pair's L partial sums can live in the registers of ONE thread, which is what lets the scan be tiled like a GEMM.
//
// A CTA owns a (16 TI rows) x (16 TJ queries) tile of pairs, a thread a TI x TJ sub-tile with TI*TJ*L accumulators (8x8x1 = 64,
// 8x4x4 = 128, 4x4x8 = 128); rows and queries stream through shared memory in chunks of 16
// elements, transposed so that a thread fetches its TI row values / TJ query values with 16-byte loads; every element costs
// the thread (TI+TJ)/4 shared loads for 2 (IP) or 3 (L2) x TI*TJ FP32 instructions.  Distances go to a [queries][rows]
// chunk matrix that stays in L2/HBM (<= 128 MB per row chunk) for the selection kernel (dense_topk_kernel: threshold
// filter + the shared sorted lists of topk.cuh), whose keys carry the RANK of the row's label.
// The lane-per-row kernel this replaces streamed every row with 32 different 512-byte-strided loads per request and
// measured 0.6 T pair-elements/s.
#include <stdlib.h>

#include <algorithm>

#include "dist_tile.cuh"
#include "flat_kernels.cuh"
#include "pq_kernels.cuh"
#include "topk.cuh"

namespace b200nn {

namespace {

constexpr int FT_DC = 16;  // elements per chunk (a multiple of every L)

// L = 1 fits two CTAs per SM in 128 registers (64 accumulators); the lane-split orders hold 128 accumulators per thread and
// run one CTA per SM with the full register file (8 warps of 128-way ILP keep the FP32 pipe fed)
template <int METRIC, int L, int TI, int TJ>
__global__ void __launch_bounds__(256, (L == 1) ? 2 : 1)
flat_tile_f32_kernel(const float* __restrict__ x, long long n_rows, int d, const float* __restrict__ q, long long nq,
                     float* __restrict__ dmat /*[nq_pad][ldm]*/, long long ldm) {
    constexpr int TR = 16 * TI, TQ = 16 * TJ, XS = TR + 4, QS = TQ + 4;
    constexpr int NX = TR * FT_DC / 256, NQ = TQ * FT_DC / 256;  // elements a thread fetches per chunk
    __shared__ __align__(16) float xs[2][FT_DC * XS];  // [t][row]
    __shared__ __align__(16) float qs[2][FT_DC * QS];  // [t][query]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * TR, q0 = (long long)blockIdx.y * TQ;
    float acc[TI][TJ][L];
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
        for (int j = 0; j < TJ; j++)
#pragma unroll
            for (int l = 0; l < L; l++) acc[i][j][l] = 0.0f;
    float px[NX], pq[NQ];
    // element e = tid + 256 i of a chunk: entity = e / 16, t = e % 16 (64-byte runs along t)
    auto fetch = [&](int t0) {
#pragma unroll
        for (int i = 0; i < NX; i++) {
            const int e = tid + 256 * i, r = e >> 4, t = t0 + (e & 15);
            px[i] = (row0 + r < n_rows && t < d) ? __ldg(x + (row0 + r) * d + t) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            const int e = tid + 256 * i, r = e >> 4, t = t0 + (e & 15);
            pq[i] = (q0 + r < nq && t < d) ? __ldg(q + (q0 + r) * d + t) : 0.0f;
        }
    };
    auto stage = [&](int b) {
#pragma unroll
        for (int i = 0; i < NX; i++) {
            const int e = tid + 256 * i;
            xs[b][(e & 15) * XS + (e >> 4)] = px[i];
        }
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            const int e = tid + 256 * i;
            qs[b][(e & 15) * QS + (e >> 4)] = pq[i];
        }
    };
    const int nch = (d + FT_DC - 1) / FT_DC;
    fetch(0);
    stage(0);
    __syncthreads();
    for (int ch = 0; ch < nch; ch++) {
        const int b = ch & 1;
        if (ch + 1 < nch) fetch((ch + 1) * FT_DC);
        // elements past d are zeros on both sides: they add +0.0f (IP: 0*0; L2: (0-0)^2) to a lane, which changes nothing
#pragma unroll
        for (int t = 0; t < FT_DC; t++) {
            float xv[TI], qv[TJ];
#pragma unroll
            for (int i = 0; i < TI; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&xs[b][t * XS + (i / 4) * 64 + ty * 4]);
                xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TJ; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&qs[b][t * QS + (j / 4) * 64 + tx * 4]);
                qv[j] = v.x; qv[j + 1] = v.y; qv[j + 2] = v.z; qv[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TI; i++)
#pragma unroll
                for (int j = 0; j < TJ; j++) {
                    float term;
                    if (METRIC == 0) term = __fmul_rn(qv[j], xv[i]);  // v1 * v2 with v1 = the query (searchKnn(query, row))
                    else { const float df = __fsub_rn(qv[j], xv[i]); term = __fmul_rn(df, df); }
                    acc[i][j][t % L] = __fadd_rn(acc[i][j][t % L], term);
                }
        }
        if (ch + 1 < nch) stage(b ^ 1);
        __syncthreads();
    }
    // lanes left to right, then 1 - sum for the inner-product "distance" (space_ip.hpp:33,130,206)
#pragma unroll
    for (int j = 0; j < TJ; j++) {
        const long long qq = q0 + (j / 4) * 64 + tx * 4 + (j & 3);
        if (qq >= nq) continue;
#pragma unroll
        for (int i = 0; i < TI; i += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float sum = acc[i + u][j][0];
#pragma unroll
                for (int l = 1; l < L; l++) sum = __fadd_rn(sum, acc[i + u][j][l]);
                o[u] = (METRIC == 0) ? __fsub_rn(1.0f, sum) : sum;
            }
            const long long r = (long long)blockIdx.x * TR + (i / 4) * 64 + ty * 4;  // chunk-relative row of o[0]
            if (r < ldm) *reinterpret_cast<float4*>(dmat + qq * ldm + r) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// A handful of queries (searchKnn is called with ONE: hnsw_sifts_retrieval/makeSearch.cpp:52): the pair tile above would spend
// 63/64 of its work on padding.  Here the rows stream through shared memory once (coalesced 16-byte loads; the scan is then
// HBM-bound: n * d * 4 bytes) and L threads share a row, thread l walking elements l, l+L, ... in order for every query --
// the reference's lane accumulators again -- after which the L partial sums are added left to right by the row's first thread.
// Row stride d + 4 words puts the 32 (row, lane) pairs of a warp on 32 different banks.
template <int METRIC, int L, int NQ>
__global__ void __launch_bounds__(256)
flat_rows_f32_kernel(const float* __restrict__ x, long long n_rows, int d, const float* __restrict__ q, int nq,
                     float* __restrict__ dmat /*[nq][ldm]*/, long long ldm) {
    extern __shared__ __align__(16) float sm[];
    constexpr int RPB = 256 / L;  // rows per pass of the CTA
    const int RS = d + 4;
    float* s_q = sm;              // [NQ][d]
    float* s_x = sm + NQ * d;     // [RPB][RS]
    const int tid = threadIdx.x, rl = tid / L, l = tid % L;
    for (int i = tid; i < NQ * d; i += 256) s_q[i] = (i / d < nq) ? q[i] : 0.0f;
    const int d4 = d >> 2;
    for (long long r0 = (long long)blockIdx.x * RPB; r0 < n_rows; r0 += (long long)gridDim.x * RPB) {
        __syncthreads();  // previous pass consumed (and s_q staged)
        for (int e = tid; e < RPB * d4; e += 256) {
            const int r = e / d4, c = e - r * d4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(x + (r0 + r) * d) + c);
            *reinterpret_cast<float4*>(s_x + r * RS + 4 * c) = v;
        }
        __syncthreads();
        float acc[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++) acc[j] = 0.0f;
        const float* xr = s_x + rl * RS;
        for (int t = l; t < d; t += L) {
            const float xv = xr[t];
#pragma unroll
            for (int j = 0; j < NQ; j++) {
                const float qv = s_q[j * d + t];
                float term;
                if (METRIC == 0) term = __fmul_rn(qv, xv);
                else { const float df = __fsub_rn(qv, xv); term = __fmul_rn(df, df); }
                acc[j] = __fadd_rn(acc[j], term);
            }
        }
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            float sum = acc[j];
#pragma unroll
            for (int o = 1; o < L; o++) {  // lanes left to right: ((a0 + a1) + a2) + ...
                const float v = __shfl_sync(0xffffffffu, acc[j], (threadIdx.x & 31) - l + o);
                sum = __fadd_rn(sum, v);
            }
            if (l == 0 && j < nq && r0 + rl < n_rows) dmat[(long long)j * ldm + r0 + rl] = (METRIC == 0) ? __fsub_rn(1.0f, sum) : sum;
        }
    }
}

template <int METRIC, int L>
int fr_launch(Ctx* ctx, const float* x, long long rows, int d, const float* q, int nq, float* dmat, long long ldm) {
    constexpr int NQ = 4;
    const size_t smem = ((size_t)NQ * d + (size_t)(256 / L) * (d + 4)) * sizeof(float);
    if (smem > 200 * 1024) return -100;  // caller falls back to the pair tile
    if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(flat_rows_f32_kernel<METRIC, L, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long passes = (rows + 256 / L - 1) / (256 / L);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(passes, (long long)ctx->sm_count * 4));
    for (int q0 = 0; q0 < nq; q0 += NQ) {
        flat_rows_f32_kernel<METRIC, L, NQ><<<grid, 256, smem, ctx->stream>>>(x, rows, d, q + (size_t)q0 * d, std::min(NQ, nq - q0), dmat + (size_t)q0 * ldm, ldm);
        ctx->launches++;
    }
    B2_CUDA(cudaGetLastError());
    return 0;
}

int fr_dispatch(Ctx* ctx, int metric, int order, const float* x, long long rows, int d, const float* q, int nq, float* dmat, long long ldm) {
    if (d % 4 != 0 || d % order != 0) return -100;
    if (metric == 0) {
        if (order == 1) return fr_launch<0, 1>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 4) return fr_launch<0, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 8) return fr_launch<0, 8>(ctx, x, rows, d, q, nq, dmat, ldm);
    } else if (metric == 1) {
        if (order == 1) return fr_launch<1, 1>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 4) return fr_launch<1, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 8) return fr_launch<1, 8>(ctx, x, rows, d, q, nq, dmat, ldm);
    }
    return -100;
}

template <int METRIC, int L, int TI, int TJ>
int ft_launch(Ctx* ctx, const float* x, long long rows, int d, const float* q, long long nq, float* dmat, long long ldm) {
    dim3 grid((unsigned)((rows + 16 * TI - 1) / (16 * TI)), (unsigned)((nq + 16 * TJ - 1) / (16 * TJ)));
    flat_tile_f32_kernel<METRIC, L, TI, TJ><<<grid, 256, 0, ctx->stream>>>(x, rows, d, q, nq, dmat, ldm);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int ft_dispatch(Ctx* ctx, int metric, int order, const float* x, long long rows, int d, const float* q, long long nq, float* dmat, long long ldm) {
    if (metric == 0) {
        if (order == 1) return ft_launch<0, 1, 8, 8>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 4) return ft_launch<0, 4, 8, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 8) return ft_launch<0, 8, 4, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
    } else if (metric == 1) {
        if (order == 1) return ft_launch<1, 1, 8, 8>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 4) return ft_launch<1, 4, 8, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
        if (order == 8) return ft_launch<1, 8, 4, 4>(ctx, x, rows, d, q, nq, dmat, ldm);
    }
    B2_FAIL(-1, "flat search: bad metric/order");
}

// bytes of the [queries][rows] distance chunk between the tile kernel and the selection.  Default: see flat_f32_plan.
static size_t ft_dmat_bytes() {
    static const size_t v = []() {
        const char* e = getenv("B200NN_FLAT_CHUNK_MB");
        const long long mb = e ? atoll(e) : 128;
        return (size_t)std::max<long long>(1, std::min<long long>(mb, 1024)) << 20;
    }();
    return v;
}

}  // namespace

// how the rows are cut: row chunks (<= 128 MB of distances each) x column slices of the selection -> lists per query
void flat_f32_plan(int sm_count, long long nq, long long n, long long* chunk_rows, int* n_chunks, int* slices) {
    long long cr = (long long)(ft_dmat_bytes() / sizeof(float)) / std::max<long long>(1, nq) / 128 * 128;
    cr = std::max<long long>(128, std::min<long long>(cr, (n + 127) / 128 * 128));
    *chunk_rows = cr;
    *n_chunks = (int)std::max<long long>(1, (n + cr - 1) / cr);
    long long s = (2LL * sm_count + nq - 1) / std::max<long long>(1, nq);
    s = std::max<long long>(1, std::min<long long>(s, std::max<long long>(1, std::min(cr, n) / 4096)));
    *slices = (int)std::min<long long>(s, 64);
}

// keys out: [n_chunks * slices][nq][k]; ids inside the keys are label RANKS
int launch_flat_scan_f32(Ctx* ctx, int metric, int order, const float* data, const uint32_t* rank, long long n, int d, const float* queries,
                         long long nq, int k, unsigned long long* out_keys) {
    if (nq <= 0) return 0;
    if (k < 1 || k > KP) B2_FAIL(-4, "flat search supports 1 <= k <= 128");
    long long cr;
    int nc, S;
    flat_f32_plan(ctx->sm_count, nq, n, &cr, &nc, &S);
    int rc;
    if ((rc = ensure_dmat(ctx, (size_t)(nq * cr)))) return rc;
    for (int c = 0; c < nc; c++) {
        const long long r0 = (long long)c * cr, rows = std::min(cr, n - r0);
        if (rows <= 0) {  // empty index: the lists stay empty
            B2_CUDA(cudaMemsetAsync(out_keys + (size_t)c * S * nq * k, 0xFF, (size_t)S * nq * k * sizeof(unsigned long long), ctx->stream));
            continue;
        }
        rc = -100;
        if (nq <= 8) rc = fr_dispatch(ctx, metric, order, data + r0 * d, rows, d, queries, (int)nq, ctx->dmat, cr);  // streaming kernel for a few queries
        if (rc == -100) rc = ft_dispatch(ctx, metric, order, data + r0 * d, rows, d, queries, nq, ctx->dmat, cr);
        if (rc) return rc;
        if ((rc = launch_dense_topk_ex(ctx, ctx->dmat, nq, rows, cr, k, S, rank + r0, out_keys + (size_t)c * S * nq * k))) return rc;
    }
    return 0;
}

}  // namespace b200nn
