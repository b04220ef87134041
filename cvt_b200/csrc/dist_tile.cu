// dist_tile.cu -- all-pairs squared distances between rows and centroids, register-tiled, in the reference's arithmetic.
//
// Users: a2 coarse assignment (IVFOPQ::Add, opq/src/IVFOPQ.cpp:107-129), a4 coarse probes (IVFOPQ::Query, :238-260) and
// the assignment pass of k-means training (f-4).  At the shipped K = 8192 this loop is what dominates the reference's
// indexing (3 K D flop per row).  The per-pair arithmetic must stay the reference's --
//     acc = 0.0f; for t ascending: tmp = a[t] - b[t]; acc = acc + tmp * tmp;      (two roundings per term, no FMA)
// -- because list ids and therefore everything downstream are compared bit for bit.  That rules the tensor cores out
// (their accumulation order and rounding are not the reference's); what is left is to make the FP32 pipe the only
// limit: 3 issue slots per (pair, t).  A one-warp-per-row kernel spends those slots waiting on a global load and a
// shared load per 3 flops (measured 0.22 T pair-elements/s); here a CTA owns a 128 x 128 tile of pairs, every thread an
// 8 x 8 sub-tile in registers, and each t costs a thread 4 shared-memory loads (16 bytes each) for 192 FP32 operations.
// Distances go to a row-chunk matrix that stays in L2/HBM for the selection kernels below (argmin with the reference's
// first-minimum rule, or the nk smallest in the reference's pop order): 4 bytes written + read per K D / 1 work.
#include <algorithm>

#include "dist_tile.cuh"

namespace b200nn {

namespace {

constexpr int TR = 128, TC = 128, DC = 16;  // tile rows, tile centroids, t per chunk
constexpr int XS = TR + 4;                   // row stride of the transposed x chunk (keeps 16-byte alignment, spreads banks)

__global__ void __launch_bounds__(256, 2)
sqdist_tile_kernel(const float* __restrict__ x, long long ld, int col0, long long n_rows, int d, const float* __restrict__ cT, int K,
                   float* __restrict__ dmat, long long ldm) {
    __shared__ __align__(16) float xs[2][DC * XS];  // [t][row]
    __shared__ __align__(16) float cs[2][DC * TC];  // [t][centroid]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.y * TR;
    const int c0 = blockIdx.x * TC;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
    float px[8], pc[8];
    // element e = tid + 256 i of a chunk.  x: row = e / 16, t = e % 16 (64-byte runs along t);  c: t = e / 128, col = e % 128
    auto fetch = [&](int t0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + 256 * i;
            const int r = e >> 4, t = t0 + (e & 15);
            px[i] = (row0 + r < n_rows && t < d) ? __ldg(x + (row0 + r) * ld + col0 + t) : 0.0f;
            const int tc = t0 + (e >> 7), cc = c0 + (e & 127);
            pc[i] = (tc < d && cc < K) ? __ldg(cT + (long long)tc * K + cc) : 0.0f;
        }
    };
    auto stage = [&](int b) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + 256 * i;
            xs[b][(e & 15) * XS + (e >> 4)] = px[i];
            cs[b][e] = pc[i];
        }
    };
    const int nch = (d + DC - 1) / DC;
    fetch(0);
    stage(0);
    __syncthreads();
    for (int ch = 0; ch < nch; ch++) {
        const int b = ch & 1;
        if (ch + 1 < nch) fetch((ch + 1) * DC);
        // padded t (>= d) contribute (0 - 0)^2 = +0.0f, which leaves a non-negative accumulator unchanged bit for bit
#pragma unroll
        for (int t = 0; t < DC; t++) {
            const float4 xa = *reinterpret_cast<const float4*>(&xs[b][t * XS + ty * 4]);
            const float4 xb = *reinterpret_cast<const float4*>(&xs[b][t * XS + 64 + ty * 4]);
            const float4 ca = *reinterpret_cast<const float4*>(&cs[b][t * TC + tx * 4]);
            const float4 cb = *reinterpret_cast<const float4*>(&cs[b][t * TC + 64 + tx * 4]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float df = __fsub_rn(xv[i], cv[j]);
                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(df, df));
                }
        }
        if (ch + 1 < nch) stage(b ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= n_rows) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = c0 + h * 64 + tx * 4;
            if (c < ldm)  // ldm is a multiple of 4: a group of 4 columns is entirely inside or outside
                *reinterpret_cast<float4*>(dmat + r * ldm + c) = make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
        }
    }
}

__device__ __forceinline__ void dt_warp_argmin(float& best, int& idx) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        const bool take = (oi >= 0) && (idx < 0 || ob < best || (ob == best && oi < idx));
        if (take) { best = ob; idx = oi; }
    }
}

// argmin per row with the reference's rule (IVFOPQ.cpp:111-127): running minimum starts at (float)UINT_MAX, strict '<',
// so the lowest index wins ties and the index stays -1 when nothing is smaller.  One warp per row.
__global__ void argmin_rows_kernel(const float* __restrict__ dmat, long long ldm, long long n_rows, int K, int* __restrict__ out_idx,
                                   float* __restrict__ out_dist) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n_rows; r += nwarps) {
        float best = 4294967296.0f;
        int idx = -1;
        for (int c = lane; c < K; c += 32) {
            const float v = dmat[r * ldm + c];
            if (v < best) { best = v; idx = c; }
        }
        dt_warp_argmin(best, idx);
        if (lane == 0) {
            out_idx[r] = idx;
            if (out_dist) out_dist[r] = best;
        }
    }
}

// the nk smallest centroids of a row under (dist, index), written in the reference's pop order (largest first,
// IVFOPQ.cpp:238-260).  Same selection as coarse_probe_kernel (pq_kernels.cu), on precomputed distances.
__global__ void probe_rows_kernel(const float* __restrict__ dmat, long long ldm, long long n_rows, int K, int nk, int* __restrict__ out_lists) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int MAXNK = 8;
    for (long long row = warp; row < n_rows; row += nwarps) {
        float bd[MAXNK];
        int bi[MAXNK];
#pragma unroll
        for (int t = 0; t < MAXNK; t++) { bd[t] = __int_as_float(0x7f800000); bi[t] = 0x7fffffff; }
        for (int c = lane; c < K; c += 32) {
            float dd = dmat[row * ldm + c];
            int id = c;
#pragma unroll
            for (int t = 0; t < MAXNK; t++) {
                if (t < nk && (dd < bd[t] || (dd == bd[t] && id < bi[t]))) {
                    const float td = bd[t]; const int ti = bi[t];
                    bd[t] = dd; bi[t] = id; dd = td; id = ti;
                }
            }
        }
        for (int r = 0; r < nk; r++) {
            float dd = bd[0]; int id = bi[0];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, dd, s);
                const int oi = __shfl_xor_sync(0xffffffffu, id, s);
                if (od < dd || (od == dd && oi < id)) { dd = od; id = oi; }
            }
            if (bi[0] == id && bd[0] == dd) {
#pragma unroll
                for (int t = 0; t + 1 < MAXNK; t++) { bd[t] = bd[t + 1]; bi[t] = bi[t + 1]; }
                bd[MAXNK - 1] = __int_as_float(0x7f800000); bi[MAXNK - 1] = 0x7fffffff;
            }
            if (lane == 0) out_lists[row * nk + (nk - 1 - r)] = id;
        }
    }
}

constexpr size_t DMAT_MAX_BYTES = 128u << 20;

}  // namespace

// the row-chunk distance matrix lives in the context (grown on demand; shared by the nearest-centroid path and the flat scan)
int ensure_dmat(Ctx* ctx, size_t elems) {
    if (elems <= ctx->dmat_elems) return 0;
    if (ctx->dmat) {
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dmat);
        ctx->dmat = nullptr;
        ctx->dmat_elems = 0;
    }
    B2_CUDA(cudaMalloc(&ctx->dmat, elems * sizeof(float)));
    ctx->dmat_elems = elems;
    return 0;
}

bool tiled_nearest_pays(int d, int K) { return d >= 32 && K >= 64; }

int launch_tiled_nearest(Ctx* ctx, const float* x, long long ld, int col0, long long n, int d, const float* cT, int K, int rule, int nk,
                         int* out_idx, float* out_dist) {
    if (n <= 0) return 0;
    if (rule == 0) nk = 1;
    if (nk < 1 || nk > 8 || nk > K) B2_FAIL(-1, "tiled_nearest: nk must be in [1, min(8, K)]");
    const long long ldm = ((long long)K + 3) / 4 * 4;
    long long rc_rows = (long long)(DMAT_MAX_BYTES / sizeof(float)) / ldm / TR * TR;
    rc_rows = std::max<long long>(TR, std::min<long long>(rc_rows, (n + TR - 1) / TR * TR));
    int rc;
    if ((rc = ensure_dmat(ctx, (size_t)(rc_rows * ldm)))) return rc;
    const unsigned ctiles = (unsigned)((K + TC - 1) / TC);
    for (long long r0 = 0; r0 < n; r0 += rc_rows) {
        const long long rows = std::min(rc_rows, n - r0);
        dim3 grid(ctiles, (unsigned)((rows + TR - 1) / TR));
        sqdist_tile_kernel<<<grid, 256, 0, ctx->stream>>>(x + r0 * ld, ld, col0, rows, d, cT, K, ctx->dmat, ldm);
        const unsigned sgrid = (unsigned)std::max<long long>(1, std::min<long long>((rows + 7) / 8, (long long)ctx->sm_count * 8));
        if (rule == 0)
            argmin_rows_kernel<<<sgrid, 256, 0, ctx->stream>>>(ctx->dmat, ldm, rows, K, out_idx + r0, out_dist ? out_dist + r0 : nullptr);
        else
            probe_rows_kernel<<<sgrid, 256, 0, ctx->stream>>>(ctx->dmat, ldm, rows, K, nk, out_idx + r0 * nk);
        ctx->launches += 2;
    }
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn

// =============================================================================================
// a3 on the register tile: residual + PQ argmin encode (IVFOPQ::Add, IVFOPQ.cpp:135-163).
// CTA = 128 rows x ONE sub-quantizer m: the rows' residual sub-vectors (x - coarse[list], the reference's own fp32 subtraction)
// and the 256 codewords of m sit in shared memory, transposed; a thread owns an 8 x 8 tile of (row, codeword) pairs per half of
// the codebook (64 accumulators), every element costs 4 shared loads (16 bytes each) for 192 FP32 instructions in the
// reference's arithmetic (t = r - c; acc = acc + t * t), and the first-minimum rule is kept by reducing (dist, index) pairs
// lexicographically -- inside the thread in index order, then across the 16 threads that share the rows.
// The one-warp-per-(row, m) kernel this replaces re-read the codebook through the read-only path for every row (24 % of the
// FP32 issue rate).
// =============================================================================================
namespace b200nn {
namespace {

template <int DS>
__global__ void __launch_bounds__(256, 2)
pq_encode_tile_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ coarse, const int* __restrict__ list,
                      const float* __restrict__ cbT /*[M][DS][256]*/, int M, unsigned char* __restrict__ codes) {
    constexpr int TR = 128, XS = TR + 4, KS = 256;
    __shared__ __align__(16) float xs[DS * XS];   // [t][row]   residuals
    __shared__ __align__(16) float cs[DS * KS];   // [t][codeword]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m = blockIdx.y;
    for (int e = tid; e < DS * KS; e += 256) cs[e] = __ldg(cbT + (long long)m * DS * KS + e);  // the codebook of m: staged once per CTA
    for (long long row0 = (long long)blockIdx.x * TR; row0 < n; row0 += (long long)gridDim.x * TR) {
    __syncthreads();  // the previous tile's residuals are consumed
    for (int e = tid; e < TR * DS; e += 256) {  // residual sub-vector of row r, element t (runs of DS floats per row)
        const int r = e / DS, t = e - r * DS;
        float v = 0.0f;
        if (row0 + r < n) {
            const int vw = list[row0 + r];
            v = __fsub_rn(__ldg(x + (row0 + r) * D + m * DS + t), __ldg(coarse + (long long)(vw < 0 ? 0 : vw) * D + m * DS + t));
        }
        xs[t * XS + r] = v;
    }
    __syncthreads();
    float best[8];
    int bidx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { best[i] = 4294967296.0f; bidx[i] = -1; }  // (float)UINT_MAX, IVFOPQ.cpp:143
#pragma unroll 1
    for (int h = 0; h < 2; h++) {  // codewords [128 h, 128 h + 128)
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
#pragma unroll 2
        for (int t = 0; t < DS; t++) {
            const float4 xa = *reinterpret_cast<const float4*>(&xs[t * XS + ty * 4]);
            const float4 xb = *reinterpret_cast<const float4*>(&xs[t * XS + 64 + ty * 4]);
            const float4 ca = *reinterpret_cast<const float4*>(&cs[t * KS + h * 128 + tx * 4]);
            const float4 cb = *reinterpret_cast<const float4*>(&cs[t * KS + h * 128 + 64 + tx * 4]);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float df = __fsub_rn(xv[i], cv[j]);
                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(df, df));
                }
        }
        // this thread's codewords in ascending index order: strict '<' keeps the first minimum
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = h * 128 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (acc[i][j] < best[i]) { best[i] = acc[i][j]; bidx[i] = c; }
        }
    }
    // across the 16 threads (tx) that hold the same rows: lexicographic (dist, index) = the sequential first-minimum rule
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int s = 8; s >= 1; s >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[i], s);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], s);
            const bool take = (oi >= 0) && (bidx[i] < 0 || ob < best[i] || (ob == best[i] && oi < bidx[i]));
            if (take) { best[i] = ob; bidx[i] = oi; }
        }
        const long long r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (tx == 0 && r < n) codes[r * M + m] = (unsigned char)bidx[i];  // elem.PQindex[i] = vw1, IVFOPQ.cpp:161
    }
    }  // row tiles of this CTA
}

}  // namespace

bool pq_encode_tile_supported(int ds, int ksub) { return ksub == 256 && (ds == 4 || ds == 8 || ds == 16); }

int launch_pq_encode_tile(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* list, const float* cbT, int M, int ksub,
                          unsigned char* codes) {
    if (n <= 0) return 0;
    const int ds = D / M;
    if (!pq_encode_tile_supported(ds, ksub)) B2_FAIL(-4, "pq_encode_tile: needs ksub = 256 and D/M in {4, 8, 16}");
    // every CTA keeps its sub-quantizer's codebook and walks a strip of row tiles: about 4 CTAs per SM in total
    const long long tiles = (n + 127) / 128;
    const dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(tiles, (4LL * ctx->sm_count + M - 1) / M)), (unsigned)M);
    switch (ds) {
        case 4: pq_encode_tile_kernel<4><<<grid, 256, 0, ctx->stream>>>(x, n, D, coarse, list, cbT, M, codes); break;
        case 8: pq_encode_tile_kernel<8><<<grid, 256, 0, ctx->stream>>>(x, n, D, coarse, list, cbT, M, codes); break;
        case 16: pq_encode_tile_kernel<16><<<grid, 256, 0, ctx->stream>>>(x, n, D, coarse, list, cbT, M, codes); break;
    }
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
