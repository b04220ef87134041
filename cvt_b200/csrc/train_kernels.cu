// train_kernels.cu -- f-4: k-means training of the coarse quantizer and the PQ codebooks on the device.
//
// The reference trains with yael's kmeans (opq/train_codebook/train_PQ_codebook.cpp:164 coarse, :229 per sub-space;
// yael is un-vendored, its initialisation is random => training is numerically UNPINNED in the reference).  What is
// built here is a fully deterministic Lloyd iteration whose every arithmetic step is fixed (capi_train.cu states the
// algorithm), so that a scalar CPU restatement reproduces the centroids bit for bit:
//   assign   : the reference's own distance arithmetic (IVFOPQ.cpp:117-122: t = a - b; acc += t * t, sequential fp32,
//              strict '<' => first minimum wins) -- the same arithmetic IVFOPQ::Add will later encode with;
//   update   : per cluster, rows in ascending row order, summed in double in blocks of KM_SUM_BLOCK rows, the block sums
//              added in block order; centroid = (float)(sum / count).
// The O(n) integer bookkeeping between the two (stable counting sort of the assignment, change detection, donors for
// empty clusters) runs on the host; the O(n k d) and O(n d) floating-point work runs here.
#include <algorithm>

#include "dist_tile.cuh"
#include "train_kernels.cuh"

namespace b200nn {

namespace {

__device__ __forceinline__ void km_warp_argmin(float& best, int& idx) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        // lexicographic (dist, index) == the sequential strict-'<' first-min rule; idx < 0 = "no centroid seen yet"
        const bool take = (oi >= 0) && (idx < 0 || ob < best || (ob == best && oi < idx));
        if (take) { best = ob; idx = oi; }
    }
}

// One warp owns R rows at a time (staged in shared memory); lane l evaluates centroids l, l+32, ... for all R rows,
// so every centroid element fetched from global memory (coalesced across lanes: cT is [d][k]) is used R times.
template <int R>
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ x, long long ld, int col0, long long n, int d, const float* __restrict__ cT, int k,
                     int* __restrict__ assign, float* __restrict__ dist) {
    extern __shared__ float s_x[];  // [warps][R][d]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* xs = s_x + (size_t)w * R * d;
    const long long groups = (n + R - 1) / R;
    for (long long g = (long long)blockIdx.x * nw + w; g < groups; g += (long long)gridDim.x * nw) {
        const long long row0 = g * R;
        for (int i = lane; i < R * d; i += 32) {
            const int r = i / d, t = i - r * d;
            xs[i] = (row0 + r < n) ? x[(row0 + r) * ld + col0 + t] : 0.0f;
        }
        __syncwarp();
        float best[R];
        int idx[R];
#pragma unroll
        for (int r = 0; r < R; r++) { best[r] = 4294967296.0f; idx[r] = -1; }  // (float)UINT_MAX, IVFOPQ.cpp:111
        for (int c = lane; c < k; c += 32) {
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll 4
            for (int t = 0; t < d; t++) {
                const float cv = __ldg(cT + (long long)t * k + c);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float df = __fsub_rn(xs[r * d + t], cv);
                    acc[r] = __fadd_rn(acc[r], __fmul_rn(df, df));
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++)
                if (acc[r] < best[r]) { best[r] = acc[r]; idx[r] = c; }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            km_warp_argmin(best[r], idx[r]);
            if (lane == 0 && row0 + r < n) { assign[row0 + r] = idx[r]; dist[row0 + r] = best[r]; }
        }
        __syncwarp();
    }
}

__global__ void kmeans_partial_kernel(const float* __restrict__ x, long long ld, int col0, int d, const int* __restrict__ row_sorted,
                                      const long long* __restrict__ blk_lo, const long long* __restrict__ blk_hi, long long n_blocks,
                                      double* __restrict__ partial) {
    const long long total = n_blocks * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / d;
        const int t = (int)(i - b * d);
        double acc = 0.0;
        const long long hi = blk_hi[b];
        for (long long p = blk_lo[b]; p < hi; p++) acc = __dadd_rn(acc, (double)__ldg(x + (long long)row_sorted[p] * ld + col0 + t));
        partial[i] = acc;
    }
}

__global__ void kmeans_finalize_kernel(const double* __restrict__ partial, const long long* __restrict__ cl_blk_off,
                                       const int* __restrict__ count, int d, int k, float* __restrict__ c, float* __restrict__ cT) {
    const long long total = (long long)k * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / d), t = (int)(i - (long long)j * d);
        const int cnt = count[j];
        if (cnt <= 0) continue;
        double sum = 0.0;
        for (long long b = cl_blk_off[j]; b < cl_blk_off[j + 1]; b++) sum = __dadd_rn(sum, partial[b * d + t]);
        const float v = (float)__ddiv_rn(sum, (double)cnt);
        c[i] = v;
        cT[(long long)t * k + j] = v;
    }
}

__global__ void kmeans_reseed_kernel(const float* __restrict__ x, long long ld, int col0, int d, const int* __restrict__ donors,
                                     const int* __restrict__ empties, int n_empty, int k, float* __restrict__ c, float* __restrict__ cT) {
    const long long total = (long long)n_empty * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / d), t = (int)(i - (long long)e * d);
        const int j = empties[e];
        const float v = x[(long long)donors[e] * ld + col0 + t];
        c[(long long)j * d + t] = v;
        cT[(long long)t * k + j] = v;
    }
}

__global__ void residual_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ coarse,
                                const int* __restrict__ assign, float* __restrict__ out) {
    const long long total = n * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D;
        const int t = (int)(i - r * D);
        out[i] = __fsub_rn(x[i], __ldg(coarse + (long long)assign[r] * D + t));
    }
}

inline unsigned grid_for(long long work, int sm_count) {
    return (unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, (long long)sm_count * 8));
}

}  // namespace

int launch_kmeans_assign(Ctx* ctx, const float* x, long long ld, int col0, long long n, int d, const float* cT, int k, int* assign,
                         float* dist) {
    if (n <= 0) return 0;
    if (d < 1 || d > 4096) B2_FAIL(-4, "kmeans: dimension must be in [1, 4096]");
    if (tiled_nearest_pays(d, k)) return launch_tiled_nearest(ctx, x, ld, col0, n, d, cT, k, 0, 1, assign, dist);
    // rows per warp and warps per CTA from the shared-memory budget (<= 48 KB, no opt-in needed)
    const int R = (d <= 384) ? 4 : 1;
    int warps = 8;
    while (warps > 1 && (size_t)warps * R * d * sizeof(float) > 48 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * R * d * sizeof(float);
    const long long groups = (n + R - 1) / R;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((groups + warps - 1) / warps, (long long)ctx->sm_count * 8));
    if (R == 4)
        kmeans_assign_kernel<4><<<grid, warps * 32, smem, ctx->stream>>>(x, ld, col0, n, d, cT, k, assign, dist);
    else
        kmeans_assign_kernel<1><<<grid, warps * 32, smem, ctx->stream>>>(x, ld, col0, n, d, cT, k, assign, dist);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_kmeans_partial(Ctx* ctx, const float* x, long long ld, int col0, int d, const int* row_sorted, const long long* blk_lo,
                          const long long* blk_hi, long long n_blocks, double* partial) {
    if (n_blocks <= 0) return 0;
    kmeans_partial_kernel<<<grid_for(n_blocks * d, ctx->sm_count), 256, 0, ctx->stream>>>(x, ld, col0, d, row_sorted, blk_lo, blk_hi, n_blocks,
                                                                                         partial);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_kmeans_finalize(Ctx* ctx, const double* partial, const long long* cl_blk_off, const int* count, int d, int k, float* c,
                           float* cT) {
    kmeans_finalize_kernel<<<grid_for((long long)k * d, ctx->sm_count), 256, 0, ctx->stream>>>(partial, cl_blk_off, count, d, k, c, cT);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_kmeans_reseed(Ctx* ctx, const float* x, long long ld, int col0, int d, const int* donors, const int* empties, int n_empty,
                         int k, float* c, float* cT) {
    if (n_empty <= 0) return 0;
    kmeans_reseed_kernel<<<grid_for((long long)n_empty * d, ctx->sm_count), 256, 0, ctx->stream>>>(x, ld, col0, d, donors, empties, n_empty, k,
                                                                                                 c, cT);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_residual(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* assign, float* out) {
    if (n <= 0) return 0;
    residual_kernel<<<grid_for(n * D, ctx->sm_count), 256, 0, ctx->stream>>>(x, n, D, coarse, assign, out);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
