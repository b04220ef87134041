// train_kernels.cuh -- k-means training kernels (f-4; definitions in train_kernels.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

constexpr int KM_SUM_BLOCK = 512;  // rows per summation block of the centroid update (part of the algorithm's definition)

// assign[i] = argmin_j sum_t (x[i*ld + col0 + t] - c[j][t])^2 (sequential fp32, first minimum wins), dist[i] = that minimum.
// cT = centroids transposed to [d][k].
int launch_kmeans_assign(Ctx* ctx, const float* x, long long ld, int col0, long long n, int d, const float* cT, int k, int* assign,
                         float* dist);
// partial[b][t] = sum in double, in list order, of x[row_sorted[i]][t] over i in [blk_lo[b], blk_hi[b])
int launch_kmeans_partial(Ctx* ctx, const float* x, long long ld, int col0, int d, const int* row_sorted, const long long* blk_lo,
                          const long long* blk_hi, long long n_blocks, double* partial);
// centroid j = (float)(sum of its partials in block order / count[j]); clusters with count 0 are left as they are.
// Writes both the row-major [k][d] copy (element stride c_ld, column offset 0) and the transposed [d][k] copy.
int launch_kmeans_finalize(Ctx* ctx, const double* partial, const long long* cl_blk_off, const int* count, int d, int k, float* c,
                           float* cT);
// centroid empties[e] = row donors[e] of x
int launch_kmeans_reseed(Ctx* ctx, const float* x, long long ld, int col0, int d, const int* donors, const int* empties, int n_empty,
                         int k, float* c, float* cT);
// out[i][t] = x[i][t] - coarse[assign[i]][t]   (CoarseQuan's residue, train_PQ_codebook.cpp:185-192)
int launch_residual(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* assign, float* out);

}  // namespace b200nn
