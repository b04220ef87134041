// dist_tile.cuh -- register-tiled all-pairs squared distances + nearest-centroid selection (definitions in dist_tile.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

// worth it from here on (below, the one-warp-per-row kernels are used)
bool tiled_nearest_pays(int d, int K);
// For every row i < n of x (columns [col0, col0+d) of a matrix with leading dimension ld) against the K centroids cT [d][K]:
//   rule 0 (a2, IVFOPQ.cpp:107-129): out_idx[i] = first minimum of the sequential fp32 squared distance (-1 if none is below
//           (float)UINT_MAX), out_dist[i] (optional) = that minimum;
//   rule 1 (a4, IVFOPQ.cpp:238-260): out_idx[i*nk .. ) = the nk smallest under (dist, index) in the reference's pop order.
int launch_tiled_nearest(Ctx* ctx, const float* x, long long ld, int col0, long long n, int d, const float* cT, int K, int rule, int nk,
                         int* out_idx, float* out_dist);

// a3 on the register tile (ksub = 256, D/M in {4, 8, 16}): codes[r*M + m] = first minimum over the 256 codewords of sub-quantizer m
// of the sequential fp32 squared distance to the residual x - coarse[list[r]]; cbT = codebooks transposed to [M][D/M][256]
bool pq_encode_tile_supported(int ds, int ksub);
int launch_pq_encode_tile(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* list, const float* cbT, int M, int ksub,
                          unsigned char* codes);
// the context's distance scratch matrix (at least `elems` floats)
int ensure_dmat(Ctx* ctx, size_t elems);

}  // namespace b200nn
