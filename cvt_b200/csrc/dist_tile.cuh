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

}  // namespace b200nn
