// flat_kernels.cuh -- launchers of the exact-scan kernels (definitions in flat_kernels.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

int flat_pick_slices(int sm_count, long long nq, long long n);
// keys out: [n_slices][nq][k]; ids inside the keys are label RANKS
int launch_flat_scan(Ctx* ctx, int metric, int order, const void* data, const uint32_t* rank, long long n, int d,
                     const void* queries, long long nq, int n_slices, int k, unsigned long long* out_keys);
// fp32 metrics (0: 1 - <q,x>, 1: L2^2) on the register-tiled kernel (flat_tile.cu): rows are cut into n_chunks row chunks x
// `slices` selection slices; keys out [n_chunks * slices][nq][k]
void flat_f32_plan(int sm_count, long long nq, long long n, long long* chunk_rows, int* n_chunks, int* slices);
int launch_flat_scan_f32(Ctx* ctx, int metric, int order, const float* data, const uint32_t* rank, long long n, int d, const float* queries,
                         long long nq, int k, unsigned long long* out_keys);
int launch_rank_to_label(Ctx* ctx, unsigned long long* ids, long long count, const unsigned long long* label_sorted);

}  // namespace b200nn
