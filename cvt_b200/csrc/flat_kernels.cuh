// flat_kernels.cuh -- launchers of the exact-scan kernels (definitions in flat_kernels.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

int flat_pick_slices(int sm_count, long long nq, long long n);
// keys out: [n_slices][nq][k]; ids inside the keys are label RANKS
int launch_flat_scan(Ctx* ctx, int metric, int order, const void* data, const uint32_t* rank, long long n, int d,
                     const void* queries, long long nq, int n_slices, int k, unsigned long long* out_keys);
int launch_rank_to_label(Ctx* ctx, unsigned long long* ids, long long count, const unsigned long long* label_sorted);

}  // namespace b200nn
