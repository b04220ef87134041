// pq_kernels.cuh -- launchers of the (O)PQ kernels (definitions in pq_kernels.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace b200nn {

int launch_rotate_perm(Ctx* ctx, const float* x, long long n, int D, const int* perm, float* y);
int launch_coarse_assign(Ctx* ctx, const float* x, long long n, int D, const float* coarseT, int K, int* out_list);
int launch_pq_encode(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* list, const float* cbT,
                     int M, int ksub, unsigned char* codes);
int launch_codes_to_scan_layout(Ctx* ctx, const unsigned char* codes, long long n, int M, uint32_t* codesT, long long n_pad);
int launch_coarse_probe(Ctx* ctx, const float* q, long long nq, int D, const float* coarseT, int K, int nk, int* out_lists);
int launch_lut_build_std(Ctx* ctx, const float* q, long long nq, int D, const int* probes, int nprobe, const float* coarse,
                         const float* cb, int M, int ksub, float* lut);
int launch_lut_build_scan(Ctx* ctx, int M, const float* q, long long nq, int D, const float* centroid, const float* cb,
                          float* lut_scan);
int scan_queries_per_cta(int M);
// Work plan of the fused scan: n_full whole-shard CTAs + n_tail tail CTAs; desc = 8 ints per tail CTA
// (two {query group, output slice, granule lo, granule hi} segments); slices = output lists per query.
struct ScanPlan {
    int n_full = 0, n_tail = 0, slices = 1;
    std::vector<int> desc;
};
void scan_plan(int sm_count, long long qgroups, long long n_granules, ScanPlan* plan);
// out_keys: [plan.slices][qgroups*QW][k]; must be pre-filled with 0xFF bytes when plan.slices > 1
size_t scan_warm_scratch_floats(int M, int n_full, int n_tail);
int launch_adc_scan_topk(Ctx* ctx, int M, const uint32_t* codesT, const float* lut_scan, long long n_rows, long long qgroups,
                         int n_full, int n_tail, const int* tail_desc_dev /* plan.desc on the device */, int k, float clamp,
                         uint32_t id_base, unsigned long long* out_keys, float* warm_scratch /* scan_warm_scratch_floats() floats */);
int launch_fill_f32(Ctx* ctx, float* p, long long n, float v);
// f-5: total[g] = sequential fp32 sum over frames of scores[f][g]; k smallest (value, index) per row of a dense matrix
int launch_frame_sum(Ctx* ctx, const float* scores, int n_frames, long long ng, float* total);
int launch_dense_topk(Ctx* ctx, const float* values, long long rows, long long n, int k, unsigned long long* out_keys);
// row stride ld, `slices` column slices per row (out_keys [slices][rows][k]), ids = id_map[column] when given
int launch_dense_topk_ex(Ctx* ctx, const float* values, long long rows, long long n, long long ld, int k, int slices, const uint32_t* id_map,
                         unsigned long long* out_keys);
int launch_ivf_scan(Ctx* ctx, const float* lut, const int* probes, const long long* list_off, const unsigned char* codes_sorted,
                    const int* slot_sorted, int M, int ksub, long long nq, int nprobe, long long out_stride, float* out);
int launch_ivf_search_topk(Ctx* ctx, const float* q_rot, long long nq, int D, const int* probes, int nprobe, const float* coarse,
                           const float* cb, int M, int ksub, const long long* list_off, const unsigned char* codes_sorted,
                           const int* row_sorted, long long n_rows, int k, float clamp, uint32_t id_base, unsigned long long* out_keys);

}  // namespace b200nn
