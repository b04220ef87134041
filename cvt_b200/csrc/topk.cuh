// topk.cuh -- exact top-k selection building blocks shared by every scan kernel.
//
// Order = the reference's: k smallest under lexicographic (dist, id)
// (std::partial_sort_copy over pair<float,uint>, opq/src/common.h:25-37; the max-heap push/pop
// rule of BruteforceSearch::searchKnn, brute_force_search/src/brutoforce.hpp:73-93).
// Records are 64-bit keys (orderable(dist) << 32 | id), compared as unsigned integers.
//
// Structure: every scan kernel filters candidates against a per-query threshold tau (the current
// k-th best), parks the rare survivors in a small per-warp staging buffer and, when that fills,
// merges it into the CTA's sorted per-query list (KP slots, shared memory) under a per-query lock.
#pragma once
#include "common.cuh"

namespace b200nn {

constexpr int KP = 128;  // list slots per query; the fused kernels support k <= KP

// Merge nb (<= 32) distinct candidate keys at cand_addr (shared, unsorted) into the ascending
// sorted list of KP keys at L_addr (shared; unused slots = KEY_MAX).  Rank based: every element
// computes its position in the union; nothing is sorted.  All 32 lanes must call.
__device__ __forceinline__ void warp_list_merge(uint32_t L_addr, uint32_t cand_addr, int nb) {
    const int lane = threadIdx.x & 31;
    unsigned long long l[KP / 32];
    int pl[KP / 32];
#pragma unroll
    for (int t = 0; t < KP / 32; t++) {
        l[t] = lds64(L_addr + (uint32_t)(lane + 32 * t) * 8u);
        pl[t] = 0;
    }
    const unsigned long long c = lane < nb ? lds64(cand_addr + (uint32_t)lane * 8u) : KEY_MAX;
    int pc = 0;
    for (int j0 = 0; j0 < nb; j0 += 4) {  // 4 broadcast reads in flight per step
        unsigned long long cj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cj[u] = (j0 + u < nb) ? lds64(cand_addr + (uint32_t)(j0 + u) * 8u) : KEY_MAX;
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int t = 0; t < KP / 32; t++) pl[t] += (cj[u] < l[t]) ? 1 : 0;
            pc += (cj[u] < c) ? 1 : 0;
        }
    }
    // number of list entries < c (lower bound in the sorted list; KP is a power of two)
    int pos = 0;
#pragma unroll
    for (int step = KP / 2; step >= 1; step >>= 1)
        if (lds64(L_addr + (uint32_t)(pos + step - 1) * 8u) < c) pos += step;
    if (lds64(L_addr + (uint32_t)pos * 8u) < c) pos += 1;
    pc += pos;
    __syncwarp();  // all reads of the old list are done
#pragma unroll
    for (int t = 0; t < KP / 32; t++) {
        const int p = lane + 32 * t + pl[t];
        if (p < KP) sts64(L_addr + (uint32_t)p * 8u, l[t]);
    }
    if (lane < nb && pc < KP) sts64(L_addr + (uint32_t)pc * 8u, c);
    __syncwarp();
}

// Locked flush of a staging buffer into the CTA list of one query; publishes the new threshold
// *tau_key = the k-th best record (KEY_MAX while the list holds fewer than k records).
__device__ __forceinline__ void warp_flush(uint32_t L_addr, int* lock, volatile unsigned long long* tau_key,
                                           uint32_t cand_addr, int nb, int k, bool lock_held = false) {
    const int lane = threadIdx.x & 31;
    if (nb <= 0) {  // warp-uniform
        if (lock_held && lane == 0) atomicExch(lock, 0);
        return;
    }
    if (lane == 0 && !lock_held) {
        while (atomicCAS(lock, 0, 1) != 0) {
        }
        __threadfence_block();
    }
    __syncwarp();
    warp_list_merge(L_addr, cand_addr, nb);
    if (lane == 0) {
        *tau_key = lds64(L_addr + (uint32_t)(k - 1) * 8u);
        __threadfence_block();
        atomicExch(lock, 0);
    }
    __syncwarp();
}

// float view of a threshold record: the k-th best distance, +inf while the list is not full
__device__ __forceinline__ float tau_f32_of(unsigned long long tkey) {
    return tkey == KEY_MAX ? __int_as_float(0x7f800000) : f32_from_orderable((uint32_t)(tkey >> 32));
}

// Merge of L sorted key lists per query (defined in topk_merge.cu).  keys[chunk][l][list_stride/k][k],
// chunk = q / (list_stride/k); a single chunk is keys[l*list_stride + q*k + j].
int launch_topk_merge(Ctx* ctx, const unsigned long long* keys, int L, long long nq, int k, long long list_stride,
                      float* out_dist_f, int* out_dist_i, unsigned long long* out_id, unsigned long long* out_key);

// the same merge with the L lists given as L device pointers ([nq][k] each; may point into peer GPUs' memory)
int launch_topk_merge_ptrs(Ctx* ctx, const unsigned long long* const* list_ptrs_dev, int L, long long nq, int k, float* out_dist_f,
                           int* out_dist_i, unsigned long long* out_id, unsigned long long* out_key);

}  // namespace b200nn
