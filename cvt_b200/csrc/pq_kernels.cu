// pq_kernels.cu -- (O)PQ / IVFOPQ hot path on sm_100a:
//   rotate (permutation gather), coarse assign, PQ argmin encode, LUT build, the TMA-staged
//   conflict-free ADC scan with fused exact top-k, and the generic IVF (list-probing) scan.
//
// Arithmetic contract: every squared distance is the reference's sequential
// fp32 loop `tmp = a-b; acc += tmp*tmp` (IVFOPQ.cpp:117-122,147-154,283-288) written with
// __fsub_rn/__fmul_rn/__fadd_rn so nothing is contracted into FMAs; every ADC score is the
// sequential fp32 sum over m = 0..M-1 starting from 0.0f (IVFOPQ.cpp:302-306).  Codes, LUTs and
// scores are therefore bit-identical to the reference's.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "dist_tile.cuh"
#include "pq_kernels.cuh"
#include "topk.cuh"

namespace b200nn {

// =============================================================================================
// a1. rotation as a permutation: y[r][i] = x[r][perm[i]]   (IVFOPQ::reorder, IVFOPQ.cpp:424-439)
// =============================================================================================
__global__ void rotate_perm_kernel(const float* __restrict__ x, long long n, int D, const int* __restrict__ perm,
                                   float* __restrict__ y) {
    const long long total = n * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D;
        const int c = (int)(i - r * D);
        y[i] = __ldg(x + r * D + perm[c]);
    }
}

// =============================================================================================
// a2. coarse assignment: argmin_i sum_j (x_j - C[i][j])^2, first minimum wins
//     (IVFOPQ::Add, IVFOPQ.cpp:107-129).  One warp per row; lane l evaluates centroids
//     l, l+32, ... (each a sequential fp32 sum), then a warp-shuffle argmin on (dist, index).
//     coarseT is the centroid table transposed to [D][K] so that lanes read consecutive addresses.
// =============================================================================================
__device__ __forceinline__ void warp_argmin(float& best, int& idx) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        // lexicographic (dist, index): identical to the sequential strict-'<' first-min rule.
        // idx == -1 marks "nothing beat UINT_MAX yet" and never wins against a real index.
        const bool take = (oi >= 0) && (idx < 0 || ob < best || (ob == best && oi < idx));
        if (take) { best = ob; idx = oi; }
    }
}

__global__ void coarse_assign_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ coarseT,
                                     int K, int* __restrict__ out_list) {
    extern __shared__ float s_x[];  // [warps][D]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* xr = s_x + w * D;
    for (long long row = (long long)blockIdx.x * nw + w; row < n; row += (long long)gridDim.x * nw) {
        for (int j = lane; j < D; j += 32) xr[j] = x[row * D + j];
        __syncwarp();
        float best = 4294967296.0f;  // (float)UINT_MAX, IVFOPQ.cpp:111
        int idx = -1;
        for (int c = lane; c < K; c += 32) {
            float acc = 0.0f;
            for (int j = 0; j < D; j++) {
                const float t = __fsub_rn(xr[j], __ldg(coarseT + (long long)j * K + c));
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            if (acc < best) { best = acc; idx = c; }
        }
        warp_argmin(best, idx);
        if (lane == 0) out_list[row] = idx;
        __syncwarp();
    }
}

// =============================================================================================
// a3. residual + PQ argmin encode (IVFOPQ::Add, IVFOPQ.cpp:135-163).  One warp per (row, m):
//     lane l evaluates codewords l, l+32, ... sequentially over d_sub, then shuffle argmin.
//     cbT = codebooks transposed to [M][d_sub][ksub] (lanes read consecutive codewords).
// =============================================================================================
template <int DS>
__global__ void pq_encode_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ coarse,
                                 const int* __restrict__ list, const float* __restrict__ cbT, int M, int ksub,
                                 unsigned char* __restrict__ codes) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = n * M;
    for (long long p = warp; p < total; p += nwarps) {
        const long long row = p / M;
        const int m = (int)(p - row * M);
        const int vw = list[row];
        float res[DS];
#pragma unroll
        for (int k = 0; k < DS; k++)  // feat_residual = x - centroid, IVFOPQ.cpp:136-139
            res[k] = __fsub_rn(__ldg(x + row * D + m * DS + k), __ldg(coarse + (long long)(vw < 0 ? 0 : vw) * D + m * DS + k));
        float best = 4294967296.0f;  // (float)UINT_MAX, IVFOPQ.cpp:143
        int idx = -1;
        const float* cb = cbT + (long long)m * DS * ksub;
        for (int j = lane; j < ksub; j += 32) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < DS; k++) {
                const float t = __fsub_rn(res[k], __ldg(cb + k * ksub + j));
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            if (acc < best) { best = acc; idx = j; }
        }
        warp_argmin(best, idx);
        if (lane == 0) codes[p] = (unsigned char)idx;  // elem.PQindex[i] = vw1, IVFOPQ.cpp:161
    }
}

// =============================================================================================
// Scan layout of the coded database in HBM.  Rows are grouped in granules of 64; inside a
// granule the M = 4G code bytes of a row are split into G little-endian 32-bit words (word h =
// bytes of sub-quantizers 4h..4h+3) and stored plane-major:
//     codesT[granule][h][row_in_granule]   (uint32)
// so that the scan can move one plane of a stage with a single 1-D TMA copy and a lane group can
// fetch the words of 4 consecutive rows with one 16-byte shared-memory load.  Rows beyond n are
// zero bytes.
// =============================================================================================
__global__ void codes_to_scan_layout_kernel(const unsigned char* __restrict__ codes, long long n, int M,
                                            uint32_t* __restrict__ codesT, long long n_pad) {
    const int G = M / 4;
    const long long total = n_pad * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long gran = i / (64 * G);
        const int rem = (int)(i - gran * 64 * G);
        const int h = rem / 64, r = rem - h * 64;
        const long long row = gran * 64 + r;
        uint32_t w = 0;
        if (row < n) {
            const unsigned char* c = codes + row * M + 4 * h;
            w = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
        }
        codesT[i] = w;
    }
}

// =============================================================================================
// a4. coarse probe selection (IVFOPQ::Query, IVFOPQ.cpp:238-260): the nk smallest centroids under
//     (dist, index), emitted in the reference's pop order (largest first).  One warp per query.
// =============================================================================================
__global__ void coarse_probe_kernel(const float* __restrict__ q, long long nq, int D, const float* __restrict__ coarseT,
                                    int K, int nk, int* __restrict__ out_lists) {
    extern __shared__ float s_q[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* xr = s_q + w * D;
    constexpr int MAXNK = 8;
    for (long long row = (long long)blockIdx.x * nw + w; row < nq; row += (long long)gridDim.x * nw) {
        for (int j = lane; j < D; j += 32) xr[j] = q[row * D + j];
        __syncwarp();
        float bd[MAXNK];
        int bi[MAXNK];
#pragma unroll
        for (int t = 0; t < MAXNK; t++) { bd[t] = __int_as_float(0x7f800000); bi[t] = 0x7fffffff; }
        for (int c = lane; c < K; c += 32) {
            float acc = 0.0f;
            for (int j = 0; j < D; j++) {
                const float t = __fsub_rn(xr[j], __ldg(coarseT + (long long)j * K + c));
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            // insert into the lane-local ascending list (indices ascend, so ties keep order)
            float d = acc; int id = c;
#pragma unroll
            for (int t = 0; t < MAXNK; t++) {
                if (t < nk && (d < bd[t] || (d == bd[t] && id < bi[t]))) {
                    const float td = bd[t]; const int ti = bi[t];
                    bd[t] = d; bi[t] = id; d = td; id = ti;
                }
            }
        }
        // pop the global minimum nk times
        for (int r = 0; r < nk; r++) {
            float d = bd[0]; int id = bi[0];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, d, s);
                const int oi = __shfl_xor_sync(0xffffffffu, id, s);
                if (od < d || (od == d && oi < id)) { d = od; id = oi; }
            }
            if (bi[0] == id && bd[0] == d) {  // the owning lane advances its list
#pragma unroll
                for (int t = 0; t + 1 < MAXNK; t++) { bd[t] = bd[t + 1]; bi[t] = bi[t + 1]; }
                bd[MAXNK - 1] = __int_as_float(0x7f800000); bi[MAXNK - 1] = 0x7fffffff;
            }
            if (lane == 0) out_lists[row * nk + (nk - 1 - r)] = id;  // reference pops largest first
        }
        __syncwarp();
    }
}

// =============================================================================================
// a5. residual + LUT build (IVFOPQ::Query, IVFOPQ.cpp:269-291).  One thread per (m, j) entry,
//     sequential over d_sub.  Two output layouts:
//       standard  lut[q][probe][m][j]
//       scan      lut_scan[qgroup][pair][j][msel][lane]  (see adc_scan_topk_kernel)
// =============================================================================================
template <int DS>
__global__ void lut_build_std_kernel(const float* __restrict__ q, long long nq, int D, const int* __restrict__ probes,
                                     int nprobe, const float* __restrict__ coarse, const float* __restrict__ cb, int M,
                                     int ksub, float* __restrict__ lut) {
    extern __shared__ float s_res[];  // [D]
    const long long qp = blockIdx.x;  // query * nprobe + probe
    const long long qi = qp / nprobe;
    const int vw = probes ? probes[qp] : 0;
    for (int j = threadIdx.x; j < D; j += blockDim.x) s_res[j] = __fsub_rn(q[qi * D + j], coarse[(long long)vw * D + j]);
    __syncthreads();
    for (int e = threadIdx.x; e < M * ksub; e += blockDim.x) {
        const int m = e / ksub, j = e - m * ksub;
        const float* c = cb + ((long long)m * ksub + j) * DS;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < DS; k++) {
            const float t = __fsub_rn(s_res[m * DS + k], __ldg(c + k));
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
        lut[qp * M * ksub + e] = acc;
    }
}

template <int G, int DS>
__global__ void lut_build_scan_kernel(const float* __restrict__ q, long long nq, int D, const float* __restrict__ centroid,
                                      const float* __restrict__ cb, float* __restrict__ lut_scan) {
    constexpr int QW = 32 / G, M = 4 * G;
    extern __shared__ float s_res[];  // [QW][D+4]
    const int RS = D + 4;
    const long long qg = blockIdx.x;
    for (int i = threadIdx.x; i < QW * D; i += blockDim.x) {
        const int ql = i / D, j = i - ql * D;
        const long long qi = qg * QW + ql;
        s_res[ql * RS + j] = qi < nq ? __fsub_rn(q[qi * D + j], centroid[j]) : 0.0f;
    }
    __syncthreads();
    float* out = lut_scan + qg * 32768;
    // One thread = one codeword (m, j) against all QW queries of the group: the codeword row is loaded once (lanes read
    // consecutive rows), the residual sub-vectors are warp-uniform shared-memory broadcasts, and the QW results are
    // adjacent in the scan layout (one run of QW floats).  blockIdx.y = which slice of the M*256 codewords: more CTAs than
    // query groups, so that a small batch (one rank's chunk of a sharded batch) still fills the machine.
    const int t_per = M * 256 / (int)gridDim.y;
    for (int t = (int)blockIdx.y * t_per + threadIdx.x; t < ((int)blockIdx.y + 1) * t_per; t += blockDim.x) {
        const int m = t >> 8, j = t & 255;
        const int h = m >> 2, pair = (m >> 1) & 1, msel = m & 1;
        const float* c = cb + ((long long)m * 256 + j) * DS;
        float cv[DS];
        if (DS % 4 == 0) {
#pragma unroll
            for (int k = 0; k < DS; k += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(c + k));
                cv[k] = v.x; cv[k + 1] = v.y; cv[k + 2] = v.z; cv[k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < DS; k++) cv[k] = __ldg(c + k);
        }
        float acc[QW];
#pragma unroll
        for (int ql = 0; ql < QW; ql++) {
            const float* r = s_res + ql * RS + m * DS;
            float a = 0.0f;
            if (DS % 4 == 0) {
#pragma unroll
                for (int k = 0; k < DS; k += 4) {
                    const float4 rv = *reinterpret_cast<const float4*>(r + k);
                    float d;
                    d = __fsub_rn(rv.x, cv[k]); a = __fadd_rn(a, __fmul_rn(d, d));
                    d = __fsub_rn(rv.y, cv[k + 1]); a = __fadd_rn(a, __fmul_rn(d, d));
                    d = __fsub_rn(rv.z, cv[k + 2]); a = __fadd_rn(a, __fmul_rn(d, d));
                    d = __fsub_rn(rv.w, cv[k + 3]); a = __fadd_rn(a, __fmul_rn(d, d));
                }
            } else {
#pragma unroll
                for (int k = 0; k < DS; k++) {
                    const float d = __fsub_rn(r[k], cv[k]);
                    a = __fadd_rn(a, __fmul_rn(d, d));
                }
            }
            acc[ql] = (qg * QW + ql < nq) ? a : 0.0f;
        }
        float4* o4 = reinterpret_cast<float4*>(out + pair * 16384 + j * 64 + msel * 32 + h * QW);
#pragma unroll
        for (int ql = 0; ql < QW; ql += 4) o4[ql >> 2] = make_float4(acc[ql], acc[ql + 1], acc[ql + 2], acc[ql + 3]);
    }
}

// =============================================================================================
// a6 + a7.  THE hot kernel: flat ADC scan with fused exact top-k.
//
// Work unit: one CTA = (group of QW = 32/G queries) x (a slice of the coded database).
// Lane layout inside every warp: lane = h*QW + q, h = lane group owning sub-quantizers
// 4h..4h+3, q = query within the group.  The CTA keeps the QW queries' LUTs resident in shared
// memory for its whole life (128 KB, loaded once with TMA bulk copies) laid out so that the
// word lane l reads always sits in bank l:
//     byte offset = pair*65536 + code*256 + msel*128 + lane*4      (m = 4h + 2*pair + msel)
// => every LUT gather instruction is 32 lookups in ONE conflict-free shared-memory wavefront,
// no matter what the code bytes are (code bytes are warp-uniform per lane group, they only
// select the row).  The LUT base is 64 KB aligned in the shared window so a single PRMT builds
// the complete address (byte0 = lane*4, byte1 = code byte, bytes 2-3 = base).
//
// A lane only owns 4 of the M sub-quantizers, so the running sum of a row travels through the
// lane groups like a systolic pipeline: group h adds its 4 entries (in order) to the value
// received from group h-1 with one warp shuffle, one block (4 rows) later.  The fp32 additions
// therefore happen in exactly the reference's order m = 0..M-1 starting from 0.0f
// (IVFOPQ.cpp:302-306) and the scores are bit-identical; the last group owns the final scores.
//
// Code bytes stream through a per-warp ring in shared memory filled by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier); each lane group reads the 32-bit code words of 4 consecutive rows
// with one 16-byte ld.shared.
//
// Top-k: final scores are compared with the query's running threshold; the rare survivors go to
// a per-warp staging buffer and are merged into the CTA's sorted per-query list (topk.cuh).
// Each CTA emits k sorted keys per query; topk_merge_kernel merges the slices.
// =============================================================================================
// VAR bit 1 (grouped check): the threshold test is made once per 4 blocks instead of once per block (see group_body).
// VAR bit 0 (used for G = 8, where three 64-row stages of 8 planes per warp do not fit beside the LUT): a ring of TWO
// whole-granule stages whose prefetch is issued in the MIDDLE of the current stage, once the lagging lane groups
// (G-1 <= 8 blocks behind) have left the previous one -- half as many TMA copies (256 B instead of 128 B each) and half
// as many stage boundaries per row as the three 32-row stages of VAR = 0.
template <int G, int WARPS_, int VAR = 0>
struct ScanCfg {
    static constexpr int QW = 32 / G;                          // queries per CTA
    static constexpr int WARPS = WARPS_;
    static constexpr int STAGE_BLOCKS = (G <= 4 || (VAR & 1)) ? 16 : 8;     // blocks (of 4 rows) per ring stage (> G-1 lag)
    static constexpr int STAGE_ROWS = STAGE_BLOCKS * 4;
    static constexpr int RING_STAGES = (VAR & 1) ? 2 : 3;     // previous (lagging lane groups) | current | prefetch
    static constexpr int PLANE_RING_BYTES = RING_STAGES * STAGE_ROWS * 4;  // per lane group
    static constexpr int RING_BYTES = G * PLANE_RING_BYTES;                // per warp
    // staging records per (warp, query): as many as shared memory allows (99 KB beside the LUT)
    static constexpr int SB = (G == 1) ? 8 : (G == 2 ? 24 : 32);
    // try to merge once half a buffer is staged (measured: merging much earlier costs more in merges
    // than the fresher threshold saves in candidates)
    static constexpr int SOFT = (G == 1) ? 4 : 10;  // measured flat optimum 8..12 (B200NN_SOFT overrides for tuning)
    static constexpr int STAGING_BYTES = QW * SB * 8;          // per warp
    static constexpr int LIST_BYTES = QW * KP * 8;
    // Warm-up: the first WARM_STAGES stages of every warp are not filtered at all -- their raw scores are
    // parked in global scratch and the CTA then selects their top-k in one dense pass (lanes over rows),
    // which is far cheaper per record than the streaming candidate path that a k(1+ln(n/k)) warm-up
    // would otherwise take for exactly these rows.
    static constexpr int WARM_STAGES = (STAGE_BLOCKS == 16) ? 2 : 4;
    static constexpr int WARM_BLOCKS = WARM_STAGES * STAGE_BLOCKS;   // 32 blocks = 128 rows per warp
    static constexpr int WARM_ROWS = WARM_BLOCKS * 4;
    static constexpr int LUT_BYTES = 131072;
    static constexpr int SMEM_BYTES = 232448;                  // 227 KB: one CTA per SM
};

template <int IMM>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}

// first look-up of a lane group: S = R*keep + lut[..]  (keep = 0 for lane group 0, which starts the
// sum at 0.0f, else 1: R*1 + x is R + x with a single rounding, R*0 + x is exactly x)
#define B2_LOOKUP4(S, R, W)                                                                   \
    S = __fmaf_rn(R, keep, lds_f32_off<0>(__byte_perm(W, basereg, 0x7604)));                    \
    S = __fadd_rn(S, lds_f32_off<128>(__byte_perm(W, basereg, 0x7614)));                        \
    S = __fadd_rn(S, lds_f32_off<65536>(__byte_perm(W, basereg, 0x7624)));                      \
    S = __fadd_rn(S, lds_f32_off<65536 + 128>(__byte_perm(W, basereg, 0x7634)));

// Work decomposition: the first n_full query groups are scanned whole by one CTA each (whole waves
// of one-CTA-per-SM).  The remaining groups (the partial last wave) form one linear stretch of
// (group, granule) work that is cut into equal pieces, one per tail CTA, so that the last wave fills
// the machine exactly whatever the group count is; a piece may straddle the boundary between two
// groups, in which case the CTA runs two segments (own LUT, own top-k lists, own output slice).
// Every (query, segment) pair pays a top-k warm-up of ~k(1 + ln(rows/k)) insertions, so pieces are
// used only where they buy balance.  tail_desc holds two {group, slice, granule lo, granule hi}
// records per tail CTA (scan_plan); the second is empty (lo >= hi) when the piece lies in one group.
template <int G, int WARPS_, bool STATS, int VAR = 0>
__global__ void __launch_bounds__(WARPS_ * 32, 1)
adc_scan_topk_kernel(const uint32_t* __restrict__ codesT,  // [granules][G][64]
                     const float* __restrict__ lut_scan,    // [qgroups][32768]
                     long long n_rows,                      // valid rows of this shard
                     long long n_granules,                  // ceil(n_rows / 64)
                     int n_full, const int4* __restrict__ tail_desc, int k, float clamp, uint32_t id_base,
                     unsigned long long* __restrict__ out_keys,  // [slice][qgroups*QW][k]
                     long long q_stride_total,                   // qgroups*QW
                     float* __restrict__ warm_scratch,           // [grid][QW][WARPS][WARM_ROWS] raw scores of the warm-up rows
                     int soft_thr,                               // staged records at which a warp tries to merge
                     unsigned long long* __restrict__ stats,     // STATS builds only (B200NN_SCAN_STATS): counters, see scan_launch
                     int* __restrict__ err_flag) {
    using C = ScanCfg<G, WARPS_, VAR>;
    constexpr int QW = C::QW;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int h = lane / QW, ql = lane - h * QW;
    const bool last_group = (h == G - 1);
    unsigned st_ev = 0, st_cand = 0, st_try = 0, st_got = 0, st_hard = 0, st_candB = 0, st_peel = 0;
    unsigned long long t_0 = 0, t_a = 0, t_b1 = 0, t_b = 0, t_b2 = 0;
    auto now = [&]() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };

    // ---- carve shared memory: LUT at a 64 KB aligned window address, the rest around it ----
    const uint32_t base = smem_u32(dyn_smem);
    const uint32_t lut = (base + 0xFFFFu) & ~0xFFFFu;
    uint32_t low_lo = base, low_hi = lut, high_lo = lut + C::LUT_BYTES, high_hi = base + C::SMEM_BYTES;
    bool ok = true;
    auto carve = [&](uint32_t bytes) -> uint32_t {
        uint32_t p;
        if (low_lo + bytes <= low_hi) { p = low_lo; low_lo += bytes; }
        else { p = high_lo; high_lo += bytes; if (high_lo > high_hi) ok = false; }
        return p;
    };
    uint32_t ring_w = 0, staging_w = 0;
    for (int i = 0; i < C::WARPS; i++) { const uint32_t p = carve(C::RING_BYTES); if (i == w) ring_w = p; }
    const uint32_t lists = carve(C::LIST_BYTES);
    for (int i = 0; i < C::WARPS; i++) { const uint32_t p = carve(C::STAGING_BYTES); if (i == w) staging_w = p; }
    const uint32_t misc = carve(2048);
    if (!ok) { if (threadIdx.x == 0) atomicExch(err_flag, 1); return; }
    // misc: [0,8) lut barrier | [64, 64+WARPS*32) per-warp full barriers | [1024, +QW*8) tau keys | [1536, +QW*4) locks
    const uint32_t lut_bar = misc;
    const uint32_t full_bar = misc + 64 + (uint32_t)w * 32;
    unsigned char* generic_base = dyn_smem - base;  // generic pointer of shared-window address 0
    volatile unsigned long long* tau_key = (volatile unsigned long long*)(generic_base + misc + 1024);
    int* locks = (int*)(generic_base + misc + 1536);
    unsigned int* warm_valid = (unsigned int*)(generic_base + misc + 1700);   // [WARPS] parked rows per warp
    unsigned int* warm_idbase = (unsigned int*)(generic_base + misc + 1800);  // [WARPS] id of a warp's first row

    // ---- barriers (once per CTA; their phases run on across segments) ----
    if (threadIdx.x == 0) mbar_init(lut_bar, 1);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < C::RING_STAGES; s++) mbar_init(full_bar + 8 * s, 1);
    }
    mbar_fence_init();
    uint32_t phases = 0;  // bit s = parity to wait for on ring slot s
    const bool whole = (int)blockIdx.x < n_full;

#pragma unroll 1
    for (int seg = 0; seg < (whole ? 1 : 2); seg++) {
    long long qg, g_lo, g_hi;
    int slice;
    if (whole) { qg = blockIdx.x; slice = 0; g_lo = 0; g_hi = n_granules; }
    else {
        const int4 d = __ldg(tail_desc + ((int)blockIdx.x - n_full) * 2 + seg);
        if (seg > 0 && d.z >= d.w) break;  // the piece lies inside one group (CTA-uniform)
        qg = d.x; slice = d.y; g_lo = d.z; g_hi = d.w;
    }
    if (STATS) { t_0 = now(); st_ev = st_cand = st_try = st_got = st_hard = st_candB = st_peel = 0; }

    // ---- this warp's stream of rows ----
    const long long stages_total = (g_hi - g_lo) * (64 / C::STAGE_ROWS);
    const long long st_per_warp = (stages_total + C::WARPS - 1) / C::WARPS;
    const long long st_lo = min(stages_total, (long long)w * st_per_warp), st_hi = min(stages_total, st_lo + st_per_warp);
    const int n_st = (int)(st_hi - st_lo);
    const long long row0 = g_lo * 64 + st_lo * C::STAGE_ROWS;  // first row of the stream
    const int nblocks = n_st * C::STAGE_BLOCKS;

    // ---- lists ----
    for (int i = threadIdx.x; i < QW * KP; i += blockDim.x) sts64(lists + (uint32_t)i * 8u, KEY_MAX);
    if (threadIdx.x < QW) { tau_key[threadIdx.x] = KEY_MAX; locks[threadIdx.x] = 0; }
    __syncthreads();

    // ---- LUT: 8 TMA bulk copies of 16 KB ----
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(lut_bar, C::LUT_BYTES);
        const float* src = lut_scan + qg * 32768;
#pragma unroll
        for (int i = 0; i < 8; i++) tma_load_1d(lut + i * 16384, src + i * 4096, 16384, lut_bar);
    }

    // ---- code prefetch ----
    // stage s of this warp = rows [row0 + s*STAGE_ROWS, +STAGE_ROWS); plane h of it is STAGE_ROWS
    // consecutive words inside granule (row/64), plane h.
    auto issue_stage = [&](int s, int slot) {
        const uint32_t bar = full_bar + 8 * slot;
        if (lane == 0) mbar_arrive_expect_tx(bar, G * C::STAGE_ROWS * 4);
        __syncwarp();
        if (lane < G) {
            const long long row = row0 + (long long)s * C::STAGE_ROWS;
            const long long gran = row >> 6;
            const int within = (int)(row & 63);
            const uint32_t* src = codesT + (gran * G + lane) * 64 + within;
            const uint32_t dst = ring_w + (uint32_t)lane * C::PLANE_RING_BYTES + (uint32_t)slot * (C::STAGE_ROWS * 4);
            tma_load_1d(dst, src, C::STAGE_ROWS * 4, bar);
        }
    };
    if (n_st > 0) issue_stage(0, 0);

    mbar_wait(lut_bar, (uint32_t)(seg & 1));  // LUT resident

    const uint32_t basereg = lut + (uint32_t)lane * 4u;
    const uint32_t plane = ring_w + (uint32_t)h * C::PLANE_RING_BYTES;
    const uint32_t my_staging = staging_w + (uint32_t)ql * (C::SB * 8);
    const float keep = (h == 0) ? 0.0f : 1.0f;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    // Per-lane threshold state (meaningful on the last lane group, which owns final scores):
    //   tkey : the query's k-th best record so far (KEY_MAX while the list is not full)
    //   tsc  : its score (+inf while not full)
    //   tau  : prefilter on the RAW score: tsc, or +inf while the clamp value itself would still
    //          qualify (clamped scores tie at `clamp`; the record comparison then decides);
    //          -inf on lanes that own no final score.
    unsigned long long tkey = KEY_MAX;
    float tsc = __int_as_float(0x7f800000);
    float tau = last_group ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
    int cnt = 0;  // staged records (last group lanes)
    auto refresh_tau = [&]() {
        if (last_group) {
            tkey = tau_key[ql];
            tsc = tau_f32_of(tkey);
            tau = (clamp <= tsc) ? __int_as_float(0x7f800000) : tsc;
        }
    };
    // byte offset of this lane's current block inside its plane ring (block index = Bg - h)
    uint32_t coff = (h == 0) ? 0u : (uint32_t)(C::PLANE_RING_BYTES - h * 16);
    // 32-bit row bookkeeping for the candidate path: valid relative rows are [0, nrel)
    const uint32_t nrel = (uint32_t)max(0LL, min((long long)nblocks * 4, n_rows - row0));
    const uint32_t idbase = id_base + (uint32_t)row0;
    // number of leading blocks of this warp whose scores are parked instead of filtered
    const int warm_nblocks = max(0, min(C::WARM_BLOCKS, nblocks) - (G - 1));

    // flush the staging buffers of the lanes in `need` (lane mask) into the CTA lists
    auto flush_lanes = [&](unsigned need, bool blocking) {
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int qsel = src - (G - 1) * QW;
            const int nb = __shfl_sync(0xffffffffu, cnt, src);
            int got = 1;
            if (!blocking) {  // try-lock: keep scanning if another warp is merging this query
                if (lane == 0) got = (atomicCAS(locks + qsel, 0, 1) == 0);
                got = __shfl_sync(0xffffffffu, got, 0);
                if (got && lane == 0) __threadfence_block();
                if (STATS) { st_try++; st_got += got; }
            } else if (STATS) st_hard++;
            if (got) {
                warp_flush(lists + (uint32_t)qsel * (KP * 8), locks + qsel, tau_key + qsel,
                           staging_w + (uint32_t)qsel * (C::SB * 8), nb, k, /*lock_held=*/!blocking);
                if (lane == src) { cnt = 0; refresh_tau(); }
            }
        }
    };

    auto rare_path = [&](int Bg, float mn, float o0, float o1, float o2, float o3) {  // the block's four final scores
        // Bg = block counter of lane group 0; this lane's block is Bg - h.  Only lanes of the last
        // lane group can satisfy mn <= tau (tau = -inf elsewhere).
        if (mn <= tau && Bg - h >= warm_nblocks) {  // blocks below warm_nblocks were parked and selected in phase B
            // usually exactly one of the 4 rows qualifies: peel minima until none is left under the threshold
            const uint32_t rel0 = (uint32_t)(Bg - h) * 4u;
            float a0 = o0, a1 = o1, a2 = o2, a3 = o3, m = mn;
            do {
                const int j = (a0 == m) ? 0 : (a1 == m) ? 1 : (a2 == m) ? 2 : 3;
                const float s = clamp < m ? clamp : m;  // std::min(score, threhold-initialised slot), IVFOPQ.cpp:410
                const uint32_t rel = rel0 + (uint32_t)j;
                if (s <= tsc && rel < nrel) {
                    // scores are sums of squares (>= +0): the orderable form is just the sign bit set
                    const uint32_t ord = __float_as_uint(s) | 0x80000000u, id = idbase + rel;
                    if (s < tsc || make_key(ord, id) < tkey) {
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(my_staging + (uint32_t)cnt * 8u), "r"(id), "r"(ord) : "memory");
                        cnt++;
                        if (STATS) st_cand++;
                    }
                }
                if (STATS) st_peel++;
                const float inf = __int_as_float(0x7f800000);
                a0 = (j == 0) ? inf : a0; a1 = (j == 1) ? inf : a1; a2 = (j == 2) ? inf : a2; a3 = (j == 3) ? inf : a3;
                m = fminf(fminf(a0, a1), fminf(a2, a3));
            } while (m <= tau && m < __int_as_float(0x7f800000));
        }
        const unsigned soft = __ballot_sync(0xffffffffu, cnt >= soft_thr);
        if (soft) {
            const unsigned hard = __ballot_sync(0xffffffffu, cnt > C::SB - 4);
            if (hard) flush_lanes(hard, true);
            else flush_lanes(soft, false);
        }
    };

    float* my_scratch = warm_scratch + (((size_t)blockIdx.x * QW + ql) * C::WARPS + w) * C::WARM_ROWS;

    auto block_body = [&](int Bg, auto warm_tag) {
        const uint4 cw = lds128(plane + coff);
        coff += 16u;
        if (coff == (uint32_t)C::PLANE_RING_BYTES) coff = 0u;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        if (G > 1) {
            r0 = __shfl_up_sync(0xffffffffu, o0, QW);
            r1 = __shfl_up_sync(0xffffffffu, o1, QW);
            r2 = __shfl_up_sync(0xffffffffu, o2, QW);
            r3 = __shfl_up_sync(0xffffffffu, o3, QW);
        }
        float s0, s1, s2, s3;
        B2_LOOKUP4(s0, r0, cw.x)
        B2_LOOKUP4(s1, r1, cw.y)
        B2_LOOKUP4(s2, r2, cw.z)
        B2_LOOKUP4(s3, r3, cw.w)
        o0 = s0; o1 = s1; o2 = s2; o3 = s3;
        if constexpr (decltype(warm_tag)::value) {
            // warm-up rows: park the raw scores (one 16-byte store), no filtering
            const int Bh = Bg - (G - 1);
            if (last_group && (unsigned)Bh < (unsigned)warm_nblocks)
                *reinterpret_cast<float4*>(my_scratch + Bh * 4) = make_float4(s0, s1, s2, s3);
        } else if constexpr (!(VAR & 2)) {
            const float mn = fminf(fminf(s0, s1), fminf(s2, s3));
            if (__any_sync(0xffffffffu, mn <= tau)) {
                if (STATS) st_ev++;
                rare_path(Bg, mn, s0, s1, s2, s3);
            }
        }
    };
    // Grouped check (VAR bit 1): four blocks are computed back to back and their 16 final scores stay in registers; ONE
    // vote + branch per group decides whether any of them meets the threshold.  The hot loop is then 4 straight-line
    // block bodies (the loads of a block overlap the add chains of the one before, nothing to serialise on) and a single
    // out-of-line copy of the candidate path, which walks the group's blocks in order.
    float ga0, ga1, ga2, ga3, gb0, gb1, gb2, gb3, gc0, gc1, gc2, gc3;  // scores of a group's first three blocks (the fourth: o0..o3)
    float gma, gmb, gmc, gmd;                                          // the four blocks' minima
    auto group_compute = [&](int Bg0) -> float {
        block_body(Bg0, std::false_type{}); ga0 = o0; ga1 = o1; ga2 = o2; ga3 = o3;
        block_body(Bg0 + 1, std::false_type{}); gb0 = o0; gb1 = o1; gb2 = o2; gb3 = o3;
        block_body(Bg0 + 2, std::false_type{}); gc0 = o0; gc1 = o1; gc2 = o2; gc3 = o3;
        block_body(Bg0 + 3, std::false_type{});
        gma = fminf(fminf(ga0, ga1), fminf(ga2, ga3));
        gmb = fminf(fminf(gb0, gb1), fminf(gb2, gb3));
        gmc = fminf(fminf(gc0, gc1), fminf(gc2, gc3));
        gmd = fminf(fminf(o0, o1), fminf(o2, o3));
        return fminf(fminf(gma, gmb), fminf(gmc, gmd));
    };
    auto group_rare = [&](int Bg0) {  // visits, in order, the blocks of the group that hold a candidate (one copy of the candidate path)
        // which blocks: four independent votes up front (a flush on the way can only tighten tau; rare_path re-tests per lane)
        unsigned hits = (__any_sync(0xffffffffu, gma <= tau) ? 1u : 0u) | (__any_sync(0xffffffffu, gmb <= tau) ? 2u : 0u) |
                        (__any_sync(0xffffffffu, gmc <= tau) ? 4u : 0u) | (__any_sync(0xffffffffu, gmd <= tau) ? 8u : 0u);
#pragma unroll 1
        while (hits) {
            const int u = __ffs(hits) - 1;
            hits &= hits - 1;
            if (STATS) st_ev++;
            const float s0 = (u == 0) ? ga0 : (u == 1) ? gb0 : (u == 2) ? gc0 : o0;
            const float s1 = (u == 0) ? ga1 : (u == 1) ? gb1 : (u == 2) ? gc1 : o1;
            const float s2 = (u == 0) ? ga2 : (u == 1) ? gb2 : (u == 2) ? gc2 : o2;
            const float s3 = (u == 0) ? ga3 : (u == 1) ? gb3 : (u == 2) ? gc3 : o3;
            rare_path(Bg0 + u, fminf(fminf(s0, s1), fminf(s2, s3)), s0, s1, s2, s3);
        }
    };
    int Bg = 0;
    // a run of blocks (a multiple of 4)
    auto run_blocks = [&](int nb, auto warm_tag) {
        if constexpr ((VAR & 2) && !decltype(warm_tag)::value) {
            // the inner loop is the hot path: straight-line group, vote, a forward branch that is normally NOT taken
            int bb = 0;
            while (bb < nb) {
                bool hit;
#pragma unroll 1
                do {
                    const float mn = group_compute(Bg);
                    bb += 4; Bg += 4;
                    hit = __any_sync(0xffffffffu, mn <= tau);
                } while (!hit && bb < nb);
                if (hit) group_rare(Bg - 4);
            }
        } else {
#pragma unroll 4
            for (int bb = 0; bb < nb; bb++, Bg++) block_body(Bg, warm_tag);
        }
    };

    int slot = 0, st = 0;
    auto next_stage = [&]() {
        const int next = (slot == C::RING_STAGES - 1) ? 0 : slot + 1;
        // the slot after the current one held stage st-2, which every lane group has left
        if (st + 1 < n_st) issue_stage(st + 1, next);
        mbar_wait(full_bar + 8 * slot, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        slot = next;
    };
    // one stage of blocks; VAR = 1: two-slot ring, the next stage is requested in the middle of this one
    auto run_stage = [&](auto warm_tag, bool refresh) {
        if constexpr (!(VAR & 1)) {
            next_stage();
            if (refresh) refresh_tau();
            run_blocks(C::STAGE_BLOCKS, warm_tag);
        } else {
            constexpr int HALF = C::STAGE_BLOCKS / 2;
            static_assert(HALF >= G - 1, "the lagging lane groups must have left the previous stage at the half-way point");
            mbar_wait(full_bar + 8 * slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            if (refresh) refresh_tau();
#pragma unroll 1
            for (int half = 0; half < 2; half++) {  // one copy of the unrolled body (instruction cache)
                if (half == 1 && st + 1 < n_st) issue_stage(st + 1, slot ^ 1);
                run_blocks(HALF, warm_tag);
            }
            slot ^= 1;
        }
    };
    // ---- phase A: warm-up stages, scores parked ----
    for (; st < min(n_st, C::WARM_STAGES); st++) run_stage(std::true_type{}, false);
    // ---- phase B: the CTA selects the top-k of all parked rows in one dense pass ----
    if (lane == 0) {
        warm_valid[w] = min((uint32_t)warm_nblocks * 4u, nrel);
        warm_idbase[w] = idbase;
    }
    __threadfence();  // parked scores visible to the selecting warps
    if (STATS) t_a = now();
    asm volatile("bar.sync 1, %0;" ::"r"(C::WARPS * 32) : "memory");
    if (STATS) t_b1 = now();
    // One warp per query holds the query's parked scores in registers (TOT/32 per lane) and finds the exact
    // k-th smallest record by bisection on the score bits (scores are >= +0, so their bit patterns order like
    // the values) and, only if the k-th score is tied, on the ids; the k survivors are rank-sorted straight
    // into the (still empty) list.  ~30 rounds of register compares instead of ~15 serialised list merges.
    {
        static_assert(C::WARM_ROWS == 128, "slot -> (warp, row) mapping below assumes 128 parked rows per warp");
        constexpr int TOT = C::WARPS * C::WARM_ROWS;  // parked slots per query
        constexpr int E = TOT / 32;                   // slots per lane: slot p = i*32 + lane -> warp i/4, row (i%4)*32 + lane
        for (int qq = w; qq < QW; qq += C::WARPS) {
            const float* qs = warm_scratch + ((size_t)blockIdx.x * QW + qq) * TOT;
            const uint32_t L = lists + (uint32_t)qq * (KP * 8);
            constexpr uint32_t NONE = 0xFFFFFFFFu;  // above every score pattern
            uint32_t x[E];
            uint32_t n_valid = 0, vmin = NONE, vmax = 0;
#pragma unroll
            for (int i = 0; i < E; i++) {
                const uint32_t off = (uint32_t)((i & 3) * 32 + lane);
                const bool valid = off < warm_valid[i >> 2];
                float sc = valid ? __ldcg(qs + i * 32 + lane) : 0.0f;
                sc = clamp < sc ? clamp : sc;
                x[i] = valid ? __float_as_uint(sc) : NONE;
                n_valid += valid ? 1u : 0u;
                if (valid) { vmin = min(vmin, x[i]); vmax = max(vmax, x[i]); }
            }
            n_valid = __reduce_add_sync(0xffffffffu, n_valid);
            vmin = __reduce_min_sync(0xffffffffu, vmin);
            vmax = __reduce_max_sync(0xffffffffu, vmax);
            const uint32_t kk = min((uint32_t)k, n_valid);
            if (kk == 0) continue;  // warp-uniform
            // smallest v with #(x <= v) >= kk
            uint32_t lo = vmin, hi = vmax;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                // four independent predicated counters: one dependent add per 4 elements instead of a 2-op chain per element
                uint32_t c4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < E; i++)
                    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c4[i & 3]) : "r"(x[i]), "r"(mid));
                const uint32_t c = __reduce_add_sync(0xffffffffu, (c4[0] + c4[1]) + (c4[2] + c4[3]));
                if (c >= kk) hi = mid; else lo = mid + 1;
            }
            const uint32_t v = lo;
            uint32_t c_lt = 0, c_le = 0;
#pragma unroll
            for (int i = 0; i < E; i++) { c_lt += (x[i] < v) ? 1u : 0u; c_le += (x[i] <= v) ? 1u : 0u; }
            c_lt = __reduce_add_sync(0xffffffffu, c_lt);
            c_le = __reduce_add_sync(0xffffffffu, c_le);
            // Tied k-th score (typically the clamp value): the smallest ids win.  The ids of the parked rows ascend with
            // the slot order p = i*32 + lane (warps own ascending row ranges), so the tie rule is a running count in the
            // compaction pass itself -- no second bisection.
            const bool tied = c_le > kk;   // warp-uniform
            const uint32_t need = kk - c_lt;  // ties to keep
            // compact the kk survivors into this warp's staging area
            uint32_t ns = 0;
            const uint32_t lt_mask = (1u << lane) - 1u;
            if (!tied) {
#pragma unroll
                for (int i = 0; i < E; i++) {
                    const bool pass = x[i] <= v;
                    const unsigned msk = __ballot_sync(0xffffffffu, pass);
                    if (pass)
                        sts64(staging_w + (ns + __popc(msk & lt_mask)) * 8u,
                              make_key(x[i] | 0x80000000u, warm_idbase[i >> 2] + (uint32_t)((i & 3) * 32 + lane)));
                    ns += __popc(msk);
                }
            } else {
                uint32_t ties_seen = 0;
#pragma unroll
                for (int i = 0; i < E; i++) {
                    const bool eq = x[i] == v;
                    const unsigned meq = __ballot_sync(0xffffffffu, eq);
                    const bool pass = x[i] < v || (eq && ties_seen + __popc(meq & lt_mask) < need);
                    ties_seen += __popc(meq);
                    const unsigned msk = __ballot_sync(0xffffffffu, pass);
                    if (pass)
                        sts64(staging_w + (ns + __popc(msk & lt_mask)) * 8u,
                              make_key(x[i] | 0x80000000u, warm_idbase[i >> 2] + (uint32_t)((i & 3) * 32 + lane)));
                    ns += __popc(msk);
                }
            }
            if (STATS) st_candB += ns;
            __syncwarp();
            // rank sort into the list (keys are distinct)
            unsigned long long c[KP / 32];
            int rk[KP / 32];
#pragma unroll
            for (int t = 0; t < KP / 32; t++) {
                c[t] = (uint32_t)(lane + 32 * t) < ns ? lds64(staging_w + (uint32_t)(lane + 32 * t) * 8u) : KEY_MAX;
                rk[t] = 0;
            }
            for (uint32_t j = 0; j < ns; j++) {
                const unsigned long long kj = lds64(staging_w + j * 8u);
#pragma unroll
                for (int t = 0; t < KP / 32; t++) rk[t] += (kj < c[t]) ? 1 : 0;
            }
#pragma unroll
            for (int t = 0; t < KP / 32; t++)
                if ((uint32_t)(lane + 32 * t) < ns) sts64(L + (uint32_t)rk[t] * 8u, c[t]);
            __syncwarp();
            if (lane == 0) tau_key[qq] = (ns >= (uint32_t)k) ? lds64(L + (uint32_t)(k - 1) * 8u) : KEY_MAX;
        }
    }
    if (STATS) t_b = now();
    asm volatile("bar.sync 1, %0;" ::"r"(C::WARPS * 32) : "memory");
    if (STATS) t_b2 = now();
    // ---- phase C: stream the rest against the thresholds ----
    for (; st < n_st; st++) run_stage(std::false_type{}, true);
    refresh_tau();
    // drain the lane-group pipeline (grouped: whole groups; the surplus blocks lie past the stream and are rejected by row)
    if constexpr (VAR & 2) run_blocks((G - 1 + 3) & ~3, std::false_type{});
    else for (int d = 0; d < G - 1; d++, Bg++) block_body(Bg, std::false_type{});

    // ---- flush what is still staged, then emit the CTA's sorted lists ----
    flush_lanes(__ballot_sync(0xffffffffu, last_group && cnt > 0), true);
    __syncthreads();
    for (int i = threadIdx.x; i < QW * k; i += blockDim.x) {
        const int qq = i / k, j = i - qq * k;
        out_keys[((long long)slice * q_stride_total + qg * QW + qq) * k + j] = lds64(lists + (uint32_t)(qq * KP + j) * 8u);
    }
    __syncthreads();  // lists, LUT and scratch are free for the next segment
    if (STATS) {
        const unsigned long long t_e = now();
        auto wsum = [&](unsigned v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
        const unsigned c_cand = wsum(st_cand), c_peel = wsum(st_peel);
        if (lane == 0) {
            atomicAdd(stats + 0, (unsigned long long)st_ev);
            atomicAdd(stats + 1, (unsigned long long)c_cand);
            atomicAdd(stats + 2, (unsigned long long)st_try);
            atomicAdd(stats + 3, (unsigned long long)st_got);
            atomicAdd(stats + 4, (unsigned long long)st_hard);
            atomicAdd(stats + 5, (unsigned long long)st_candB);
            atomicAdd(stats + 6, (unsigned long long)c_peel);
            if (w == 0) {
                atomicAdd(stats + 8, t_a - t_0);    // phase A (warp 0)
                atomicAdd(stats + 9, t_b1 - t_a);   // wait at barrier 1
                atomicAdd(stats + 10, t_b - t_b1);  // phase B
                atomicAdd(stats + 11, t_b2 - t_b);  // wait at barrier 2
                atomicAdd(stats + 12, t_e - t_b2);  // phase C + emit
                atomicAdd(stats + 13, 1ull);
            }
        }
    }
    }  // segment
}

// =============================================================================================
// Generic IVF scan (true list probing, any K): one CTA per (query, probe).  The LUT for the
// probed list lives in shared memory, lanes take rows of the list.  Two outputs:
//   group mode: match[q][group] = min(score, match)   (IVFOPQ::QueryThrehold, IVFOPQ.cpp:403-411)
//   row mode  : scores[q][row]  = min(score, clamp-initialised slot)   (videoId = row)
// Scores are non-negative, so an integer atomicMin on the float bits is an exact float min.
// =============================================================================================
__global__ void ivf_scan_kernel(const float* __restrict__ lut,            // [nq*nprobe][M*ksub]
                                const int* __restrict__ probes,           // [nq*nprobe]
                                const long long* __restrict__ list_off,   // [K+1] CSR offsets (sorted order)
                                const unsigned char* __restrict__ codes_sorted,  // [n][M]
                                const int* __restrict__ slot_sorted,      // [n] group id or original row
                                int M, int ksub, int nprobe, long long out_stride, float* __restrict__ out) {
    extern __shared__ float s_lut[];
    const long long qp = blockIdx.x;
    const long long qi = qp / nprobe;
    const int vw = probes[qp];
    for (int e = threadIdx.x; e < M * ksub; e += blockDim.x) s_lut[e] = lut[qp * M * ksub + e];
    __syncthreads();
    const long long lo = list_off[vw], hi = list_off[vw + 1];
    int* o = (int*)(out + qi * out_stride);
    for (long long r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        const unsigned char* c = codes_sorted + r * M;
        float score = 0.0f;
        for (int m = 0; m < M; m++) score = __fadd_rn(score, s_lut[m * ksub + c[m]]);
        atomicMin(o + slot_sorted[r], __float_as_int(score));
    }
}

// =============================================================================================
// f-1: true IVF search with per-row ids (videoId = row), fused: one CTA per query walks its nprobe
// lists; for each it builds the residual LUT in shared memory (IVFOPQ.cpp:376-398), scans the list
// (IVFOPQ.cpp:403-411) and feeds the exact top-k lists -- no dense [queries x rows] score matrix.
// Reference semantics kept: every row starts at `clamp` (matchScore initialised to threhold,
// IVFOPQ.cpp:369) and get_sort_results breaks ties by id, so when fewer than k probed rows score
// below the clamp the tail is the smallest row ids at exactly `clamp` (probed or not).
// =============================================================================================
// CTA-wide selection buffer shared by the list/matrix selection kernels: keys are appended to s_list[0, *s_cnt) (warp-
// aggregated counter) and this sorts them ascending with a bitonic network (padded with KEY_MAX to a power of two >= 256),
// keeps the best k and publishes the k-th best as the new threshold.  Must be called by every thread of the CTA.
__device__ __forceinline__ void cta_sort_trim(unsigned long long* s_list, int* s_cnt, unsigned long long* s_tau, int k) {
    __syncthreads();
    const int have = *s_cnt;
    int n2 = 256;
    while (n2 < have) n2 <<= 1;
    for (int i = have + threadIdx.x; i < n2; i += blockDim.x) s_list[i] = KEY_MAX;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (stride - 1)), j = i + stride;  // i has bit `stride` clear
                const unsigned long long a = s_list[i], b = s_list[j];
                const bool up = (i & size) == 0;
                if ((a > b) == up) { s_list[i] = b; s_list[j] = a; }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        const int keep = min(have, k);
        *s_cnt = keep;
        *s_tau = keep >= k ? s_list[k - 1] : KEY_MAX;
    }
    __syncthreads();
}

constexpr int IVF_PG = 4;  // probes whose LUTs are built together (shared memory: IVF_PG x M x ksub floats)
template <int DS>
__global__ void __launch_bounds__(256)
ivf_search_topk_kernel(const float* __restrict__ q_rot, long long nq, int D, const int* __restrict__ probes, int nprobe,
                       const float* __restrict__ coarse, const float* __restrict__ cb, int M, int ksub,
                       const long long* __restrict__ list_off, const unsigned char* __restrict__ codes_sorted,
                       const int* __restrict__ row_sorted, long long n_rows, int k, float clamp, uint32_t id_base, int n_pg /* <= IVF_PG */,
                       unsigned long long* __restrict__ out_keys) {
    // Selection: the probed lists are short (1 M rows over 8192 lists: ~120 rows each), so a query sees a few hundred
    // candidates in all.  They are appended to ONE buffer per CTA (warp-aggregated slot counter) and the buffer is sorted
    // by a bitonic network when it is about to overflow and once at the end -- no per-warp staging, no lock, no
    // serialised list merges (those were ~12 lock-ordered merges per query, the bulk of the kernel's time).
    constexpr int BUF = 1024;
    extern __shared__ __align__(16) unsigned char dsm[];
    float* s_lut = reinterpret_cast<float*>(dsm);                 // [n_pg][M*ksub]
    float* s_res = s_lut + n_pg * M * ksub;                       // [n_pg][D]
    __shared__ __align__(16) unsigned long long s_list[BUF];      // [0, s_cnt): the best so far (first k sorted after a trim) + new candidates
    __shared__ unsigned long long s_tau;                           // k-th best record after the last trim (KEY_MAX: fewer than k yet)
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long q = blockIdx.x;
    if (threadIdx.x == 0) { s_tau = KEY_MAX; s_cnt = 0; }
    auto sort_and_trim = [&]() { cta_sort_trim(s_list, &s_cnt, &s_tau, k); };
    const uint32_t clamp_ord = f32_orderable(clamp);
    // The LUTs of up to IVF_PG probes are built in ONE pass over the codebook (a codeword row is loaded once, with 16-byte
    // loads, and meets every probe's residual): a third of the L2 traffic and of the CTA barriers of a LUT per probe.
    const int LUTN = M * ksub;
    for (int p0 = 0; p0 < nprobe; p0 += n_pg) {
        const int pg = min(n_pg, nprobe - p0);
        __syncthreads();  // previous LUTs fully consumed; lists initialised
        for (int i = threadIdx.x; i < pg * D; i += blockDim.x) {
            const int pp = i / D, j = i - pp * D;
            const int vw = probes[q * nprobe + p0 + pp];
            s_res[pp * D + j] = __fsub_rn(q_rot[q * D + j], coarse[(long long)vw * D + j]);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < LUTN; e += blockDim.x) {
            const int m = e / ksub;
            const float* c = cb + (long long)e * DS;
            float cv[DS];
            if (DS % 4 == 0) {
#pragma unroll
                for (int t = 0; t < DS; t += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(c + t));
                    cv[t] = v.x; cv[t + 1] = v.y; cv[t + 2] = v.z; cv[t + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int t = 0; t < DS; t++) cv[t] = __ldg(c + t);
            }
#pragma unroll
            for (int pp = 0; pp < IVF_PG; pp++) {
                if (pp < pg) {
                    const float* r = s_res + pp * D + m * DS;
                    float acc = 0.0f;
#pragma unroll
                    for (int t = 0; t < DS; t++) {
                        const float d = __fsub_rn(r[t], cv[t]);
                        acc = __fadd_rn(acc, __fmul_rn(d, d));
                    }
                    s_lut[pp * LUTN + e] = acc;
                }
            }
        }
        __syncthreads();
        for (int pp = 0; pp < pg; pp++) {
            const int vw = probes[q * nprobe + p0 + pp];
            const float* lut = s_lut + pp * LUTN;
            const long long lo = list_off[vw], hi = list_off[vw + 1];
            for (long long r0 = lo; r0 < hi; r0 += blockDim.x) {  // CTA-uniform passes of 256 rows
                // same decision in every thread: read the counter, barrier, only then may anyone append (see dense_topk_kernel)
                const int cnt0 = s_cnt;
                __syncthreads();
                if (cnt0 > BUF - (int)blockDim.x) sort_and_trim();
                const unsigned long long tau = s_tau;
                const long long r = r0 + threadIdx.x;
                bool pass = false;
                unsigned long long key = 0;
                if (r < hi) {
                    const unsigned char* c = codes_sorted + r * M;
                    float score = 0.0f;
                    if ((M & 15) == 0) {  // 16 code bytes per load
                        for (int m0 = 0; m0 < M; m0 += 16) {
                            const uint4 cw = __ldg(reinterpret_cast<const uint4*>(c + m0));
                            const uint32_t wd[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
                            for (int u = 0; u < 16; u++)
                                score = __fadd_rn(score, lut[(m0 + u) * ksub + ((wd[u >> 2] >> (8 * (u & 3))) & 0xFFu)]);
                        }
                    } else {
                        for (int m = 0; m < M; m++) score = __fadd_rn(score, lut[m * ksub + c[m]]);
                    }
                    if (score < clamp) {  // rows at or above the clamp are indistinguishable from unprobed rows
                        key = make_key(f32_orderable(score), id_base + (uint32_t)row_sorted[r]);
                        pass = key < tau;
                    }
                }
                const unsigned msk = __ballot_sync(0xffffffffu, pass);
                if (msk) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt, __popc(msk));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (pass) s_list[base + __popc(msk & ((1u << lane) - 1))] = key;
                }
                __syncthreads();  // s_cnt settled before the next pass reads it
            }
        }
    }
    sort_and_trim();
    for (int i = s_cnt + threadIdx.x; i < k; i += blockDim.x) s_list[i] = KEY_MAX;  // unfilled slots (the sort's padding may end below k)
    __syncthreads();
    // Fewer than k probed rows scored below the clamp: the tail is (clamp, smallest ids not already in the list), IVFOPQ.cpp:369 +
    // common.h:25-37.  The list holds real < k records, so among the ids [0, 2k) at least k are absent: thread t < 2k tests id t
    // against the list, a ballot/prefix count gives every absent id its slot.
    __shared__ int s_real, s_wcnt[8];
    if (w == 0) {
        int real = 0;
        for (int j = lane; j < k; j += 32) real += (s_list[j] != KEY_MAX) ? 1 : 0;
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) real += __shfl_xor_sync(0xffffffffu, real, sft);
        if (lane == 0) s_real = real;
    }
    __syncthreads();
    const int real = s_real;
    if (real < k) {  // CTA-uniform
        const int t = threadIdx.x;  // 256 threads >= 2k
        bool absent = t < 2 * k && (long long)t < n_rows;
        if (absent)
            for (int j = 0; j < real; j++) absent &= ((uint32_t)(s_list[j] & 0xFFFFFFFFull) != id_base + (uint32_t)t);
        const unsigned msk = __ballot_sync(0xffffffffu, absent);
        if (lane == 0) s_wcnt[w] = __popc(msk);
        __syncthreads();
        int before = __popc(msk & ((1u << lane) - 1));
        for (int i = 0; i < w; i++) before += s_wcnt[i];
        if (absent && real + before < k) s_list[real + before] = make_key(clamp_ord, id_base + (uint32_t)t);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) out_keys[q * k + j] = s_list[j];
}

// =============================================================================================
// f-5: the reference's actual result type -- per-VIDEO scores of a multi-frame query.
//   frame sum : score_total[id] = sum over frames f = 0..F-1 of matchScore[f][id], sequential fp32 from 0.0f
//               (opq/src/multi_frame_index_test.cpp:59-67)
//   selection : get_sort_results(score_total, k) = the k smallest (score, id) pairs, ascending
//               (opq/src/common.h:25-37)
// =============================================================================================
__global__ void frame_sum_kernel(const float* scores, int n_frames, long long ng, float* total) {  // total may alias row 0 of scores
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int f = 0; f < n_frames; f++) acc = __fadd_rn(acc, scores[(long long)f * ng + g]);
        total[g] = acc;
    }
}

// k smallest (value, id) records of each row of a dense fp32 matrix (row stride ld); CTA (row, slice) scans its share of the
// n columns.  id = id_map[column] when a map is given (the flat index passes the RANK of every row's label, so that ties
// resolve on the label as std::pair<dist_t, labeltype> does), else the column index.  out_keys [slice][rows][k].
__global__ void __launch_bounds__(256)
dense_topk_kernel(const float* __restrict__ values, long long n, long long ld, int k, const uint32_t* __restrict__ id_map,
                  unsigned long long* __restrict__ out_keys) {
    // passes of 1024 columns, four coalesced loads per thread, the NEXT pass's loads already in flight while this one is
    // filtered (the matrix comes from DRAM: the kernel is bound by bytes in flight; eight loads per thread with a 32 KB
    // buffer measured slower); a value meets the threshold's SCORE first and only a survivor fetches its id; survivors go
    // to the CTA buffer (cta_sort_trim): one sort after the first pass sets the threshold, after which a pass rarely holds
    // a candidate at all.
    constexpr int PER = 4, PASS = PER * 256, BUF = 2 * PASS;
    __shared__ __align__(16) unsigned long long s_list[BUF];
    __shared__ unsigned long long s_tau;
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31;
    const float* v = values + (long long)blockIdx.x * ld;
    const long long c_lo = (n * blockIdx.y) / gridDim.y, c_hi = (n * (blockIdx.y + 1)) / gridDim.y;
    if (threadIdx.x == 0) { s_tau = KEY_MAX; s_cnt = 0; }
    float x[PER], nx[PER];
    auto fetch = [&](float (&d)[PER], long long c0) {
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const long long col = c0 + j * 256 + threadIdx.x;
            d[j] = col < c_hi ? __ldcs(v + col) : 0.0f;  // read once: streaming
        }
    };
    fetch(x, c_lo);
    __syncthreads();
    for (long long c0 = c_lo; c0 < c_hi; c0 += PASS) {
        if (c0 + PASS < c_hi) fetch(nx, c0 + PASS);
        // The decision to sort must be the same in every thread: all read the counter first, THEN a barrier, and only
        // then may anyone append (a fast warp's appends of this pass would otherwise be seen by a slow warp's test).
        const int cnt0 = s_cnt;
        unsigned long long tau = s_tau;
        __syncthreads();
        if (cnt0 > BUF - PASS || (tau == KEY_MAX && cnt0 >= k)) {
            cta_sort_trim(s_list, &s_cnt, &s_tau, k);
            tau = s_tau;
        }
        const uint32_t tau_ord = (uint32_t)(tau >> 32);
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const long long col = c0 + j * 256 + threadIdx.x;
            bool pass = false;
            unsigned long long key = 0;
            if (col < c_hi) {
                const uint32_t ord = f32_orderable(x[j]);
                if (ord <= tau_ord) {
                    key = make_key(ord, id_map ? __ldg(id_map + col) : (uint32_t)col);
                    pass = key < tau;
                }
            }
            const unsigned msk = __ballot_sync(0xffffffffu, pass);
            if (msk) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_cnt, __popc(msk));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) s_list[base + __popc(msk & ((1u << lane) - 1))] = key;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; j++) x[j] = nx[j];
    }
    cta_sort_trim(s_list, &s_cnt, &s_tau, k);
    unsigned long long* out = out_keys + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * k;
    for (int j = threadIdx.x; j < k; j += blockDim.x) out[j] = s_list[j];  // slots past the columns scanned hold KEY_MAX (the sort's padding)
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// =============================================================================================
// host launchers
// =============================================================================================
static inline unsigned grid_for(long long work, int per_block, int sm_count, int waves = 8) {
    long long g = (work + per_block - 1) / per_block;
    const long long cap = (long long)sm_count * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

int launch_rotate_perm(Ctx* ctx, const float* x, long long n, int D, const int* perm, float* y) {
    if (n == 0) return 0;
    rotate_perm_kernel<<<grid_for(n * D, 256, ctx->sm_count), 256, 0, ctx->stream>>>(x, n, D, perm, y);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_coarse_assign(Ctx* ctx, const float* x, long long n, int D, const float* coarseT, int K, int* out_list) {
    if (n == 0) return 0;
    if (tiled_nearest_pays(D, K)) return launch_tiled_nearest(ctx, x, D, 0, n, D, coarseT, K, 0, 1, out_list, nullptr);
    const int warps = 8;
    coarse_assign_kernel<<<grid_for(n, warps, ctx->sm_count), warps * 32, warps * D * sizeof(float), ctx->stream>>>(
        x, n, D, coarseT, K, out_list);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_pq_encode(Ctx* ctx, const float* x, long long n, int D, const float* coarse, const int* list, const float* cbT,
                     int M, int ksub, unsigned char* codes) {
    if (n == 0) return 0;
    const int ds = D / M;
    if (pq_encode_tile_supported(ds, ksub) && n >= 64 && !getenv("B200NN_NO_ENCODE_TILE"))  // the register-tiled kernel (dist_tile.cu)
        return launch_pq_encode_tile(ctx, x, n, D, coarse, list, cbT, M, ksub, codes);
    const unsigned grid = grid_for(n * M, 8, ctx->sm_count, 16);
#define B2_ENC(DS_)                                                                                              \
    case DS_:                                                                                                    \
        pq_encode_kernel<DS_><<<grid, 256, 0, ctx->stream>>>(x, n, D, coarse, list, cbT, M, ksub, codes);        \
        break;
    switch (ds) {
        B2_ENC(1) B2_ENC(2) B2_ENC(4) B2_ENC(8) B2_ENC(16) B2_ENC(32)
        default: B2_FAIL(-4, "pq_encode: unsupported sub-vector dimension (D/M must be 1,2,4,8,16,32)");
    }
#undef B2_ENC
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_codes_to_scan_layout(Ctx* ctx, const unsigned char* codes, long long n, int M, uint32_t* codesT, long long n_pad) {
    if (n_pad == 0) return 0;
    codes_to_scan_layout_kernel<<<grid_for(n_pad * (M / 4), 256, ctx->sm_count), 256, 0, ctx->stream>>>(codes, n, M, codesT, n_pad);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_coarse_probe(Ctx* ctx, const float* q, long long nq, int D, const float* coarseT, int K, int nk, int* out_lists) {
    if (nq == 0) return 0;
    if (nk > 8 || nk < 1 || nk > K) B2_FAIL(-1, "coarse_probe: nprobe must be in [1, min(8, K)]");
    if (tiled_nearest_pays(D, K)) return launch_tiled_nearest(ctx, q, D, 0, nq, D, coarseT, K, 1, nk, out_lists, nullptr);
    const int warps = 4;
    coarse_probe_kernel<<<grid_for(nq, warps, ctx->sm_count), warps * 32, warps * D * sizeof(float), ctx->stream>>>(
        q, nq, D, coarseT, K, nk, out_lists);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_lut_build_std(Ctx* ctx, const float* q, long long nq, int D, const int* probes, int nprobe, const float* coarse,
                         const float* cb, int M, int ksub, float* lut) {
    if (nq == 0) return 0;
    const int ds = D / M;
    const unsigned grid = (unsigned)(nq * nprobe);
#define B2_LUT(DS_)                                                                                             \
    case DS_:                                                                                                   \
        lut_build_std_kernel<DS_><<<grid, 256, D * sizeof(float), ctx->stream>>>(q, nq, D, probes, nprobe, coarse, cb, \
                                                                                 M, ksub, lut);                  \
        break;
    switch (ds) {
        B2_LUT(1) B2_LUT(2) B2_LUT(4) B2_LUT(8) B2_LUT(16) B2_LUT(32)
        default: B2_FAIL(-4, "lut_build: unsupported sub-vector dimension");
    }
#undef B2_LUT
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

template <int G>
static int lut_scan_dispatch(Ctx* ctx, const float* q, long long nq, int D, const float* centroid, const float* cb,
                             float* lut_scan) {
    constexpr int QW = 32 / G, M = 4 * G;
    const int ds = D / M;
    const unsigned qgroups = (unsigned)((nq + QW - 1) / QW);
    unsigned parts = 1;  // entry slices per query group: aim at >= 4 CTAs per SM
    while (parts < 8 && (unsigned long long)qgroups * parts < 4ull * ctx->sm_count) parts *= 2;
    const dim3 grid(qgroups, parts);
    const size_t smem = (size_t)QW * (D + 4) * sizeof(float);
#define B2_LS(DS_)                                                                                                  \
    case DS_:                                                                                                       \
        if (smem > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(lut_build_scan_kernel<G, DS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        lut_build_scan_kernel<G, DS_><<<grid, 512, smem, ctx->stream>>>(q, nq, D, centroid, cb, lut_scan);           \
        break;
    switch (ds) {
        B2_LS(1) B2_LS(2) B2_LS(4) B2_LS(8) B2_LS(16) B2_LS(32)
        default: B2_FAIL(-4, "lut_build_scan: unsupported sub-vector dimension");
    }
#undef B2_LS
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_lut_build_scan(Ctx* ctx, int M, const float* q, long long nq, int D, const float* centroid, const float* cb,
                          float* lut_scan) {
    if (nq == 0) return 0;
    switch (M) {
        case 4: return lut_scan_dispatch<1>(ctx, q, nq, D, centroid, cb, lut_scan);
        case 8: return lut_scan_dispatch<2>(ctx, q, nq, D, centroid, cb, lut_scan);
        case 16: return lut_scan_dispatch<4>(ctx, q, nq, D, centroid, cb, lut_scan);
        case 32: return lut_scan_dispatch<8>(ctx, q, nq, D, centroid, cb, lut_scan);
        default: B2_FAIL(-4, "fast ADC scan supports M in {4, 8, 16, 32}");
    }
}

int scan_queries_per_cta(int M) { return M >= 4 ? 128 / M : 0; }

// Work plan of the scan (see adc_scan_topk_kernel): n_full query groups get one whole-shard CTA each
// (full waves of one-CTA-per-SM); the (group, granule) work of the partial last wave is cut into
// n_tail equal pieces, one per tail CTA, each piece = one or two segments.
void scan_plan(int sm_count, long long qgroups, long long n_granules, ScanPlan* plan) {
    const long long waves = qgroups / sm_count;
    const long long rem = qgroups - waves * sm_count;
    plan->n_full = (int)(waves * sm_count);
    plan->n_tail = 0;
    plan->slices = 1;
    plan->desc.clear();
    if (rem == 0) return;
    const long long work = rem * n_granules;
    // one piece per SM, but no piece below 16 granules (1024 rows) and never fewer pieces than groups
    long long T = std::min<long long>(sm_count, std::max<long long>(rem, work / 16));
    // when whole slices per group already fill >= 90 % of the SMs keep the pieces aligned with the group
    // boundaries (no two-segment CTAs: one warm-up each; measured 0.6 % faster at cfg3's 68 tail groups)
    const long long s_int = sm_count / rem;
    if (T == sm_count && rem * s_int * 10 >= (long long)sm_count * 9) T = rem * s_int;
    plan->n_tail = (int)T;
    plan->desc.assign((size_t)T * 8, 0);
    std::vector<int> next_slice((size_t)rem, 0);
    for (long long t = 0; t < T; t++) {
        int* d = plan->desc.data() + t * 8;
        if (n_granules == 0 || T == rem) {  // one whole group per CTA (also the empty shard)
            d[0] = (int)(plan->n_full + t); d[1] = 0; d[2] = 0; d[3] = (int)n_granules;
            next_slice[t] = 1;
            continue;
        }
        const long long lo = work * t / T, hi = work * (t + 1) / T;
        const long long g1 = lo / n_granules;
        d[0] = (int)(plan->n_full + g1); d[1] = next_slice[g1]++;
        d[2] = (int)(lo - g1 * n_granules); d[3] = (int)(std::min(hi, (g1 + 1) * n_granules) - g1 * n_granules);
        if (hi > (g1 + 1) * n_granules) {  // the piece runs on into the next group
            const long long g2 = g1 + 1;
            d[4] = (int)(plan->n_full + g2); d[5] = next_slice[g2]++;
            d[6] = 0; d[7] = (int)(hi - g2 * n_granules);
        }
    }
    for (int v : next_slice) plan->slices = std::max(plan->slices, v);
}

template <int G, int WARPS_, int VAR = 0>
static int scan_launch(Ctx* ctx, const uint32_t* codesT, const float* lut_scan, long long n_rows, long long qgroups,
                       int n_full, int n_tail, const int4* tail_desc, int k, float clamp, uint32_t id_base,
                       unsigned long long* out_keys, float* warm_scratch) {
    using C = ScanCfg<G, WARPS_, VAR>;
    // the opt-in is per device (a process may hold contexts on several GPUs): set it on every launch, it is a cheap host call
    B2_CUDA(cudaFuncSetAttribute(adc_scan_topk_kernel<G, WARPS_, false, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const long long n_gran = (n_rows + 63) / 64;
    const unsigned grid = (unsigned)(n_full + n_tail);
    int soft = C::SOFT;
    if (const char* e = getenv("B200NN_SOFT")) soft = std::max(1, std::min(C::SB - 4, atoi(e)));
    if (getenv("B200NN_SCAN_STATS")) {  // development aid: counters of the candidate path, printed per launch
        B2_CUDA(cudaFuncSetAttribute(adc_scan_topk_kernel<G, WARPS_, true, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        unsigned long long* d_stats = nullptr;
        unsigned long long hs[16];
        B2_CUDA(cudaMalloc(&d_stats, sizeof(hs)));
        B2_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(hs), ctx->stream));
        adc_scan_topk_kernel<G, WARPS_, true, VAR><<<grid, C::WARPS * 32, C::SMEM_BYTES, ctx->stream>>>(
            codesT, lut_scan, n_rows, n_gran, n_full, tail_desc, k, clamp, id_base, out_keys, qgroups * C::QW, warm_scratch, soft,
            d_stats, ctx->d_err);
        B2_CUDA(cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        B2_CUDA(cudaFree(d_stats));
        const double nc = (double)std::max<unsigned long long>(1, hs[13]);
        fprintf(stderr,
                "[scan stats] grid %u rows %lld soft %d | events %llu cand %llu (per query-CTA %.1f) peel %llu | try %llu got %llu hard %llu | "
                "phaseB cand %llu | per-CTA us: A %.1f bar1 %.1f B %.1f bar2 %.1f C %.1f\n",
                grid, n_rows, soft, hs[0], hs[1], (double)hs[1] / ((double)grid * C::QW), hs[6], hs[2], hs[3], hs[4], hs[5],
                hs[8] / nc * 1e-3, hs[9] / nc * 1e-3, hs[10] / nc * 1e-3, hs[11] / nc * 1e-3, hs[12] / nc * 1e-3);
        ctx->launches++;
        return 0;
    }
    adc_scan_topk_kernel<G, WARPS_, false, VAR><<<grid, C::WARPS * 32, C::SMEM_BYTES, ctx->stream>>>(
        codesT, lut_scan, n_rows, n_gran, n_full, tail_desc, k, clamp, id_base, out_keys, qgroups * C::QW, warm_scratch, soft,
        nullptr, ctx->d_err);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

// floats of warm-up scratch the scan needs for a given plan (per CTA: 32 lanes' worth of queries x WARPS x 128 rows)
size_t scan_warm_scratch_floats(int M, int n_full, int n_tail) {
    return (size_t)(n_full + n_tail) * (size_t)(128 / M) * 16 * 128;
}

template <int G>
static int scan_dispatch(Ctx* ctx, const uint32_t* codesT, const float* lut_scan, long long n_rows, long long qgroups,
                         int n_full, int n_tail, const int4* tail_desc, int k, float clamp, uint32_t id_base,
                         unsigned long long* out_keys, float* warm_scratch) {
    // kernel variant: bit 0 = two-slot ring (G = 8 only), bit 1 = grouped threshold check; B200NN_SCAN_VAR overrides (tuning)
    int var = (G == 8) ? 3 : 2;
    if (const char* e = getenv("B200NN_SCAN_VAR")) var = atoi(e);
    if (G != 8) var &= 2;
#define B2_SCAN_VAR(V_)                                                                                                  \
    case V_:                                                                                                             \
        return scan_launch<G, 16, V_>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, \
                                      out_keys, warm_scratch);
    switch (var) {
        B2_SCAN_VAR(2)
        default: break;
    }
    if constexpr (G == 8) {
        switch (var) {
            B2_SCAN_VAR(1) B2_SCAN_VAR(3)
            default: break;
        }
    }
#undef B2_SCAN_VAR
    return scan_launch<G, 16, 0>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, out_keys,
                                 warm_scratch);
}

int launch_adc_scan_topk(Ctx* ctx, int M, const uint32_t* codesT, const float* lut_scan, long long n_rows, long long qgroups,
                         int n_full, int n_tail, const int* tail_desc_dev, int k, float clamp, uint32_t id_base,
                         unsigned long long* out_keys, float* warm_scratch) {
    const int4* tail_desc = reinterpret_cast<const int4*>(tail_desc_dev);
    if (n_tail > 0 && !tail_desc) B2_FAIL(-1, "adc_scan: tail CTAs without descriptors");
    if (k < 1 || k > KP) B2_FAIL(-4, "fused ADC top-k supports 1 <= k <= 128");
    switch (M) {
        case 4: return scan_dispatch<1>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, out_keys, warm_scratch);
        case 8: return scan_dispatch<2>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, out_keys, warm_scratch);
        case 16: return scan_dispatch<4>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, out_keys, warm_scratch);
        case 32: return scan_dispatch<8>(ctx, codesT, lut_scan, n_rows, qgroups, n_full, n_tail, tail_desc, k, clamp, id_base, out_keys, warm_scratch);
        default: B2_FAIL(-4, "fast ADC scan supports M in {4, 8, 16, 32}");
    }
}

int launch_frame_sum(Ctx* ctx, const float* scores, int n_frames, long long ng, float* total) {
    if (ng <= 0) return 0;
    frame_sum_kernel<<<grid_for(ng, 256, ctx->sm_count), 256, 0, ctx->stream>>>(scores, n_frames, ng, total);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_dense_topk_ex(Ctx* ctx, const float* values, long long rows, long long n, long long ld, int k, int slices, const uint32_t* id_map,
                         unsigned long long* out_keys) {
    if (rows <= 0) return 0;
    if (k < 1 || k > KP) B2_FAIL(-4, "dense top-k supports 1 <= k <= 128");
    if (n > 0xFFFFFFFFLL) B2_FAIL(-4, "dense top-k: more than 2^32 columns");
    dense_topk_kernel<<<dim3((unsigned)rows, (unsigned)std::max(1, slices)), 256, 0, ctx->stream>>>(values, n, ld, k, id_map, out_keys);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_dense_topk(Ctx* ctx, const float* values, long long rows, long long n, int k, unsigned long long* out_keys) {
    return launch_dense_topk_ex(ctx, values, rows, n, n, k, 1, nullptr, out_keys);
}

int launch_fill_f32(Ctx* ctx, float* p, long long n, float v) {
    if (n == 0) return 0;
    fill_f32_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(p, n, v);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_ivf_scan(Ctx* ctx, const float* lut, const int* probes, const long long* list_off, const unsigned char* codes_sorted,
                    const int* slot_sorted, int M, int ksub, long long nq, int nprobe, long long out_stride, float* out) {
    if (nq == 0) return 0;
    const size_t smem = (size_t)M * ksub * sizeof(float);
    if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(ivf_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ivf_scan_kernel<<<(unsigned)(nq * nprobe), 256, smem, ctx->stream>>>(lut, probes, list_off, codes_sorted, slot_sorted, M,
                                                                         ksub, nprobe, out_stride, out);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_ivf_search_topk(Ctx* ctx, const float* q_rot, long long nq, int D, const int* probes, int nprobe, const float* coarse,
                           const float* cb, int M, int ksub, const long long* list_off, const unsigned char* codes_sorted,
                           const int* row_sorted, long long n_rows, int k, float clamp, uint32_t id_base, unsigned long long* out_keys) {
    if (nq <= 0) return 0;
    if (k < 1 || k > KP) B2_FAIL(-4, "IVF search supports 1 <= k <= 128");
    const int ds = D / M;
    const size_t per_probe = ((size_t)M * ksub + D) * sizeof(float);
    const int n_pg = (int)std::max<size_t>(1, std::min<size_t>(std::min(IVF_PG, nprobe), (160u << 10) / per_probe));
    const size_t smem = (size_t)n_pg * per_probe;
#define B2_IVF(DS_)                                                                                                          \
    case DS_:                                                                                                                \
        if (smem > 40 * 1024)                                                                                                \
            B2_CUDA(cudaFuncSetAttribute(ivf_search_topk_kernel<DS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ivf_search_topk_kernel<DS_><<<(unsigned)nq, 256, smem, ctx->stream>>>(q_rot, nq, D, probes, nprobe, coarse, cb, M, ksub, \
                                                                             list_off, codes_sorted, row_sorted, n_rows, k, clamp, \
                                                                             id_base, n_pg, out_keys);                        \
        break;
    switch (ds) {
        B2_IVF(1) B2_IVF(2) B2_IVF(4) B2_IVF(8) B2_IVF(16) B2_IVF(32)
        default: B2_FAIL(-4, "ivf_search: unsupported sub-vector dimension");
    }
#undef B2_IVF
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
