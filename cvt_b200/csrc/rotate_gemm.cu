// rotate_gemm.cu -- dense projections Y[n,N] = (X[n,K] - mean) * V^T on the tensor cores.  Two users:
//   a1  the OPQ rotation as a dense matrix (K = N = D, no mean), bit-identical to the reference's permutation;
//   f-3 the PCA front end cvtk::PCAUtils::reduceDim (pca_train_project/pca_online/pca_utils.cc:25-35):
//       cv::PCA::project = (x - mean) * eigenvectors^T, then the per-row L2 normalisation, fused in the epilogue.
// Written for a1 first, hence the name:
// rotate_gemm.cu -- a1 as a dense rotation: Y[n,D] = X[n,D] * R^T on the 5th-generation tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM), at ~fp32 accuracy through an error-free
// split of both operands ("3xTF32"):
//     x = x0 + x1 + x2   (each part has <= 11 significant bits: exactly representable in TF32)
//     R = r0 + r1        (pre-split once per model on the host)
//     y = x2*r0 + x1*r1 + x1*r0 + x0*r1 + x0*r0        (small terms first)
// Every product of TF32-exact operands is exact in fp32; the only roundings are the fp32
// accumulations in TMEM, and those TRUNCATE (measured: a single accumulator over K = 1024 ends ~60 ulp
// short, always toward zero).  So the leading product x0*r0 and the four correction products (2^-11 of the
// result and less) accumulate in two separate TMEM column ranges and are added once in the epilogue:
// one truncation per k-step on the leading sum instead of five.  For a PERMUTATION matrix (r1 = 0, entries 0/1) every partial sum is
// exactly representable, so the result is bit-identical to the gather of IVFOPQ::reorder
// (opq/src/IVFOPQ.cpp:424-439) -- tests/test_rotate_gemm_gpu.py checks that on the device.
//
// Tile: one CTA computes 128 rows x NB (64|128) output columns; K is walked in chunks of 32.
//   A (rows of X): read with 16-byte global loads, split in registers, written as three K-major
//      no-swizzle "core matrix" tiles (8 rows x 16 bytes, LBO = K-direction, SBO = row-group direction).
//   B (R parts):   pre-arranged in global memory in the same canonical order, so a chunk is ONE
//      contiguous TMA bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier.
//   MMA: thread 0 issues 4 k-steps x 5 products of tcgen05.mma (M=128, N=NB, K=8) per chunk and
//      tcgen05.commit's to the barrier that frees the double-buffered stage.
//   Epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> 16-byte global stores.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "rotate_gemm.cuh"

namespace b200nn {

constexpr int RG_M = 128;   // rows per tile (= TMEM lanes)
constexpr int RG_KC = 32;   // K elements per chunk (4 MMA k-steps of 8)
constexpr int RG_THREADS = 128;

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | swizzle none
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// MODE 0: plain store.  MODE 1 (needs N == NB, one CTA owns whole output rows): the reference's row normalisation
//   normMat = row * row.t() (OpenCV gemm: float products accumulated in double, stored as float);
//   denomv = max(1e-12, (double)sqrt(normMat)) as float; row /= denomv        (pca_utils.cc:28-34)
template <int NB, int MODE>
__global__ void __launch_bounds__(RG_THREADS, 1)
rotate_gemm_tf32x3_kernel(const float* __restrict__ x, long long n, int K, int N, const float* __restrict__ mean,
                          const float* __restrict__ bplanes, float* __restrict__ y) {
    constexpr uint32_t A_PLANE = RG_M * RG_KC * 4;       // 16 KB
    constexpr uint32_t B_PLANE = NB * RG_KC * 4;         // 8 / 16 KB
    constexpr uint32_t A_LBO = RG_M * 16, B_LBO = NB * 16, SBO = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t sA = s_base;                            // [2 stages][3 planes][A_PLANE]
    const uint32_t sB = sA + 2 * 3 * A_PLANE;              // [2 stages][2 planes][B_PLANE]
    const uint32_t bars = sB + 2 * 2 * B_PLANE;            // b_full[2], stage_free[2], accum_full, tmem slot
    const uint32_t b_full = bars, stage_free = bars + 16, accum_full = bars + 32, tmem_slot = bars + 48;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nblk = blockIdx.x;                           // output column block (fastest: shares X rows in L2)
    const long long tile = blockIdx.y;
    const int nchunks = K / RG_KC;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * NB)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(b_full, 1); mbar_init(b_full + 8, 1);
        mbar_init(stage_free, 1); mbar_init(stage_free + 8, 1);
        mbar_init(accum_full, 1);
        mbar_fence_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(RG_M >> 4) << 24);

    const long long row = tile * RG_M + tid;
    const bool valid = row < n;
    const float* xrow = x + row * K;
    const uint32_t a_row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;

    // the row's next K chunk is fetched into registers while the current one is split and multiplied (the global-load
    // latency would otherwise sit on the critical path of every chunk: 4 warps per SM cannot hide it by themselves).
    // Measured: 0.499 -> 0.435 ms for 131 072 x 1024 -> 128.  Fetching THREE chunks ahead (four rotating buffers, 190
    // registers) was tried and measured slower (0.481 ms; 0.894 vs 0.653 ms at 2048 -> 256): dropped.
    float4 cur[RG_KC / 4], nxt[RG_KC / 4];
    auto fetch = [&](float4 (&v)[RG_KC / 4], int kc) {
#pragma unroll
        for (int c = 0; c < RG_KC / 4; c++)
            v[c] = valid ? __ldg(reinterpret_cast<const float4*>(xrow + kc * RG_KC + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch(cur, 0);
    for (int kc = 0; kc < nchunks; kc++) {
        const int st = kc & 1, use = kc >> 1;
        if (kc + 1 < nchunks) fetch(nxt, kc + 1);
        if (kc >= 2) mbar_wait(stage_free + 8 * st, (uint32_t)(use - 1) & 1u);  // MMAs of chunk kc-2 have drained this stage
        if (tid == 0) {
            mbar_arrive_expect_tx(b_full + 8 * st, 2 * B_PLANE);
            tma_load_1d(sB + (uint32_t)st * 2 * B_PLANE, bplanes + ((size_t)nblk * nchunks + kc) * (2 * NB * RG_KC), 2 * B_PLANE,
                        b_full + 8 * st);
        }
        const uint32_t a0 = sA + (uint32_t)st * 3 * A_PLANE + a_row_off;
#pragma unroll
        for (int c = 0; c < RG_KC / 4; c++) {
            float4 v = cur[c];
            if (mean != nullptr) {  // cv::PCA::project subtracts the mean first (one fp32 rounding), then multiplies
                const float4 mv = __ldg(reinterpret_cast<const float4*>(mean + kc * RG_KC + c * 4));
                v.x = __fsub_rn(v.x, mv.x); v.y = __fsub_rn(v.y, mv.y); v.z = __fsub_rn(v.z, mv.z); v.w = __fsub_rn(v.w, mv.w);
            }
            const float h0 = tf32_trunc(v.x), h1 = tf32_trunc(v.y), h2 = tf32_trunc(v.z), h3 = tf32_trunc(v.w);
            const float r0 = __fsub_rn(v.x, h0), r1 = __fsub_rn(v.y, h1), r2 = __fsub_rn(v.z, h2), r3 = __fsub_rn(v.w, h3);
            const float m0 = tf32_trunc(r0), m1 = tf32_trunc(r1), m2 = tf32_trunc(r2), m3 = tf32_trunc(r3);
            const float l0 = tf32_trunc(__fsub_rn(r0, m0)), l1 = tf32_trunc(__fsub_rn(r1, m1)), l2 = tf32_trunc(__fsub_rn(r2, m2)),
                        l3 = tf32_trunc(__fsub_rn(r3, m3));
            const uint32_t o = a0 + (uint32_t)c * A_LBO;
            sts128(o, h0, h1, h2, h3);
            sts128(o + A_PLANE, m0, m1, m2, m3);
            sts128(o + 2 * A_PLANE, l0, l1, l2, l3);
        }
#pragma unroll
        for (int c = 0; c < RG_KC / 4; c++) cur[c] = nxt[c];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            mbar_wait(b_full + 8 * st, (uint32_t)use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t A = sA + (uint32_t)st * 3 * A_PLANE, B = sB + (uint32_t)st * 2 * B_PLANE;
#pragma unroll
            for (int s = 0; s < RG_KC / 8; s++) {
                const uint32_t ak = (uint32_t)s * 2 * A_LBO, bk = (uint32_t)s * 2 * B_LBO;
                const uint64_t da0 = umma_desc_kmajor(A + ak, A_LBO, SBO), da1 = umma_desc_kmajor(A + A_PLANE + ak, A_LBO, SBO),
                               da2 = umma_desc_kmajor(A + 2 * A_PLANE + ak, A_LBO, SBO);
                const uint64_t db0 = umma_desc_kmajor(B + bk, B_LBO, SBO), db1 = umma_desc_kmajor(B + B_PLANE + bk, B_LBO, SBO);
                // two accumulators (the tensor core TRUNCATES every fp32 accumulation): the four correction products,
                // 2^-11 and less of the result, go to columns [NB, 2NB); the leading product x0*r0 to columns [0, NB).
                // The first MMA into an accumulator overwrites it.  The epilogue adds the two (one rounding).
                umma_tf32(tmem_base + NB, da2, db0, idesc, (kc | s) != 0);
                umma_tf32(tmem_base + NB, da1, db1, idesc, 1);
                umma_tf32(tmem_base + NB, da1, db0, idesc, 1);
                umma_tf32(tmem_base + NB, da0, db1, idesc, 1);
                umma_tf32(tmem_base, da0, db0, idesc, (kc | s) != 0);
            }
            umma_commit(stage_free + 8 * st);
            if (kc == nchunks - 1) umma_commit(accum_full);
        }
    }

    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* yrow = y + row * N + (size_t)nblk * NB;
    auto tld32 = [&](uint32_t (&r)[32], int c0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    };
    // y = leading accumulator + correction accumulator, one fp32 rounding
    auto tld_sum = [&](float (&v)[32], int c0) {
        uint32_t r[32], q[32];
        tld32(r, c0);
        tld32(q, NB + c0);
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = __fadd_rn(__uint_as_float(r[i]), __uint_as_float(q[i]));
    };
    float denomv = 1.0f;
    if (MODE == 1) {  // thread = row: sum of squares over the row's N outputs, in column order, accumulated in double
        double ss = 0.0;
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 32) {
            float v[32];
            tld_sum(v, c0);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const double dv = (double)v[i];
                ss = __dadd_rn(ss, __dmul_rn(dv, dv));
            }
        }
        const double d = (double)__fsqrt_rn((float)ss);
        denomv = (float)(d > 1e-12 ? d : 1e-12);
    }
#pragma unroll 1
    for (int c0 = 0; c0 < NB; c0 += 32) {
        float v[32];
        tld_sum(v, c0);
        if (valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (MODE == 1) { o.x = __fdiv_rn(o.x, denomv); o.y = __fdiv_rn(o.y, denomv); o.z = __fdiv_rn(o.z, denomv); o.w = __fdiv_rn(o.w, denomv); }
                *reinterpret_cast<float4*>(yrow + c0 + i) = o;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * NB)) : "memory");
}

// ---- host ----
int rotate_gemm_nblock(int D) { return (D % 128 == 0) ? 128 : 64; }

bool rotate_gemm_supported(int D) { return D >= 64 && D % 64 == 0; }

// output-column block of a projection to N columns: with the fused row normalisation one CTA must own whole rows
int proj_gemm_nblock(int N, bool l2norm) {
    if (l2norm) return (N == 64 || N == 128 || N == 256) ? N : 0;
    return (N % 64 == 0 && N >= 64) ? ((N % 128 == 0) ? 128 : 64) : 0;
}
bool proj_gemm_supported(int K, int N, bool l2norm) { return K >= 32 && K % 32 == 0 && proj_gemm_nblock(N, l2norm) != 0; }

// V [N,K] row-major (y = V x) -> two TF32-exact parts in the canonical K-major core-matrix order,
// one contiguous block per (output block, K chunk): [nblk][kc][part][ (k/4) | (n/8) | n%8 | k%4 ]
void proj_gemm_pack(const float* V, int N, int K, int NB, std::vector<float>& out) {
    const int nblocks = N / NB, nchunks = K / RG_KC;
    out.assign((size_t)N * K * 2, 0.0f);
    auto trunc = [](float v) {
        union { float f; uint32_t u; } c;
        c.f = v;
        c.u &= 0xFFFFE000u;
        return c.f;
    };
    for (int nb = 0; nb < nblocks; nb++)
        for (int kc = 0; kc < nchunks; kc++) {
            float* blk = out.data() + ((size_t)nb * nchunks + kc) * (2 * NB * RG_KC);
            for (int nl = 0; nl < NB; nl++)
                for (int kl = 0; kl < RG_KC; kl++) {
                    const float v = V[(size_t)(nb * NB + nl) * K + kc * RG_KC + kl];
                    const float r0 = trunc(v), r1 = trunc(v - r0);
                    const size_t off = (size_t)(kl / 4) * (NB * 4) + (size_t)(nl / 8) * 32 + (nl % 8) * 4 + (kl % 4);
                    blk[off] = r0;
                    blk[(size_t)NB * RG_KC + off] = r1;
                }
        }
}

void rotate_gemm_pack_R(const float* R, int D, std::vector<float>& out) { proj_gemm_pack(R, D, D, rotate_gemm_nblock(D), out); }

template <int NB, int MODE>
static int proj_launch(Ctx* ctx, const float* x, long long n, int K, int N, const float* mean, const float* bplanes, float* y) {
    const size_t smem = 2 * 3 * (size_t)RG_M * RG_KC * 4 + 2 * 2 * (size_t)NB * RG_KC * 4 + 64;
    B2_CUDA(cudaFuncSetAttribute(rotate_gemm_tf32x3_kernel<NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_tiles = (n + RG_M - 1) / RG_M;
    const long long max_tiles = 65535;  // gridDim.y limit: fold the row tiles into several launches
    for (long long t0 = 0; t0 < n_tiles; t0 += max_tiles) {
        const long long tiles = std::min<long long>(max_tiles, n_tiles - t0);
        const long long rows0 = t0 * RG_M, rows = std::min<long long>(n - rows0, tiles * RG_M);
        dim3 g((unsigned)(N / NB), (unsigned)tiles);
        rotate_gemm_tf32x3_kernel<NB, MODE><<<g, RG_THREADS, smem, ctx->stream>>>(x + rows0 * K, rows, K, N, mean, bplanes, y + rows0 * N);
        ctx->launches++;
    }
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_proj_gemm(Ctx* ctx, const float* x, long long n, int K, int N, const float* mean, const float* bplanes, bool l2norm, float* y) {
    if (n <= 0) return 0;
    if (!proj_gemm_supported(K, N, l2norm)) B2_FAIL(-4, "projection GEMM needs K % 32 == 0 and N % 64 == 0 (N in {64, 128, 256} with the fused L2 normalisation)");
    const int NB = proj_gemm_nblock(N, l2norm);
    if (l2norm) {
        if (NB == 64) return proj_launch<64, 1>(ctx, x, n, K, N, mean, bplanes, y);
        if (NB == 128) return proj_launch<128, 1>(ctx, x, n, K, N, mean, bplanes, y);
        return proj_launch<256, 1>(ctx, x, n, K, N, mean, bplanes, y);
    }
    if (NB == 128) return proj_launch<128, 0>(ctx, x, n, K, N, mean, bplanes, y);
    return proj_launch<64, 0>(ctx, x, n, K, N, mean, bplanes, y);
}

int launch_rotate_gemm(Ctx* ctx, const float* x, long long n, int D, const float* bplanes, float* y) {
    if (n <= 0) return 0;
    if (!rotate_gemm_supported(D)) B2_FAIL(-4, "dense rotation needs D % 64 == 0");
    return launch_proj_gemm(ctx, x, n, D, D, nullptr, bplanes, false, y);
}

}  // namespace b200nn
