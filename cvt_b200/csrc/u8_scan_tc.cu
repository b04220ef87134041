// u8_scan_tc.cu -- cfg2: exact L2 scan over uint8 vectors as a dense contraction on the tensor cores.
//
//   dist(q,x) = sum (q_i - x_i)^2 = |q|^2 + |x|^2 - 2 <q,x>        (L2SqrI, hnswlib/space_l2.h:186-219)
// <q,x> over uint8 is exact in int32 (128 * 255^2 = 8.3e6), so the scan is a u8 x u8 -> s32 GEMM:
// tcgen05.mma kind::i8, M = 128 queries, N = 256 database rows, K = D, accumulators in TMEM
// (2 x 256 columns, double buffered), followed by a fused epilogue that never writes distances:
// tcgen05.ld brings 32 lanes x 32 columns to registers, d' = |x|^2 - 2 acc is compared with the
// query's threshold, and the rare survivors go through the exact (dist, label) top-k lists.
//
// Roles (288 threads): warps 0-3 and 4-7 = two epilogue groups, group g drains columns [128 g, 128 g + 128)
// of every accumulator (so the MMA of the next tile always overlaps the drain of this one); inside a group
// thread t owns query t = TMEM lane t and keeps that query's sorted list PRIVATE: survivors are inserted at
// once by the owning thread -- no staging, no locks, no warp collectives -- and each group emits its own
// lists (2 output slices per CTA, merged by topk_merge_kernel).  Two groups = two warps per scheduler, which
// hides the dependent-ALU latency of the filter.  Warp 8, one elected thread = TMA producer + MMA issuer.
// (Measured and dropped: sharing a per-query distance bound between the CTAs of different slices through
// global atomics -- the k-th best of a partial list bounds the global k-th only as well as ONE slice does.)
//   B tiles (256 rows x D bytes) are stored in HBM already in the K-major no-swizzle core-matrix
//   order (u8_rows_to_canonical_kernel), so a tile is ONE contiguous TMA bulk copy; the rows' |x|^2 and
//   label ranks ride along as a second 2 KB copy, so the epilogue never touches global memory.
//   The epilogue keeps two tcgen05.ld of 32 columns in flight per warp (register double buffer) and
//   filters a chunk with one running minimum per query; the exact mask is built only for chunks that hit.
// A launch covers a RANGE of tiles: the caller (capi_flat.cu) scans a large index in passes of growing size -- a 16 k-row
// prefix, then 8x as many rows, then the rest -- and hands each pass the exact k-th best distance of everything scanned
// before it as the starting threshold (an upper bound on the final k-th best: exact, ties kept), so that by the time the bulk
// of the rows streams by, a 32 x 32 chunk of accumulators rarely holds a survivor at all.
// Supported: D % 32 == 0, D <= 256, k <= 120 as far as the lists fit shared memory beside the operand tiles (k <= 120 at
// D = 128 with one epilogue group); otherwise the dp4a kernel of flat_kernels.cu is used.
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "topk.cuh"
#include "u8_scan_tc.cuh"

namespace b200nn {

constexpr int TC_M = 128;      // queries per CTA (TMEM lanes)
constexpr int TC_N = 256;      // database rows per tile
constexpr int TC_KP = 120;     // most list slots per query (what fits shared memory at D = 128, one epilogue group)
constexpr int TC_MAX_SMEM = 232448;  // 227 KB
constexpr int TC_META_BYTES = 2 * TC_N * 4;  // per tile: 256 x |x|^2 then 256 x label rank

__device__ __forceinline__ uint64_t tc_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// THREAD-private top-k list of k keys in shared memory, kept UNSORTED with its maximum tracked in registers
// (tkey at slot tpos; KEY_MAX while a slot is still empty): an insertion overwrites the maximum and rescans the k
// slots for the new one -- k independent loads instead of a dependent shift chain.  Sorted once at the end.
// Layout: slot-major, slot j of query t at L_addr + (j * 128) * 8 with L_addr already offset by t: the threads of a warp
// walk the same slot of 32 different lists together, which is conflict-free for any k.
constexpr uint32_t TC_SLOT_STRIDE = 128u * 8u;
// `fill` = occupied slots: while the list is still filling, a key just takes the next free slot (no rescan; tkey stays
// KEY_MAX, so everything is still accepted); the maximum is looked up once when the k-th slot is taken and after every
// replacement from then on.
__device__ __forceinline__ void list_replace_max(uint32_t L_addr, unsigned long long key, int k, unsigned long long& tkey, int& tpos, int& fill) {
    if (fill < k) {
        sts64(L_addr + (uint32_t)fill * TC_SLOT_STRIDE, key);
        if (++fill < k) return;
    } else {
        sts64(L_addr + (uint32_t)tpos * TC_SLOT_STRIDE, key);
    }
    unsigned long long mx = 0;
    int mp = 0;
#pragma unroll 4
    for (int j = 0; j < k; j++) {
        const unsigned long long v = lds64(L_addr + (uint32_t)j * TC_SLOT_STRIDE);
        if (v >= mx) { mx = v; mp = j; }
    }
    tkey = mx;
    tpos = mp;
}

constexpr int TC_BOUND_LISTS = 32;          // lists per query whose minima are shared (shared-bound mode)
constexpr int TC_INF = 0x7f7f7f7f;          // "no distance yet" (memset pattern; above every real distance)

template <int TC_GROUPS>
__global__ void __launch_bounds__(TC_GROUPS * 128 + 64, 1)
u8_scan_tc_kernel(const unsigned char* __restrict__ xcan,   // [tiles][D/16][32][8][16] canonical B tiles
                  const int* __restrict__ xmeta,             // [tiles][2][256]: |x|^2 (padded rows: large), label rank
                  long long n, long long tile0, long long n_tiles,  // rows indexed; this launch scans tiles [tile0, tile0 + n_tiles)
                  int D, const unsigned char* __restrict__ queries, long long nq, int n_slices, int k,
                  const int* __restrict__ init_thr, int init_stride,  // optional: an upper bound on each query's k-th best distance
                  int* __restrict__ gmin,  // optional [nq][32], preset to TC_INF: shared-bound mode (see the bound warp below)
                  int NS,                                              // B-tile ring stages (2..4, as many as shared memory allows)
                  unsigned long long* __restrict__ out_keys /*[slice * groups + group][nq][k]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t A_BYTES = (uint32_t)TC_M * D, B_BYTES = (uint32_t)TC_N * D;
    const uint32_t sA = s_base;
    const uint32_t sB = sA + A_BYTES;                  // NS stages
    // row-meta ring (|x|^2 + label rank): MS = NS + 2 stages, because a tile's meta is read by its epilogue, which
    // may still run after the tile's shared-memory B stage has been released (the MMA retires first); stage t % MS
    // is rewritten for tile t + MS, whose load is issued only after the epilogue of tile t has signalled
    const int MS = NS + 2;
    const uint32_t sXN = sB + (uint32_t)NS * B_BYTES;  // MS stages x (256 norms + 256 ranks)
    const uint32_t sList = sXN + (uint32_t)MS * TC_META_BYTES;       // [groups][128][k] keys
    constexpr int TC_THREADS = TC_GROUPS * 128 + 64, TC_PRODUCER_WARP = TC_GROUPS * 4, TC_BOUND_WARP = TC_GROUPS * 4 + 1;
    const uint32_t sScratch = sList + (uint32_t)(TC_GROUPS * TC_M * k) * 8u;   // [warps][32 columns][32 lanes] words
    const uint32_t sQn = sScratch + TC_GROUPS * 4 * 32 * 32 * 4;             // [128] |q|^2
    const uint32_t bars = sQn + TC_M * 4;
    const uint32_t b_full = bars, b_empty = bars + 32, acc_full = bars + 64, acc_empty = bars + 80, tmem_slot = bars + 96;
    const uint32_t sDone = bars + 112;   // epilogue warps that have finished their tiles
    const uint32_t sTp = bars + 128;     // [128] shared bound on each query's k-th best DISTANCE (TC_INF: none yet)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long q0 = (long long)blockIdx.x * TC_M;
    const int slice = blockIdx.y;
    const long long t_lo = tile0 + (n_tiles * slice) / n_slices, t_hi = tile0 + (n_tiles * (slice + 1)) / n_slices;
    const int T = (int)(t_hi - t_lo);
    const uint32_t A_LBO = (TC_M / 8) * 128, B_LBO = (TC_N / 8) * 128, SBO = 128;
    const int ksteps = D / 32;

    if (warp == TC_PRODUCER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (lane == 0) {
            for (int i = 0; i < 4; i++) {
                mbar_init(b_full + 8 * i, 1);
                mbar_init(b_empty + 8 * i, 1);
            }
            for (int i = 0; i < 2; i++) {
                mbar_init(acc_full + 8 * i, 1);
                mbar_init(acc_empty + 8 * i, TC_GROUPS * 4);  // one arrival per epilogue warp
            }
            mbar_fence_init();
        }
    }
    // ---- queries -> canonical A tile, |q|^2, lists ----
    for (int i = tid; i < TC_GROUPS * TC_M * k; i += TC_THREADS) sts64(sList + (uint32_t)i * 8u, KEY_MAX);
    if (tid < TC_M) {
        int qn = 0;
        const long long qi = q0 + tid;
        const uint32_t row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;
        for (int c = 0; c < D / 16; c++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (qi < nq) v = __ldg(reinterpret_cast<const uint4*>(queries + qi * D) + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) qn = (int)__dp4a(w[e], w[e], (unsigned)qn);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sA + (uint32_t)c * A_LBO + row_off), "r"(v.x), "r"(v.y), "r"(v.z),
                         "r"(v.w)
                         : "memory");
        }
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(sQn + (uint32_t)tid * 4u), "r"(qn) : "memory");
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(sTp + (uint32_t)tid * 4u), "r"(TC_INF) : "memory");
        if (tid == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(sDone), "r"(0) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == TC_PRODUCER_WARP) {
        // ===== TMA producer + MMA issuer (one thread) =====
        if (lane == 0 && T > 0) {
            // D = S32 (2<<4), A = B = unsigned 8-bit (0), K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            auto load_tile = [&](int t) {
                const int sb = t % NS;
                mbar_arrive_expect_tx(b_full + 8 * sb, B_BYTES + TC_META_BYTES);
                tma_load_1d(sB + (uint32_t)sb * B_BYTES, xcan + (size_t)(t_lo + t) * B_BYTES, B_BYTES, b_full + 8 * sb);
                tma_load_1d(sXN + (uint32_t)(t % MS) * TC_META_BYTES, xmeta + (size_t)(t_lo + t) * 2 * TC_N, TC_META_BYTES, b_full + 8 * sb);
            };
            for (int t = 0; t < NS - 1 && t < T; t++) load_tile(t);  // NS-1 tiles in flight ahead of the MMA
            for (int t = 0; t < T; t++) {
                const int sb = t % NS, s = t & 1, use = t >> 1;
                mbar_wait(b_full + 8 * sb, (uint32_t)(t / NS) & 1u);
                if (t >= 2) mbar_wait(acc_empty + 8 * s, (uint32_t)(use - 1) & 1u);  // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)s * TC_N;
                for (int ks = 0; ks < ksteps; ks++) {
                    const uint64_t da = tc_desc_kmajor(sA + (uint32_t)ks * 2 * A_LBO, A_LBO, SBO);
                    const uint64_t db = tc_desc_kmajor(sB + (uint32_t)sb * B_BYTES + (uint32_t)ks * 2 * B_LBO, B_LBO, SBO);
                    umma_i8(acc, da, db, idesc, ks != 0);
                }
                tc_commit(b_empty + 8 * sb);  // smem stage reusable once these MMAs retire
                tc_commit(acc_full + 8 * s);  // accumulator ready for the epilogue
                // refill: tile t+NS-1 goes to the stage tile t-1 used, once the MMAs of tile t-1 have retired
                const int nxt = t + NS - 1;
                if (nxt < T) {
                    if (t >= 1) mbar_wait(b_empty + 8 * ((t - 1) % NS), (uint32_t)((t - 1) / NS) & 1u);
                    load_tile(nxt);
                }
            }
        }
    } else if (warp == TC_BOUND_WARP) {
        // ===== shared bound (one-pass mode) =====
        // Every (slice, group) list of a query covers its own rows, so the k-th smallest of the lists' MINIMA is the
        // distance of at least k distinct rows: an upper bound on the query's final k-th best that is almost as tight as a
        // merge of the lists would give (the k best rows mostly sit in different lists), available after the first few
        // tiles and without any pass structure or barrier.  The first TC_BOUND_LISTS lists of a query publish their
        // minimum (a plain store whenever it improves); this warp keeps re-reading them (thread = query, 4 rounds),
        // selects the k-th smallest by bisection in registers and posts it in shared memory, where the epilogue
        // threads pick it up once per tile.  Nothing ever waits on it: a stale value is merely a weaker bound.
        if (gmin != nullptr) {
            unsigned sleep_ns = 0;  // the bound moves fast at first and hardly at all later: back off
            for (int r = 0;; r = (r + 1) & (TC_M / 32 - 1)) {
                int done = 0;
                if (lane == 0) asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(done) : "r"(sDone));
                done = __shfl_sync(0xffffffffu, done, 0);
                if (done >= TC_GROUPS * 4) break;
                {
                    const int q = r * 32 + lane;
                    int v[TC_BOUND_LISTS];
                    const int4* src = reinterpret_cast<const int4*>(gmin + (q0 + q) * TC_BOUND_LISTS);
#pragma unroll
                    for (int i = 0; i < TC_BOUND_LISTS / 4; i++) {
                        int4 t = make_int4(TC_INF, TC_INF, TC_INF, TC_INF);
                        if (q0 + q < nq) t = __ldcg(src + i);
                        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
                    }
                    int lo = TC_INF, hi = -1, fin = 0;
#pragma unroll
                    for (int i = 0; i < TC_BOUND_LISTS; i++)
                        if (v[i] < TC_INF) { fin++; lo = min(lo, v[i]); hi = max(hi, v[i]); }
                    int res = TC_INF;
                    if (fin >= k) {
                        while (lo < hi) {  // smallest value with at least k minima at or below it
                            const int mid = lo + ((hi - lo) >> 1);
                            int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
                            for (int i = 0; i < TC_BOUND_LISTS; i += 4) {
                                c0 += (v[i] <= mid) ? 1 : 0; c1 += (v[i + 1] <= mid) ? 1 : 0;
                                c2 += (v[i + 2] <= mid) ? 1 : 0; c3 += (v[i + 3] <= mid) ? 1 : 0;
                            }
                            if ((c0 + c1) + (c2 + c3) >= k) hi = mid; else lo = mid + 1;
                        }
                        res = lo;
                    }
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(sTp + (uint32_t)q * 4u), "r"(res) : "memory");
                }
                if (r == TC_M / 32 - 1) {
                    if (sleep_ns) __nanosleep(sleep_ns);
                    sleep_ns = min(4000u, sleep_ns * 2u + 250u);
                }
            }
        }
    } else {
        // ===== epilogue group g = warp / 4: thread owns query ql = tid % 128 (TMEM lane ql), columns [128 g, 128 g + 128) of every tile =====
        const int grp = warp >> 2, ql = tid & (TC_M - 1);
        const bool qvalid = q0 + ql < nq;
        const uint32_t myList = sList + (uint32_t)(grp * TC_M * k + ql) * 8u;  // slot j at + j * TC_SLOT_STRIDE
        const uint32_t scratch = sScratch + (uint32_t)warp * (32 * 32 * 4) + (uint32_t)lane * 4u;
        int qn;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(qn) : "r"(sQn + (uint32_t)ql * 4u));
        // tp = threshold on d' = |x|^2 - 2<q,x>  (dist - |q|^2): the caller's bound (the k-th best distance over a
        // sample of the rows: a row beyond it cannot be in the top-k; ties are kept), tightened by the own list once full
        int tp = 0x7fffffff;
        if (init_thr != nullptr && qvalid) tp = __ldg(init_thr + (q0 + ql) * init_stride) - qn;
        unsigned long long tkey = KEY_MAX;
        int tpos = 0, fill = 0;
        // shared-bound mode: this list publishes its minimum if it is one of the query's first TC_BOUND_LISTS lists
        const int list_id = slice * TC_GROUPS + grp;
        const bool publish = gmin != nullptr && qvalid && list_id < TC_BOUND_LISTS;
        int* const gslot = gmin + (q0 + ql) * TC_BOUND_LISTS + list_id;
        int lmin = TC_INF;
        auto tld32 = [&](uint32_t (&r)[32], uint32_t taddr) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
                  "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
                  "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
        };
        // one chunk of 32 columns: d' = |x|^2 - 2 acc, one running minimum per query as the filter; a query
        // that hits walks its 32 values and inserts the survivors straight into its own sorted list
        // (thread-private: no staging, no warp collectives, the threshold tightens at once)
        auto process = [&](uint32_t (&r)[32], int c0, uint32_t meta_s, long long row_base) {
            int mn = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const uint4 xv = lds128(meta_s + (uint32_t)(c0 + i) * 4u);  // broadcast: |x|^2 of 4 columns
                r[i] = (uint32_t)((int)xv.x - 2 * (int)r[i]);
                r[i + 1] = (uint32_t)((int)xv.y - 2 * (int)r[i + 1]);
                r[i + 2] = (uint32_t)((int)xv.z - 2 * (int)r[i + 2]);
                r[i + 3] = (uint32_t)((int)xv.w - 2 * (int)r[i + 3]);
                mn = min(mn, min((int)r[i], (int)r[i + 1]));
                mn = min(mn, min((int)r[i + 2], (int)r[i + 3]));
            }
            if (qvalid && mn <= tp) {
                // park this lane's 32 values (word i*32+lane: conflict-free) so that the few passing columns can be
                // fetched by dynamic index, and walk the mask of passing columns
                uint32_t pm = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(scratch + (uint32_t)(i * 128)), "r"(r[i]) : "memory");
                    if ((int)r[i] <= tp) pm |= 1u << i;
                }
                while (pm) {
                    const int i = __ffs(pm) - 1;
                    pm &= pm - 1;
                    int dp;
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(dp) : "r"(scratch + (uint32_t)(i * 128)));
                    if (dp <= tp && row_base + c0 + i < n) {  // tp may have tightened since the mask was built
                        uint32_t rk;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rk) : "r"(meta_s + (uint32_t)(TC_N + c0 + i) * 4u));
                        const unsigned long long key = make_key(s32_orderable(dp + qn), rk);
                        if (key < tkey) {
                            list_replace_max(myList, key, k, tkey, tpos, fill);
                            if (publish && dp + qn < lmin) {
                                lmin = dp + qn;
                                asm volatile("st.global.cg.s32 [%0], %1;" ::"l"(gslot), "r"(lmin) : "memory");
                            }
                            if (tkey != KEY_MAX) tp = min(tp, s32_from_orderable((uint32_t)(tkey >> 32)) - qn);  // list full: k-th distance
                        }
                    }
                }
            }
            __syncwarp();
        };
        for (int t = 0; t < T; t++) {
            const int s = t & 1, use = t >> 1;
            mbar_wait(acc_full + 8 * s, (uint32_t)use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row_base = (t_lo + t) * TC_N;
            if (gmin != nullptr) {  // the bound posted by the bound warp (a DISTANCE; tp is relative to |q|^2)
                int sh;
                asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(sh) : "r"(sTp + (uint32_t)ql * 4u));
                tp = min(tp, sh - qn);
            }
            const uint32_t meta_s = sXN + (uint32_t)(t % MS) * TC_META_BYTES;
            const int cg = grp * (TC_N / TC_GROUPS);
            const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * TC_N + cg);
            uint32_t ra[32], rb[32];
            tld32(ra, tacc);
#pragma unroll 1
            for (int c0 = 0; c0 < TC_N / TC_GROUPS; c0 += 64) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tld32(rb, tacc + (uint32_t)(c0 + 32));  // in flight while chunk c0 is filtered
                process(ra, cg + c0, meta_s, row_base);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 64 < TC_N / TC_GROUPS) tld32(ra, tacc + (uint32_t)(c0 + 64));
                process(rb, cg + c0 + 32, meta_s, row_base);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * s);
        }
        __syncwarp();
        if (lane == 0) asm volatile("red.shared.add.s32 [%0], 1;" ::"r"(sDone) : "memory");
        if (qvalid) {  // emit the list ascending: rank of an entry = how many entries order before it
            unsigned long long* out = out_keys + ((long long)(slice * TC_GROUPS + grp) * nq + q0 + ql) * k;
            for (int j = 0; j < k; j++) {
                const unsigned long long e = lds64(myList + (uint32_t)j * TC_SLOT_STRIDE);
                int rank = 0;
                for (int i = 0; i < k; i++) {
                    const unsigned long long o = lds64(myList + (uint32_t)i * TC_SLOT_STRIDE);
                    rank += (o < e || (o == e && i < j)) ? 1 : 0;  // empty slots (KEY_MAX) tie: index order
                }
                out[rank] = e;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_PRODUCER_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// row-major u8 rows -> canonical B tiles + per-tile row meta (|x|^2, label rank)
__global__ void u8_rows_to_canonical_kernel(const unsigned char* __restrict__ rows, const uint32_t* __restrict__ rank, long long n, int D,
                                            unsigned char* __restrict__ xcan, int* __restrict__ xmeta, long long n_pad) {
    // one thread per (row, 16-byte chunk)
    const int chunks = D / 16;
    const long long total = n_pad * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / chunks;
        const int c = (int)(i - row * chunks);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < n) v = *reinterpret_cast<const uint4*>(rows + row * D + c * 16);
        const long long tile = row / TC_N;
        const int rl = (int)(row - tile * TC_N);
        unsigned char* dst = xcan + (size_t)tile * TC_N * D + (size_t)c * (TC_N / 8) * 128 + (size_t)(rl >> 3) * 128 + (rl & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = v;
    }
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_pad; row += (long long)gridDim.x * blockDim.x) {
        int s = 0x3fffffff;  // padded rows can never beat a real one
        if (row < n) {
            s = 0;
            for (int j = 0; j < D; j++) {
                const int v = rows[row * D + j];
                s += v * v;
            }
        }
        const long long tile = row / TC_N;
        const int rl = (int)(row - tile * TC_N);
        xmeta[tile * 2 * TC_N + rl] = s;
        xmeta[tile * 2 * TC_N + TC_N + rl] = row < n ? (int)rank[row] : -1;
    }
}

static size_t tc_smem_bytes(int D, int k, int groups, int stages) {
    return (size_t)TC_M * D + (size_t)stages * TC_N * D + (size_t)(stages + 2) * TC_META_BYTES +
           (size_t)groups * (TC_M * (size_t)k * 8 + 4 * 32 * 32 * 4) + TC_M * 4 + 128 + TC_M * 4 + 128;
}
// Two epilogue groups when their lists and scratch fit beside the operand tiles, else one.  (Four groups -- 16 epilogue warps,
// 64 accumulator columns each -- were measured: the bulk pass of cfg2 went from 233 to 223 us only, because the drain is
// bound by the TMEM read port, 64 B/clk per SM = 2048 clk for the 128 KB of int32 accumulators of a tile, not by issue or
// latency; the extra lists made the merges dearer than that gain.)
int u8_scan_tc_bound_lists() { return TC_BOUND_LISTS; }
int u8_scan_tc_lists_per_slice(int D, int k) { return tc_smem_bytes(D, k, 2, 2) <= (size_t)TC_MAX_SMEM ? 2 : 1; }

bool u8_scan_tc_supported(int D, int k) {
    return D % 32 == 0 && D >= 32 && D <= 256 && k >= 1 && k <= TC_KP && tc_smem_bytes(D, k, 1, 2) <= (size_t)TC_MAX_SMEM;
}

int launch_u8_rows_to_canonical(Ctx* ctx, const unsigned char* rows, const uint32_t* rank, long long n, int D, unsigned char* xcan,
                                int* xmeta, long long n_pad) {
    if (n_pad <= 0) return 0;
    const long long work = n_pad * (D / 16);
    u8_rows_to_canonical_kernel<<<(unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, (long long)ctx->sm_count * 16)), 256,
                                  0, ctx->stream>>>(rows, rank, n, D, xcan, xmeta, n_pad);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int u8_scan_tc_slices(int sm_count, long long nq, long long n_tiles, int min_tiles_per_slice) {
    const long long qt = (nq + TC_M - 1) / TC_M, tiles = n_tiles;
    long long s = sm_count / qt;  // one wave of one-CTA-per-SM
    s = std::max<long long>(1, std::min<long long>(s, std::max<long long>(1, tiles / std::max(1, min_tiles_per_slice))));
    return (int)std::min<long long>(s, 1024);
}

int launch_u8_scan_tc(Ctx* ctx, const unsigned char* xcan, const int* xmeta, long long n, long long tile0, long long n_tiles, int D,
                      const unsigned char* queries, long long nq, int n_slices, int k, const int* init_thr, int init_stride, int* gmin,
                      unsigned long long* out_keys) {
    if (nq <= 0 || n_tiles <= 0) return 0;
    if (!u8_scan_tc_supported(D, k)) B2_FAIL(-4, "u8 tensor-core scan: needs D % 32 == 0, D <= 256 and k lists that fit shared memory");
    const int groups = u8_scan_tc_lists_per_slice(D, k);
    int stages = 4;  // B-tile ring: as deep as shared memory allows (hides the TMA latency behind two epilogues and more)
    while (stages > 2 && tc_smem_bytes(D, k, groups, stages) > (size_t)TC_MAX_SMEM) stages--;
    const size_t smem = tc_smem_bytes(D, k, groups, stages);
    dim3 grid((unsigned)((nq + TC_M - 1) / TC_M), (unsigned)n_slices);
    if (groups == 2) {
        B2_CUDA(cudaFuncSetAttribute(u8_scan_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        u8_scan_tc_kernel<2><<<grid, 2 * 128 + 64, smem, ctx->stream>>>(xcan, xmeta, n, tile0, n_tiles, D, queries, nq, n_slices, k, init_thr, init_stride, gmin, stages, out_keys);
    } else {
        B2_CUDA(cudaFuncSetAttribute(u8_scan_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        u8_scan_tc_kernel<1><<<grid, 1 * 128 + 64, smem, ctx->stream>>>(xcan, xmeta, n, tile0, n_tiles, D, queries, nq, n_slices, k, init_thr, init_stride, gmin, stages, out_keys);
    }
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
