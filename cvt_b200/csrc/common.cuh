// common.cuh -- shared host/device helpers for the b200nn CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace b200nn {

// ------------------------------------------------------------------------------------ errors
void set_last_error(const std::string& msg);

#define B2_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            b200nn::set_last_error(std::string(#call) + ": " + cudaGetErrorString(_e) +     \
                                   " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

#define B2_FAIL(code, msg)              \
    do {                                \
        b200nn::set_last_error(msg);    \
        return (code);                  \
    } while (0)

// ------------------------------------------------------------------------------------ context
struct Ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // own_stream or an adopted external stream
    uint64_t launches = 0;
    cudaEvent_t events[16] = {};
    int* d_err = nullptr;  // device-side error flag (smem layout overflow etc.)
    float* dmat = nullptr;  // row-chunk distance matrix of the tiled nearest-centroid path (dist_tile.cu), grown on demand
    size_t dmat_elems = 0;
    // host -> device staging for the add paths: a copy stream, two pinned buffers and their events (created on first use)
    cudaStream_t copy_stream = nullptr;
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    cudaEvent_t copy_done[2] = {}, compute_done[2] = {};
};
int ensure_copy_engine(Ctx* ctx, size_t bytes_per_buffer);  // capi_ctx.cu

// ------------------------------------------------------------------------------------ keys
// 64-bit sortable record: (orderable(dist) << 32) | id.  Comparing records as unsigned integers
// is exactly the reference's lexicographic (dist, id) order (std::pair<float,uint> operator<,
// opq/src/common.h:35; brutoforce.hpp:81-91).
__host__ __device__ __forceinline__ uint32_t f32_orderable(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_orderable(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint32_t s32_orderable(int32_t v) { return (uint32_t)v ^ 0x80000000u; }
__host__ __device__ __forceinline__ int32_t s32_from_orderable(uint32_t o) { return (int32_t)(o ^ 0x80000000u); }
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t ord, uint32_t id) { return ((uint64_t)ord << 32) | id; }
static constexpr uint64_t KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared through the TMA engine (SASS: UBLKCP), completion on mbarrier.
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long lds64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
#endif  // __CUDACC__

inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace b200nn
