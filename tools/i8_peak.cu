// i8_peak.cu -- the int8 tensor-core peak of this GPU, measured: a bare tcgen05.mma kind::i8 loop (u8 x u8 -> s32, M = 128,
// N = 256, K = 32 per instruction, operands resident in shared memory in the same K-major no-swizzle core-matrix layout the
// scan kernel uses, accumulator in TMEM), one CTA per SM, no loads, no epilogue.  This is the denominator of the
// tensor-bound roofline of u8_scan_tc_kernel (bench.py reads profiles/i8_mma_peak.json) instead of the "2 x bf16" proxy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/i8_peak tools/i8_peak.cu && tools/bin/i8_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

template <int KSTEPS>
__global__ void __launch_bounds__(128, 1) i8_mma_loop(int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    constexpr int M = 128, N = 256, D = 32 * KSTEPS;
    const uint32_t sA = smem_u32(smem), sB = sA + M * D;
    for (int i = threadIdx.x; i < (M + N) * D / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01020304u * (i & 7);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_a = smem_u32(&bar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t A_LBO = (M / 8) * 128, B_LBO = (N / 8) * 128, SBO = 128;
        for (int it = 0; it < iters; it++) {
            const uint32_t acc = tmem + (uint32_t)(it & 1) * N;  // two accumulators, as the scan kernel alternates
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                const uint64_t da = desc_kmajor(sA + ks * 2 * A_LBO, A_LBO, SBO), db = desc_kmajor(sB + ks * 2 * B_LBO, B_LBO, SBO);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(ks != 0))
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a)
            : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    constexpr int KSTEPS = 4;  // D = 128: the cfg2 tile
    const int smem = (128 + 256) * 32 * KSTEPS;
    cudaFuncSetAttribute(i8_mma_loop<KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int sms = prop.multiProcessorCount, iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    i8_mma_loop<KSTEPS><<<sms, 128, smem>>>(200);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        i8_mma_loop<KSTEPS><<<sms, 128, smem>>>(iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
    const double ops = 2.0 * 128 * 256 * 32 * KSTEPS * (double)iters * sms;
    printf("{\"tops\": %.1f, \"ms\": %.3f, \"sms\": %d, \"tile\": \"M128 N256 K%d\", \"iters_per_cta\": %d, \"gpu\": \"%s\", "
           "\"how\": \"bare tcgen05.mma.cta_group::1.kind::i8 loop, one CTA per SM, operands resident in shared memory, best of 5 (tools/i8_peak.cu)\"}\n",
           ops / (best * 1e-3) / 1e12, best, sms, 32 * KSTEPS, iters, prop.name);
    return 0;
}
