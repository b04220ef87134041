// sq_kernels.cu -- cvtk::quant::Int8Quan arithmetic on the GPU
// (scalar_quantization/scalar_quantization/int8_quan.cc).  PARITY UNPINNED at the faiss 1.5.3
// boundary (no trained model is shipped, faiss is un-vendored): these kernels follow the
// reference's own in-tree restatement operation by operation, with explicit IEEE intrinsics so
// that nothing is fused or reordered:
//   L2NormalizeVector  int8_quan.cc:46-56   float product, double accumulate (sequential), double
//                                           sqrt, denominator narrowed to float, float divide
//   Int8Encode         int8_quan.cc:72-94   (x - vmin) / vdiff, clamp [0,1], (int)(255 * xi)
//   Int8Decode         int8_quan.cc:117-132 vmin + vdiff * (b + 0.5) / 255.0 in double, one rounding
//   faiss float decode int8_quan.cc:96-115  vmin + ((b + 0.5f) / 255.0f) * vdiff in float
#include <algorithm>

#include "sq_kernels.cuh"

namespace b200nn {

// One thread per row: the double sum is sequential over i exactly like the reference loop, so the
// norm (and therefore every quotient and every code) is bit-identical.  Rows of neighbouring
// threads are d floats apart; 16-byte loads keep each touched line hot in L1 for the 8 passes.
__global__ void sq_encode_kernel(float* __restrict__ x, long long n, int d, const float* __restrict__ vmin,
                                 const float* __restrict__ vdiff, int l2norm, unsigned char* __restrict__ codes) {
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
        float* v = x + row * d;
        float denorm = 1.0f;
        if (l2norm) {
            double accum = 0.0;
            for (int i = 0; i < d; ++i) {
                const float e = v[i];
                accum = __dadd_rn(accum, (double)__fmul_rn(e, e));
            }
            accum = __dsqrt_rn(accum);
            denorm = (float)(1e-12 < accum ? accum : 1e-12);  // std::max((double)1e-12, accum) narrowed to float
        }
        unsigned char* c = codes + row * d;
        for (int i = 0; i < d; ++i) {
            float e = v[i];
            if (l2norm) {
                e = __fdiv_rn(e, denorm);
                v[i] = e;  // the reference normalises its argument in place (int8_quan.cc:76-78)
            }
            float xi = 0.0f;
            const float vd = __ldg(vdiff + i);
            if (vd != 0.0f) xi = __fdiv_rn(__fsub_rn(e, __ldg(vmin + i)), vd);
            if (xi < 0.0f) xi = 0.0f;
            if (xi > 1.0f) xi = 1.0f;
            c[i] = (unsigned char)__float2int_rz(__fmul_rn(255.0f, xi));
        }
    }
}

__global__ void sq_decode_kernel(const unsigned char* __restrict__ codes, long long n, int d, const float* __restrict__ vmin,
                                 const float* __restrict__ vdiff, int faiss_float_variant, float* __restrict__ x) {
    const long long total = n * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % d);
        const unsigned char b = codes[i];
        if (faiss_float_variant) {
            const float xi = __fdiv_rn(__fadd_rn((float)b, 0.5f), 255.0f);
            x[i] = __fadd_rn(__ldg(vmin + j), __fmul_rn(xi, __ldg(vdiff + j)));
        } else {
            const double t = __ddiv_rn(__dmul_rn((double)__ldg(vdiff + j), __dadd_rn((double)b, 0.5)), 255.0);
            x[i] = __double2float_rn(__dadd_rn((double)__ldg(vmin + j), t));
        }
    }
}

// faiss RS_minmax (rangestat_arg = 0): per-dimension min and max - min (sq_train.cpp:100-132).
__global__ void sq_minmax_init_kernel(uint32_t* scratch, int d) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < d; j += gridDim.x * blockDim.x) {
        scratch[j] = 0xFFFFFFFFu;  // running min (orderable)
        scratch[d + j] = 0u;       // running max (orderable)
    }
}
__global__ void sq_minmax_kernel(const float* __restrict__ x, long long n, int d, long long rows_per_block,
                                 uint32_t* __restrict__ scratch) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        for (long long r = r0; r < r1; r++) {
            const uint32_t o = f32_orderable(x[r * d + j]);
            lo = min(lo, o);
            hi = max(hi, o);
        }
        if (r1 > r0) {
            atomicMin(scratch + j, lo);
            atomicMax(scratch + d + j, hi);
        }
    }
}
__global__ void sq_minmax_final_kernel(const uint32_t* __restrict__ scratch, int d, float* __restrict__ vmin,
                                       float* __restrict__ vdiff) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < d; j += gridDim.x * blockDim.x) {
        const float lo = f32_from_orderable(scratch[j]), hi = f32_from_orderable(scratch[d + j]);
        vmin[j] = lo;
        vdiff[j] = __fsub_rn(hi, lo);
    }
}

int launch_sq_encode(Ctx* ctx, float* x, long long n, int d, const float* vmin, const float* vdiff, int l2norm,
                     unsigned char* codes) {
    if (n <= 0) return 0;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n + 127) / 128, (long long)ctx->sm_count * 16));
    sq_encode_kernel<<<grid, 128, 0, ctx->stream>>>(x, n, d, vmin, vdiff, l2norm, codes);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_sq_decode(Ctx* ctx, const unsigned char* codes, long long n, int d, const float* vmin, const float* vdiff,
                     int faiss_float_variant, float* x) {
    if (n <= 0) return 0;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n * d + 255) / 256, (long long)ctx->sm_count * 16));
    sq_decode_kernel<<<grid, 256, 0, ctx->stream>>>(codes, n, d, vmin, vdiff, faiss_float_variant, x);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

int launch_sq_train_minmax(Ctx* ctx, const float* x, long long n, int d, uint32_t* scratch, float* vmin, float* vdiff) {
    sq_minmax_init_kernel<<<(d + 255) / 256, 256, 0, ctx->stream>>>(scratch, d);
    const long long rpb = 512;
    if (n > 0) sq_minmax_kernel<<<(unsigned)((n + rpb - 1) / rpb), 128, 0, ctx->stream>>>(x, n, d, rpb, scratch);
    sq_minmax_final_kernel<<<(d + 255) / 256, 256, 0, ctx->stream>>>(scratch, d, vmin, vdiff);
    ctx->launches += 3;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
