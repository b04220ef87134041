// front_kernels.cu -- f-3, the steps that PRODUCE the vectors fed to the quantized path:
//   rootSIFT  (siftsIDX::rootSift, hnsw_sifts_retrieval/siftsIndex.cpp:54-71; same code makeSIFTs.cpp:79-95)
//   PCA projection + L2 normalisation (cvtk::PCAUtils::reduceDim, pca_train_project/pca_online/pca_utils.cc:25-35)
//     -> the tcgen05 split-TF32 GEMM of rotate_gemm.cu with the mean subtraction in the loader and the
//        normalisation in the epilogue.
#include <algorithm>

#include "front_kernels.cuh"

namespace b200nn {

// rootSIFT, in place.  d = abs(d); sums = reduce(d, SUM); d = sqrt(d / (sums + eps)); cv::normalize(row, NORM_L2).
// OpenCV accumulates both the row sum and the squared norm in double; one THREAD walks one row in column
// order so that the two double sums are bit-identical to a sequential host loop.  Rows are staged through
// shared memory (row stride d+1 words: conflict-free for thread-per-row access), global accesses stay coalesced.
__global__ void rootsift_kernel(float* __restrict__ x, long long n, int d, float eps) {
    extern __shared__ float s_rows[];  // [blockDim.x][d + 1]
    const int RS = d + 1;
    const long long row0 = (long long)blockIdx.x * blockDim.x;
    const int rows = (int)min((long long)blockDim.x, n - row0);
    for (long long i = threadIdx.x; i < (long long)rows * d; i += blockDim.x) {
        const int r = (int)(i / d), c = (int)(i - (long long)r * d);
        s_rows[r * RS + c] = fabsf(x[row0 * d + i]);
    }
    __syncthreads();
    if ((int)threadIdx.x < rows) {
        float* v = s_rows + threadIdx.x * RS;
        double sum = 0.0;
        for (int j = 0; j < d; j++) sum = __dadd_rn(sum, (double)v[j]);
        const float den = __fadd_rn((float)sum, eps);
        double s2 = 0.0;
        for (int j = 0; j < d; j++) {
            const float t = __fsqrt_rn(__fdiv_rn(v[j], den));
            v[j] = t;
            s2 = __dadd_rn(s2, __dmul_rn((double)t, (double)t));
        }
        const double nrm = __dsqrt_rn(s2);
        const double scale = nrm > 2.220446049250313e-16 ? __ddiv_rn(1.0, nrm) : 0.0;  // cv::normalize's DBL_EPSILON guard
        for (int j = 0; j < d; j++) v[j] = (float)__dmul_rn((double)v[j], scale);
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < (long long)rows * d; i += blockDim.x) {
        const int r = (int)(i / d), c = (int)(i - (long long)r * d);
        x[row0 * d + i] = s_rows[r * RS + c];
    }
}

int launch_rootsift(Ctx* ctx, float* x, long long n, int d, float eps) {
    if (n <= 0) return 0;
    if (d < 1 || d > 1024) B2_FAIL(-4, "rootsift: descriptor length must be in [1, 1024]");
    int threads = 128;
    while (threads > 32 && (size_t)threads * (d + 1) * sizeof(float) > 96 * 1024) threads >>= 1;
    const size_t smem = (size_t)threads * (d + 1) * sizeof(float);
    if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(rootsift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rootsift_kernel<<<(unsigned)((n + threads - 1) / threads), threads, smem, ctx->stream>>>(x, n, d, eps);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200nn
