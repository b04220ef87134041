// rotate_gemm.cuh -- dense rotation on the tcgen05 tensor cores (definitions in rotate_gemm.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace b200nn {

bool rotate_gemm_supported(int D);
// host: split R into two TF32-exact parts arranged for the kernel (2*D*D floats)
void rotate_gemm_pack_R(const float* R, int D, std::vector<float>& out);
// device: y[n,D] = x[n,D] * R^T
int launch_rotate_gemm(Ctx* ctx, const float* x, long long n, int D, const float* bplanes, float* y);

// General projection y[n,N] = (x[n,K] - mean) * V^T on the same kernel (f-3, cvtk::PCAUtils::reduceDim):
// K % 32 == 0, N % 64 == 0; with l2norm the reference's per-row normalisation is fused in the epilogue and
// N must be 64, 128 or 256 (one CTA owns whole output rows).
bool proj_gemm_supported(int K, int N, bool l2norm);
int proj_gemm_nblock(int N, bool l2norm);
void proj_gemm_pack(const float* V, int N, int K, int NB, std::vector<float>& out);  // V [N,K] row-major -> 2*N*K floats
int launch_proj_gemm(Ctx* ctx, const float* x, long long n, int K, int N, const float* mean /*[K] or NULL*/, const float* bplanes,
                     bool l2norm, float* y);

}  // namespace b200nn
