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

}  // namespace b200nn
