// sq_kernels.cuh -- launchers of the scalar-quantizer kernels (definitions in sq_kernels.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

int launch_sq_encode(Ctx* ctx, float* x, long long n, int d, const float* vmin, const float* vdiff, int l2norm,
                     unsigned char* codes);
int launch_sq_decode(Ctx* ctx, const unsigned char* codes, long long n, int d, const float* vmin, const float* vdiff,
                     int faiss_float_variant, float* x);
// vmin/vdiff are device outputs [d]; scratch = 2*d uint32 device words
int launch_sq_train_minmax(Ctx* ctx, const float* x, long long n, int d, uint32_t* scratch, float* vmin, float* vdiff);

}  // namespace b200nn
