// front_kernels.cuh -- front-end kernels (definitions in front_kernels.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

// rootSIFT in place over n rows of d floats (siftsIndex.cpp:54-71)
int launch_rootsift(Ctx* ctx, float* x, long long n, int d, float eps);

}  // namespace b200nn
