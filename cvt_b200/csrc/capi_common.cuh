// capi_common.cuh -- host-side helpers shared by the C ABI translation units.
#pragma once
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/b200nn.h"
#include "common.cuh"

struct b200nn_ctx {
    b200nn::Ctx c;
    std::mutex mu;
};

namespace b200nn {

// growable device array (capacity doubling; contents preserved)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // make sure at least n elements fit; keep the first `keep` elements
    int reserve(size_t n, size_t keep, cudaStream_t s) {
        if (n <= cap) return 0;
        size_t ncap = cap ? cap : 1024;
        while (ncap < n) ncap *= 2;
        T* np = nullptr;
        B2_CUDA(cudaMalloc(&np, ncap * sizeof(T)));
        if (p && keep) B2_CUDA(cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (p) {
            B2_CUDA(cudaStreamSynchronize(s));
            cudaFree(p);
        }
        p = np;
        cap = ncap;
        return 0;
    }
    // scratch use: exact-ish size, contents not preserved
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        B2_CUDA(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
        return 0;
    }
};

inline int check_dev_err(Ctx* c, const char* what) {
    int h = 0;
    B2_CUDA(cudaMemcpyAsync(&h, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    if (h) {
        cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
        B2_FAIL(-2, std::string(what) + ": device-side configuration error (shared-memory carve-up overflow)");
    }
    return 0;
}

}  // namespace b200nn
