// train_pq_gpu.hpp -- header-only C++ shim: class TrainPQ with the reference's public interface
// (opq/train_codebook/train_PQ_codebook.h:23-54) over the b200nn C ABI, so that
// opq/train_codebook/train_PQ.cpp compiles against the GPU trainer by swapping the include.
//
// The reference calls yael's kmeans (train_PQ_codebook.cpp:164,229; un-vendored, random initialisation).  Here
// both stages run b200nn_kmeans -- a deterministic Lloyd iteration on the device (include/b200nn.h) -- so the
// centroids differ from any yael run but are reproducible; quality is judged by the quantisation error.
//   CoarseQuan == b200nn_kmeans(rows, K, seed)            + the residue of every row (fp32 subtraction)
//   ProdQuan   == b200nn_kmeans(sub-space m, pq_k, seed + 1 + m) for every m
// which is exactly what b200nn_pq_train computes in one call (tests compare the two).
// Differences from the reference, all documented defects:
//   * SaveCodebook writes the reorder tail as int32[D], the layout IVFOPQ::LoadModel reads; the reference writes
//     sizeof(int)*D bytes of a `long int` array (train_PQ_codebook.cpp:286);
//   * `maxTrainFeatNum` keeps its meaning (min with the rows in the file, :63) -- note the reference default 0
//     therefore trains on nothing.
// Knobs the reference does not have: public `seed` and `max_iter` (0 = until convergence, as niter = 0 there),
// also settable through $B200NN_KMEANS_SEED / $B200NN_KMEANS_MAX_ITER for the unmodified main.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../b200nn.h"
#include "hnswlib_gpu.hpp"  // b200nn::default_ctx / check

class TrainPQ {
public:
    TrainPQ(std::string modelFile, int maxTrainFeatNum = 0, int featDim = 128, int coarseK = 8192, int pq_k = 256, int pq_m = 16)
        : seed(0x5EED0003ull), max_iter(0), m_maxTrainFeatNum(maxTrainFeatNum), m_featNum(0), m_featDim(featDim), m_coarseK(coarseK),
          m_pq_k(pq_k), m_pq_m(pq_m), m_pq_step(featDim / pq_m) {
        if (const char* e = getenv("B200NN_KMEANS_SEED")) seed = strtoull(e, NULL, 0);
        if (const char* e = getenv("B200NN_KMEANS_MAX_ITER")) max_iter = atoi(e);
        // the reorder file is a raw `long int[featDim]` (train_PQ_codebook.cpp:13-22)
        printf("feature dimension: %d\n", featDim);
        std::vector<long int> r(m_featDim, 0);
        std::ifstream fin(modelFile.c_str(), std::ios::binary);
        fin.read((char*)r.data(), sizeof(long int) * m_featDim);
        reorder_.assign(r.begin(), r.end());
        for (int i = 0; i < m_featDim; i++) std::cout << reorder_[i] << " ";
        std::cout << std::endl;
    }

    // train_PQ_codebook.cpp:44-124: raw float32 rows, at most maxTrainFeatNum of them, reordered on load
    void LoadFeatureSample(std::string srcDir) {
        m_srcDir = srcDir;
        FILE* f = fopen(srcDir.c_str(), "rb");
        if (f == NULL) {
            std::cout << "Fail to open the source file " << srcDir << std::endl;
            return;
        }
        std::cout << "Load training features, please wait...\n";
        fseek(f, 0, SEEK_END);
        const long long len = ftell(f);
        fseek(f, 0, SEEK_SET);
        const int num = (int)(len / ((long long)m_featDim * sizeof(float)));
        m_featNum = std::min(num, m_maxTrainFeatNum);
        std::vector<float> raw((size_t)std::max(m_featNum, 0) * m_featDim);
        const size_t got = fread(raw.data(), sizeof(float), raw.size(), f);
        fclose(f);
        if (got != raw.size()) m_featNum = (int)(got / m_featDim);
        m_feat.assign((size_t)m_featNum * m_featDim, 0.0f);
        for (int m = 0; m < m_featNum; m++)
            for (int n = 0; n < m_featDim; n++) m_feat[(size_t)m * m_featDim + n] = raw[(size_t)m * m_featDim + reorder_[n]];
        std::cout << m_featNum << " training features loaded!\n";
    }

    void reorder(float* feat) {  // train_PQ_codebook.cpp:126-141
        if (feat == NULL) return;
        std::vector<float> t(m_featDim);
        for (int i = 0; i < m_featDim; i++) t[i] = feat[reorder_[i]];
        for (int i = 0; i < m_featDim; i++) feat[i] = t[i];
    }

    void IFVPQ() {
        CoarseQuan();
        ProdQuan();
    }

    void CoarseQuan() {  // train_PQ_codebook.cpp:150-199
        m_coarse.assign((size_t)m_coarseK * m_featDim, 0.0f);
        std::vector<int32_t> assign(m_featNum);
        b200nn::check(b200nn_kmeans(b200nn::default_ctx(), m_feat.data(), (size_t)m_featNum, m_featDim, m_coarseK, max_iter, seed,
                                    m_coarse.data(), assign.data(), NULL, NULL, NULL));
        std::vector<int> size(m_coarseK, 0);
        for (int i = 0; i < m_featNum; i++) size[assign[i]]++;
        for (int i = 0; i < m_coarseK; i++)
            if (size[i] <= 0) std::cout << "warning: empty cluster 1: " << i << std::endl;
        m_residue.assign(m_feat.size(), 0.0f);
        for (int i = 0; i < m_featNum; i++)
            for (int j = 0; j < m_featDim; j++)
                m_residue[(size_t)i * m_featDim + j] = m_feat[(size_t)i * m_featDim + j] - m_coarse[(size_t)assign[i] * m_featDim + j];
        std::cout << "finish coarse quantization." << std::endl;
    }

    void ProdQuan() {  // train_PQ_codebook.cpp:201-244
        std::cout << "product quantization......." << std::endl;
        m_codebooks.assign((size_t)m_pq_m * m_pq_k * m_pq_step, 0.0f);
        std::vector<float> v((size_t)m_featNum * m_pq_step);
        for (int i = 0; i < m_pq_m; i++) {
            for (int j = 0; j < m_featNum; j++)
                memcpy(v.data() + (size_t)j * m_pq_step, m_residue.data() + (size_t)j * m_featDim + (size_t)i * m_pq_step, sizeof(float) * m_pq_step);
            b200nn::check(b200nn_kmeans(b200nn::default_ctx(), v.data(), (size_t)m_featNum, m_pq_step, m_pq_k, max_iter, seed + 1 + (uint64_t)i,
                                        m_codebooks.data() + (size_t)i * m_pq_k * m_pq_step, NULL, NULL, NULL, NULL));
        }
    }

    void SaveCodebook(std::string desDir) {  // train_PQ_codebook.cpp:247-288 (file name scheme :272)
        std::ostringstream name;
        name << desDir << "/OPQ_db_" << m_featNum << "_dim_" << m_featDim << "_k_" << m_coarseK << "_PQ_m" << m_pq_m << "_k" << m_pq_k << ".fvecs";
        m_desDir = name.str();
        b200nn::check(b200nn_pq_write_model(m_desDir.c_str(), m_featDim, m_coarseK, m_pq_m, m_pq_k, m_coarse.data(), m_codebooks.data(),
                                            reorder_.data()));
    }

    const std::string& saved_path() const { return m_desDir; }
    uint64_t seed;
    int max_iter;

private:
    std::string m_srcDir, m_desDir;
    int m_maxTrainFeatNum, m_featNum, m_featDim, m_coarseK, m_pq_k, m_pq_m, m_pq_step;
    std::vector<float> m_feat, m_residue, m_coarse, m_codebooks;
    std::vector<int32_t> reorder_;
};
