// ivfopq_gpu.hpp -- header-only C++ shim: class IVFOPQ with the reference's public interface
// (opq/src/IVFOPQ.h:31-47) over the b200nn C ABI, so that opq/src/multi_frame_index_test.cpp
// compiles against the GPU index by swapping the include.
//
// Differences from the reference, all documented defects:
//   * Query works right after IndexDatabase (the reference dereferences a NULL m_ivfSize, D-2);
//   * LoadIndex reads what SaveIndex writes (the reference's LoadIndex cannot, D-2); as in the
//     reference's own main, call LoadModel first (the index file carries no reorder table);
//   * no debug printing of every code / LUT (D-5); M <= 32 instead of 16 (D-3).
// Same: one videoId per feature FILE in IndexDatabase (IVFOPQ.cpp:198-201), `threhold = 1.0` clamp
// (IVFOPQ.cpp:5), LoadModel returns 0/1 and prints the reference's message on failure.
#pragma once
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../b200nn.h"
#include "hnswlib_gpu.hpp"  // b200nn::default_ctx / check

typedef unsigned char uchar;
const int max_path = 260;
struct ImgNameStruct {
    char ptr[max_path];
};

// opq/src/common.h:62-72,46-60 -- the 2-D array helpers the reference's callers use with LoadSingleFeatFile
template <typename T>
void Init2DArray(T**& f, int row, int col) {
    T* block = new T[(size_t)row * col]();
    f = new T*[row];
    for (int i = 0; i < row; i++) f[i] = block + (size_t)i * col;
}
template <typename T>
void Delete2DArray(T**& f) {
    if (f != NULL) {
        delete[] f[0];
        delete[] f;
        f = NULL;
    }
}

// opq/src/common.h:25-37: the results_per_query smallest (score, id) pairs, ascending
inline std::vector<std::pair<float, unsigned> > get_sort_results(const std::vector<float>& match_score, int results_per_query) {
    std::vector<std::pair<float, unsigned> > all(match_score.size()), top(results_per_query);
    for (size_t i = 0; i < match_score.size(); i++) all[i] = std::make_pair(match_score[i], (unsigned)i);
    std::partial_sort_copy(all.begin(), all.end(), top.begin(), top.end());
    return top;
}

class IVFOPQ {
public:
    // $B200NN_DEVICES="0,1,..." with more than one entry: the index is row-sharded over those GPUs (b200nn_mpq_*),
    // same answers as on one GPU; otherwise the single default context (b200nn_pq_*).
    IVFOPQ() : m_imgLocation(NULL), m_maxIndexNum(1 << 30), devices_(b200nn::env_devices()) {}
    explicit IVFOPQ(int maxIndexNum) : m_imgLocation(NULL), m_maxIndexNum(maxIndexNum), devices_(b200nn::env_devices()) {}
    ~IVFOPQ() {
        drop();
        delete[] m_imgLocation;
    }
    IVFOPQ(const IVFOPQ&) = delete;
    IVFOPQ& operator=(const IVFOPQ&) = delete;

    // IVFOPQ.cpp:64-102
    int LoadModel(std::string modelFile) {
        printf("load training file....\n");
        std::ifstream fin(modelFile.c_str(), std::ios::binary);
        if (!fin.is_open()) {
            printf("Can not open the model file!\n");
            return 0;
        }
        int hdr[4];
        fin.read((char*)hdr, sizeof hdr);
        D_ = hdr[0]; K_ = hdr[1]; M_ = hdr[2]; ksub_ = hdr[3];
        fin.seekg((std::streamoff)16 + 4LL * K_ * D_ + 4LL * M_ * ksub_ * (D_ / M_));
        perm_.resize(D_);
        fin.read((char*)perm_.data(), 4LL * D_);
        if (!fin) { printf("Can not open the model file!\n"); return 0; }
        drop();
        const int rc = multi() ? b200nn_mpq_load_model(devices_.data(), (int)devices_.size(), modelFile.c_str(), &mh_)
                               : b200nn_pq_load_model(b200nn::default_ctx(), modelFile.c_str(), &h_);
        if (rc != 0) {
            printf("%s\n", b200nn_last_error());
            return 0;
        }
        printf("feature dimension: %d, coarse codebook size: %d, subspace dimension: %d, number of subspace codebook: %d, finetune codebook size: %d\n",
               D_, K_, D_ / M_, M_, ksub_);
        return 1;
    }

    // IVFOPQ.cpp:441-462: raw float32 [n, D] file, every row reordered
    void LoadSingleFeatFile(std::string srcFile, float**& m_ppFeat, int& m_frameNum) {
        std::vector<float> raw;
        if (!read_raw(srcFile, raw, m_frameNum)) return;
        Init2DArray(m_ppFeat, m_frameNum, D_);
        if (m_frameNum > 0)
            b200nn::check(multi() ? b200nn_mpq_rotate(mh_, raw.data(), (size_t)m_frameNum, m_ppFeat[0])
                                  : b200nn_pq_rotate(h_, raw.data(), (size_t)m_frameNum, m_ppFeat[0]));
    }

    // IVFOPQ.cpp:105-174: rows are already reordered; every row gets the current videoId (m_imgNum)
    void Add(float** m_ppFeat, const int m_frameNum) {
        if (m_frameNum <= 0) return;
        std::vector<int32_t> gid(m_frameNum, (int32_t)m_imgNum_);
        b200nn::check(multi() ? b200nn_mpq_add_rotated(mh_, m_ppFeat[0], (size_t)m_frameNum, gid.data())
                              : b200nn_pq_add_rotated(h_, m_ppFeat[0], (size_t)m_frameNum, gid.data()));
    }

    // IVFOPQ.cpp:176-211
    void IndexDatabase(std::vector<std::string> featFiles) {
        const int num = std::min((int)featFiles.size(), m_maxIndexNum);
        delete[] m_imgLocation;
        m_imgLocation = new ImgNameStruct[num > 0 ? num : 1];
        for (int i = 0; i < num; i++) {
            std::cout << featFiles.at(i) << std::endl;
            float** feat = NULL;
            int n = 0;
            LoadSingleFeatFile(featFiles.at(i), feat, n);
            if (n > 0) {
                Add(feat, n);
                memset(m_imgLocation[m_imgNum_].ptr, 0, max_path);
                strncpy(m_imgLocation[m_imgNum_].ptr, featFiles.at(i).c_str(), max_path - 1);
                m_imgNum_++;
            }
            Delete2DArray(feat);
        }
    }

    // IVFOPQ.cpp:213-320 / 322-422 (identical maths; Query additionally prints its LUTs in the reference)
    void Query(std::string featFile, std::vector<std::vector<float> >& matchScore, int nk = 3) { QueryThrehold(featFile, matchScore, nk); }
    void QueryThrehold(std::string featFile, std::vector<std::vector<float> >& matchScore, int nk = 3) {
        std::vector<float> raw;
        int n = 0;
        if (!read_raw(featFile, raw, n) || n == 0) return;
        uint64_t ng = 0;
        b200nn::check(multi() ? b200nn_mpq_info(mh_, NULL, NULL, NULL, NULL, NULL, &ng, NULL, NULL)
                              : b200nn_pq_info(h_, NULL, NULL, NULL, NULL, NULL, &ng));
        const uint64_t ng_lib = ng;
        ng = std::max<uint64_t>(ng, (uint64_t)m_imgNum_);
        std::vector<float> flat((size_t)n * ng, 1.0f);
        if (ng_lib == ng) {
            if (ng) b200nn::check(multi() ? b200nn_mpq_scores(mh_, raw.data(), (size_t)n, nk, flat.data())
                                          : b200nn_pq_scores(h_, raw.data(), (size_t)n, nk, flat.data()));
        } else if (ng_lib) {  // the path table is longer than the highest videoId stored: the extra columns keep the clamp value
            std::vector<float> part((size_t)n * ng_lib);
            b200nn::check(multi() ? b200nn_mpq_scores(mh_, raw.data(), (size_t)n, nk, part.data())
                                  : b200nn_pq_scores(h_, raw.data(), (size_t)n, nk, part.data()));
            for (int f = 0; f < n; f++) std::copy(part.begin() + (size_t)f * ng_lib, part.begin() + (size_t)(f + 1) * ng_lib, flat.begin() + (size_t)f * ng);
        }
        matchScore.resize(n);
        for (int f = 0; f < n; f++) matchScore[f].assign(flat.begin() + (size_t)f * ng, flat.begin() + (size_t)(f + 1) * ng);
    }

    // IVFOPQ.cpp:516-583 byte format; file name composed as the reference does
    void SaveIndex(std::string desDir) {
        std::cout << "save index file..." << std::endl;
        std::vector<const char*> paths(m_imgNum_);
        for (int i = 0; i < m_imgNum_; i++) paths[i] = m_imgLocation[i].ptr;
        // Add() called directly tags rows with videoId = m_imgNum without advancing it (IVFOPQ.cpp:132), so the index may
        // hold one group more than there are paths: the library is told how many entries the table has
        b200nn::check(multi() ? b200nn_mpq_save_index(mh_, desDir.c_str(), paths.data(), paths.size())
                              : b200nn_pq_save_index_n(h_, desDir.c_str(), paths.data(), paths.size()));
    }

    // reads SaveIndex's output (see header note); LoadModel must have been called (for the reorder table)
    void LoadIndex(std::string srcFile) {
        std::cout << "load index..." << std::endl;
        std::ifstream fin(srcFile.c_str(), std::ios::binary);
        if (!fin.is_open()) {
            std::cout << "Can not open the index file." << std::endl;
            exit(0);  // the reference's behaviour, IVFOPQ.cpp:468-472
        }
        int hdr[5];
        fin.read((char*)hdr, sizeof hdr);
        if (multi()) {
            b200nn_mpq_t nh = NULL;
            b200nn::check(b200nn_mpq_load_index(devices_.data(), (int)devices_.size(), srcFile.c_str(), perm_.empty() ? NULL : perm_.data(), 1.0f, &nh));
            drop();
            mh_ = nh;
        } else {
            b200nn_pq_t nh = NULL;
            b200nn::check(b200nn_pq_load_index(b200nn::default_ctx(), srcFile.c_str(), perm_.empty() ? NULL : perm_.data(), 1.0f, &nh));
            drop();
            h_ = nh;
        }
        D_ = hdr[0]; K_ = hdr[1]; M_ = hdr[2]; ksub_ = hdr[3];
        m_imgNum_ = hdr[4];
        // trailing path table: imgNum x char[260]
        delete[] m_imgLocation;
        m_imgLocation = new ImgNameStruct[m_imgNum_ > 0 ? m_imgNum_ : 1];
        fin.seekg(0, std::ios::end);
        const std::streamoff end = fin.tellg();
        fin.seekg(end - (std::streamoff)m_imgNum_ * max_path);
        for (int i = 0; i < m_imgNum_; i++) fin.read(m_imgLocation[i].ptr, max_path);
    }

    ImgNameStruct* m_imgLocation;  // public in the reference (IVFOPQ.h:45)

    // extensions
    b200nn_pq_t handle() const { return h_; }
    b200nn_mpq_t multi_handle() const { return mh_; }
    int imgNum() const { return m_imgNum_; }

private:
    bool read_raw(const std::string& path, std::vector<float>& raw, int& n) {
        std::ifstream fin(path.c_str(), std::ios::binary);
        if (!fin.is_open()) {
            std::cout << "Error open the feat file: " << path << std::endl;  // IVFOPQ.cpp:446
            n = 0;
            return false;
        }
        fin.seekg(0, std::ios::end);
        const long long bytes = (long long)fin.tellg();
        fin.seekg(0, std::ios::beg);
        n = (int)(bytes / (4LL * D_));
        raw.resize((size_t)n * D_);
        fin.read((char*)raw.data(), 4LL * n * D_);
        return true;
    }
    bool multi() const { return devices_.size() > 1; }
    void drop() {
        if (h_) b200nn_pq_destroy(h_);
        if (mh_) b200nn_mpq_destroy(mh_);
        h_ = NULL;
        mh_ = NULL;
    }
    b200nn_pq_t h_ = NULL;
    b200nn_mpq_t mh_ = NULL;
    std::vector<int> devices_;
    int D_ = 0, K_ = 0, M_ = 0, ksub_ = 0, m_imgNum_ = 0, m_maxIndexNum;
    std::vector<int32_t> perm_;
};
