// capi_pq_internal.cuh -- host-side pieces of the (O)PQ C ABI shared with the multi-GPU layer (capi_multi.cu).
#pragma once
#include <string>
#include <vector>

#include "capi_common.cuh"

struct PQHostRows {  // rows of an index on the host, insertion order: coarse list id, videoId, M code bytes
    std::vector<int> lists, groups;
    std::vector<unsigned char> codes;
};

std::string pq_index_file_name(const char* dir_or_path, long long n_groups, int D, int K, int M, int ksub);
int pq_model_to_host(b200nn_pq_t p, std::vector<float>* coarse, std::vector<float>* cb, std::vector<int>* perm);
int pq_rows_to_host(b200nn_pq_t p, PQHostRows* rows);
int pq_write_index_file(const std::string& path, int D, int K, int M, int ksub, long long n_groups, const float* coarse, const float* cb,
                        long long n, const int* lists, const int* groups, const unsigned char* codes, const char* const* group_paths,
                        size_t n_paths);
int pq_parse_index_file(const char* path, int* D, int* K, int* M, int* ksub, long long* n_groups, std::vector<float>* coarse,
                        std::vector<float>* cb, PQHostRows* rows);
