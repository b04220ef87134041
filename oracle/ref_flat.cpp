// ref_flat.cpp -- drives the UNMODIFIED reference hnswlib headers (compiled in place from
// /root/reference, never copied): BruteforceSearch<dist_t>::addPoint/searchKnn/saveIndex with the
// reference's own distance functions.  Pins oracle/cvt_oracle.c's flat_* restatement and
// generates tests/golden/.  TEST INFRASTRUCTURE ONLY.
//
// Two flavours (oracle/Makefile):
//   -DREF_BF   : brute_force_search/src/{brutoforce,space_ip}.hpp  (InnerProductSpace only);
//                built without -mavx  -> SSE 4-lane order (the CLI's own flags, CMakeLists.txt:4)
//                built with    -mavx  -> AVX 8-lane order
//   -DREF_HNSW : hnsw_sifts_retrieval/hnswlib/hnswlib.h umbrella (needs -mavx: space_l2.h:13
//                force-defines USE_AVX): L2Space (AVX order), L2SpaceI (exact int),
//                InnerProductSpace (SSE order; its AVX path is `#if 0`'d, space_ip.h:45-46)
//
// usage: ref_flat <metric ip|l2|l2i> <data.bin> <labels.bin|-> <queries.bin> <n> <d> <nq> <k> <out.bin> [save_index_path]
#include <deque>
#include <mutex>
#include <vector>
#include <iostream>
#include <fstream>
#include <string>
#include <stdexcept>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#if defined(REF_BF)
#include "brutoforce.hpp"
#include "space_ip.hpp"
#elif defined(REF_HNSW)
#include "hnswlib.h"
#else
#error "define REF_BF or REF_HNSW"
#endif

template <typename T> static std::vector<T> slurp(const char* path, size_t count) {
    std::vector<T> v(count);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(T), count, f) != count) { fprintf(stderr, "short read %s\n", path); exit(1); }
    fclose(f);
    return v;
}

template <typename dist_t, typename elem_t>
static int run(hnswlib::SpaceInterface<dist_t>* space, const char* data_p, const char* lab_p, const char* q_p,
               size_t n, size_t d, size_t nq, size_t k, const char* out_p, const char* save_p) {
    std::vector<elem_t> data = slurp<elem_t>(data_p, n * d);
    std::vector<elem_t> q = slurp<elem_t>(q_p, nq * d);
    std::vector<uint64_t> labels(n);
    if (std::string(lab_p) == "-") for (size_t i = 0; i < n; i++) labels[i] = i;
    else labels = slurp<uint64_t>(lab_p, n);
    hnswlib::BruteforceSearch<dist_t>* alg = new hnswlib::BruteforceSearch<dist_t>(space, n);
    for (size_t i = 0; i < n; i++) alg->addPoint((void*)(data.data() + i * d), (hnswlib::labeltype)labels[i]);
    if (save_p) alg->saveIndex(save_p);
    FILE* fo = fopen(out_p, "wb");
    for (size_t i = 0; i < nq; i++) {
        std::priority_queue<std::pair<dist_t, hnswlib::labeltype> > r = alg->searchKnn((void*)(q.data() + i * d), k);
        std::vector<std::pair<dist_t, hnswlib::labeltype> > v;
        while (!r.empty()) { v.push_back(r.top()); r.pop(); }
        for (size_t j = v.size(); j-- > 0;) {  // ascending, as brute_force.cpp:96-101 prints
            dist_t dd = v[j].first;
            uint64_t ll = v[j].second;
            fwrite(&dd, sizeof(dist_t), 1, fo);
            fwrite(&ll, sizeof(uint64_t), 1, fo);
        }
    }
    fclose(fo);
    delete alg;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 10) { fprintf(stderr, "usage: see header\n"); return 2; }
    std::string metric = argv[1];
    size_t n = atoll(argv[5]), d = atoll(argv[6]), nq = atoll(argv[7]), k = atoll(argv[8]);
    const char* save_p = argc > 10 ? argv[10] : NULL;
    if (metric == "ip") {
        hnswlib::InnerProductSpace sp(d);
        return run<float, float>(&sp, argv[2], argv[3], argv[4], n, d, nq, k, argv[9], save_p);
    }
#if defined(REF_HNSW)
    if (metric == "l2") {
        hnswlib::L2Space sp(d);
        return run<float, float>(&sp, argv[2], argv[3], argv[4], n, d, nq, k, argv[9], save_p);
    }
    if (metric == "l2i") {
        hnswlib::L2SpaceI sp(d);
        return run<int, unsigned char>(&sp, argv[2], argv[3], argv[4], n, d, nq, k, argv[9], save_p);
    }
#endif
    fprintf(stderr, "metric %s not available in this flavour\n", metric.c_str());
    return 2;
}
