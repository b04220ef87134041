// ref_hnsw_build.cpp -- TEST INFRASTRUCTURE: writes the index file hnsw_sifts_retrieval/makeSearch.cpp loads, with the
// reference's own hnswlib (compiled in place from /root/reference/hnsw_sifts_retrieval/hnswlib; see oracle/Makefile), the way
// makeIdx.cpp:321-364,396 does: HierarchicalNSW<float>(InnerProductSpace(128), max_elements, M = 32, efConstruction = 80),
// addPoint(row, label = running count), saveIndex.
//   ref_hnsw_build <rows.f32> <n> <dim> <M> <efConstruction> <out index.bin>
#include <deque>
#include <fstream>
#include <iostream>
#include <mutex>
#include <queue>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "hnswlib.h"

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage: ref_hnsw_build <rows.f32> <n> <dim> <M> <ef> <out>\n"); return 2; }
    const size_t n = (size_t)atoll(argv[2]), d = (size_t)atoll(argv[3]);
    std::vector<float> x(n * d);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(x.data(), 4, n * d, f) != n * d) { fprintf(stderr, "cannot read rows\n"); return 2; }
    fclose(f);
    hnswlib::InnerProductSpace space(d);
    hnswlib::HierarchicalNSW<float> alg(&space, n, (size_t)atoll(argv[4]), (size_t)atoll(argv[5]));
    for (size_t i = 0; i < n; i++) alg.addPoint((void*)(x.data() + i * d), i);
    alg.saveIndex(argv[6]);
    return 0;
}
