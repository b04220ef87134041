// brute_search -- ground-truth generator on the GPU index, same record files and output as the
// reference CLI (brute_force_search/src/brute_force.cpp) but with paths on the command line:
//   brute_search <db.bin> <querys.bin> <index.bin> <gt.txt> [topK=100] [dim=128]
// File formats: int32 num; num x { int32 idLen; char id[idLen]; int32 dim; float feat[dim] };
// gt line: "<qid> topK: <id_1> ... dists: <ip_1> ... \n" ascending distance, printed value = 1 - dist.
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "b200nn/hnswlib_gpu.hpp"

using namespace hnswlib;

static bool read_records(const std::string& path, int dim, std::vector<std::string>& ids, std::vector<float>& feats) {
    std::ifstream fp(path.c_str(), std::ios::in | std::ios::binary);
    if (!fp) return false;
    int num = 0;
    fp.read((char*)&num, sizeof(int));
    ids.reserve(num);
    feats.reserve((size_t)num * dim);
    std::vector<float> row(dim);
    for (int i = 0; i < num; i++) {
        int id_len = 0, d = 0;
        fp.read((char*)&id_len, sizeof(int));
        std::string id(id_len, '\0');
        fp.read(&id[0], id_len);
        fp.read((char*)&d, sizeof(int));
        if (d != dim) {
            std::cout << "file error";
            exit(1);
        }
        fp.read((char*)row.data(), sizeof(float) * d);
        ids.push_back(std::string(id.c_str()));  // the reference builds std::string from a C string
        feats.insert(feats.end(), row.begin(), row.end());
    }
    return (bool)fp;
}

int main(int argc, const char* argv[]) {
    if (argc < 5) {
        std::cerr << "usage: brute_search db.bin querys.bin index.bin gt.txt [topK] [dim]\n";
        return 2;
    }
    const int topK = argc > 5 ? atoi(argv[5]) : 100, dim = argc > 6 ? atoi(argv[6]) : 128;
    std::vector<std::string> db_ids, q_ids;
    std::vector<float> db, q;
    if (!read_records(argv[1], dim, db_ids, db) || !read_records(argv[2], dim, q_ids, q)) {
        std::cerr << "cannot read input records\n";
        return 1;
    }
    InnerProductSpace ipspace(dim);
    BruteforceSearch<float> alg(&ipspace, db_ids.size());
    for (size_t i = 0; i < db_ids.size(); i++) alg.addPoint((void*)(db.data() + i * dim), (labeltype)i);
    alg.saveIndex(argv[3]);
    printf("brute force index finished building\n");
    // all queries in one batch (the reference loops searchKnn per query)
    std::vector<std::priority_queue<std::pair<float, labeltype> > > res = alg.searchKnnBatch(q.data(), q_ids.size(), topK);
    std::ofstream gt(argv[4]);
    for (size_t i = 0; i < q_ids.size(); i++) {
        std::vector<std::pair<std::string, float> > rev;
        while (!res[i].empty()) {
            rev.push_back(std::make_pair(db_ids[res[i].top().second], 1.0 - res[i].top().first));
            res[i].pop();
        }
        std::stringstream ids, dists;
        for (int j = (int)rev.size() - 1; j >= 0; j--) {
            ids << rev[j].first << " ";
            dists << rev[j].second << " ";
        }
        gt << q_ids[i] << " topK: " << ids.str() << "dists: " << dists.str() << "\n";
    }
    return 0;
}
