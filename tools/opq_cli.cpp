// opq_cli -- the two mains of opq/src/multi_frame_index_test.cpp with command-line paths instead of
// the reference's hard-coded /Users/... constants, written against the drop-in IVFOPQ class:
//   opq_cli index <model> <feat_list.txt> <out_dir>
//   opq_cli query <model> <index.fvecs> <result.txt> <query_feat.bin>... [--nk N] [--show K]
#include <stdlib.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "b200nn/compat/common.h"
#include "b200nn/ivfopq_gpu.hpp"

int main(int argc, char* argv[]) {
    if (argc < 5) {
        std::cerr << "usage: opq_cli index <model> <list.txt> <out_dir> | opq_cli query <model> <index> <result.txt> <q.bin>...\n";
        return 2;
    }
    const std::string mode = argv[1];
    if (mode == "index") {
        std::vector<std::string> feat_files;
        get_vector_of_strings_from_file_lines(argv[3], feat_files);
        std::cout << feat_files.size() << std::endl;
        IVFOPQ index(1000);  // maxImageNum, multi_frame_index_test.cpp:14
        if (!index.LoadModel(argv[2])) return 1;
        index.IndexDatabase(feat_files);
        index.SaveIndex(argv[4]);
        return 0;
    }
    if (mode == "query") {
        int nk = 3, show = 5;  // num_nearest, num_show (multi_frame_index_test.cpp:34-35)
        std::vector<std::string> queries;
        for (int i = 5; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--nk" && i + 1 < argc) nk = atoi(argv[++i]);
            else if (a == "--show" && i + 1 < argc) show = atoi(argv[++i]);
            else queries.push_back(a);
        }
        IVFOPQ search;
        if (!search.LoadModel(argv[2])) return 1;
        search.LoadIndex(argv[3]);
        std::ofstream fout(argv[4]);
        fout.precision(9);
        for (size_t i = 0; i < queries.size(); i++) {
            std::vector<std::vector<float> > score;
            search.Query(queries[i], score, nk);
            const int frames = (int)score.size();
            if (frames > 0) {
                const int imgs = (int)score[0].size();
                std::vector<float> total(imgs, 0.0f);
                for (int j = 0; j < frames; j++)
                    for (int k = 0; k < imgs; k++) total[k] += score[j][k];  // frame-summed, :60-67
                const int kk = std::min(show, imgs);
                std::vector<std::pair<float, unsigned> > result = get_sort_results(total, kk);
                fout << queries[i] << "  " << i << std::endl;
                for (int j = 0; j < kk; j++) fout << get_base_name(search.m_imgLocation[result[j].second].ptr) << " ";
                fout << std::endl;
                for (int j = 0; j < kk; j++) fout << result[j].first << " ";
                fout << "\n\n";
            }
            std::cout << "query ID: " << i << ", frame_num:" << frames << std::endl;
        }
        return 0;
    }
    return 2;
}
