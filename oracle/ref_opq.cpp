// ref_opq.cpp -- drives the UNMODIFIED reference IVFOPQ (compiled in place from
// /root/reference/opq/src/IVFOPQ.cpp, never copied) so that its outputs can pin the oracle
// restatement (oracle/cvt_oracle.c), generate tests/golden/, and serve as the CPU baseline that
// bench.py times (`--impl reference`, cpu_baseline.kind = "reference").
//
// TEST INFRASTRUCTURE ONLY: the product never links or executes this.
//
// Build (oracle/Makefile):  g++ -O2 -std=c++11 -fopenmp -I$(REF)/opq/src ref_opq.cpp $(REF)/opq/src/IVFOPQ.cpp
//
// The reference cannot run Query as shipped (m_ivfSize is only allocated by the broken LoadIndex,
// SURVEY.md App. D-2), so -- exactly as SURVEY.md App. F documents -- this harness includes the std
// headers first and then opens the class with `#define private public`; the reference sources
// stay untouched.
#include <iostream>
#include <sstream>
#include <fstream>
#include <string>
#include <vector>
#include <queue>
#include <algorithm>
#include <string.h>
#include <sys/time.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <omp.h>
#include "common.h"
#define private public
#include "IVFOPQ.h"
#undef private

static double now_s() {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

// The reference prints every code (Add), the model header (LoadModel) and a statistic
// (QueryThrehold) to stdout; park stdout on /dev/null while it runs.
static int g_saved_stdout = -1;
static void mute() {
    fflush(stdout);
    if (g_saved_stdout < 0) g_saved_stdout = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
}
static void unmute() {
    fflush(stdout);
    std::cout.flush();
    if (g_saved_stdout >= 0) dup2(g_saved_stdout, 1);
}

static void set_ivf_sizes(IVFOPQ& index) {
    // what LoadIndex would have done (IVFOPQ.cpp:493-498)
    if (index.m_ivfSize) delete[] index.m_ivfSize;
    index.m_ivfSize = new int[index.m_coarseK];
    for (int i = 0; i < index.m_coarseK; i++) index.m_ivfSize[i] = (int)index.m_ivfList[i].size();
}

// Adds one feature file.  per_row: every row gets its own videoId (m_imgNum++ per row);
// otherwise the whole file shares one id, as IndexDatabase does (IVFOPQ.cpp:198-201).
// Records (list, group, code) per row by watching which inverted list grew.
static void add_file(IVFOPQ& index, const std::string& path, bool per_row, std::vector<int>& row_list,
                     std::vector<int>& row_group, std::vector<unsigned char>& row_codes) {
    float** feat = NULL;
    int n = 0;
    index.LoadSingleFeatFile(path, feat, n);  // applies reorder() to every row (:459-461)
    if (n == 0) return;
    const int K = index.m_coarseK, M = index.m_pq_m;
    std::vector<size_t> before(K);
    for (int r = 0; r < n; r++) {
        // find the list that grows: only the argmin list changes, so remember sizes lazily
        float* rowp[1] = {feat[r]};
        // cheap trick: total size known, scan lists after Add for the one whose back() is new
        for (int i = 0; i < K; i++) before[i] = index.m_ivfList[i].size();
        index.Add(rowp, 1);
        int vw = -1;
        for (int i = 0; i < K; i++)
            if (index.m_ivfList[i].size() != before[i]) { vw = i; break; }
        row_list.push_back(vw);
        row_group.push_back(index.m_imgNum);
        const IVFelem& e = index.m_ivfList[vw].back();
        for (int m = 0; m < M; m++) row_codes.push_back(e.PQindex[m]);
        if (per_row) index.m_imgNum++;
    }
    if (!per_row) index.m_imgNum++;
    Delete2DArray(feat);
}

template <typename T> static void wr(FILE* f, const T& v) { fwrite(&v, sizeof(T), 1, f); }

// run <model> <out.bin> <nk> <topk> <per_row 0|1> <n_db_files> db... <n_query_files> query...
static int cmd_run(int argc, char** argv) {
    if (argc < 8) return 2;
    std::string model = argv[2], out = argv[3];
    int nk = atoi(argv[4]), topk = atoi(argv[5]);
    bool per_row = atoi(argv[6]) != 0;
    int ndb = atoi(argv[7]);
    std::vector<std::string> db, qf;
    int a = 8;
    for (int i = 0; i < ndb; i++) db.push_back(argv[a++]);
    int nqf = atoi(argv[a++]);
    for (int i = 0; i < nqf; i++) qf.push_back(argv[a++]);

    mute();
    IVFOPQ index(1 << 30);
    if (!index.LoadModel(model)) { unmute(); fprintf(stderr, "cannot load model\n"); return 1; }
    std::vector<int> row_list, row_group;
    std::vector<unsigned char> row_codes;
    for (size_t i = 0; i < db.size(); i++) add_file(index, db[i], per_row, row_list, row_group, row_codes);
    set_ivf_sizes(index);

    // cross-check against the reference's own IndexDatabase when ids are per file
    int consistent = 1;
    if (!per_row) {
        IVFOPQ index2(1 << 30);
        index2.LoadModel(model);
        index2.IndexDatabase(db);
        for (int i = 0; i < index.m_coarseK && consistent; i++) {
            if (index.m_ivfList[i].size() != index2.m_ivfList[i].size()) { consistent = 0; break; }
            for (size_t j = 0; j < index.m_ivfList[i].size(); j++) {
                const IVFelem &x = index.m_ivfList[i][j], &y = index2.m_ivfList[i][j];
                if (x.videoId != y.videoId || memcmp(x.PQindex, y.PQindex, index.m_pq_m)) { consistent = 0; break; }
            }
        }
    }

    std::vector<std::vector<float> > all_scores;  // one row per query frame over all query files
    std::vector<int> frames_per_file;
    for (size_t i = 0; i < qf.size(); i++) {
        std::vector<std::vector<float> > score;
        index.QueryThrehold(qf[i], score, nk);
        frames_per_file.push_back((int)score.size());
        for (size_t f = 0; f < score.size(); f++) all_scores.push_back(score[f]);
    }
    unmute();

    FILE* fo = fopen(out.c_str(), "wb");
    if (!fo) return 1;
    fwrite("ROPQ", 1, 4, fo);
    int n_rows = (int)row_list.size(), n_groups = index.m_imgNum, nq = (int)all_scores.size();
    int kk = std::min(topk, n_groups);
    wr(fo, (int)1); wr(fo, index.m_featDim); wr(fo, index.m_coarseK); wr(fo, index.m_pq_m); wr(fo, index.m_pq_k);
    wr(fo, n_rows); wr(fo, n_groups); wr(fo, nq); wr(fo, nk); wr(fo, kk); wr(fo, consistent);
    wr(fo, (int)qf.size());
    for (size_t i = 0; i < qf.size(); i++) wr(fo, frames_per_file[i]);
    fwrite(row_list.data(), sizeof(int), n_rows, fo);
    fwrite(row_group.data(), sizeof(int), n_rows, fo);
    fwrite(row_codes.data(), 1, row_codes.size(), fo);
    for (int f = 0; f < nq; f++) fwrite(all_scores[f].data(), sizeof(float), n_groups, fo);
    // per-frame top-k through the reference's own get_sort_results (common.h:25-37)
    for (int f = 0; f < nq; f++) {
        std::vector<std::pair<float, uint> > r = get_sort_results(all_scores[f], kk);
        for (int j = 0; j < kk; j++) { wr(fo, r[j].first); wr(fo, (unsigned)r[j].second); }
    }
    // per query FILE: frame-summed scores + get_sort_results, as multi_frame_index_test.cpp:56-68
    int off = 0;
    for (size_t i = 0; i < qf.size(); i++) {
        std::vector<float> total(n_groups, 0.0f);
        for (int j = 0; j < frames_per_file[i]; j++)
            for (int k = 0; k < n_groups; k++) total.at(k) += all_scores[off + j].at(k);
        off += frames_per_file[i];
        std::vector<std::pair<float, uint> > r = get_sort_results(total, kk);
        for (int j = 0; j < kk; j++) { wr(fo, r[j].first); wr(fo, (unsigned)r[j].second); }
    }
    fclose(fo);
    printf("{\"rows\": %d, \"groups\": %d, \"queries\": %d, \"consistent\": %d}\n", n_rows, n_groups, nq, consistent);
    return 0;
}

// bench <model> <db_feat_file> <query_feat_file> <nk> <topk> <n_queries> <threads> <tmpdir> [repeat] [dump.bin]
// Builds the index with the reference's own Add (per-row ids), then times
// QueryThrehold + get_sort_results over n_queries query rows sharded across OpenMP threads
// (the reference itself is single-threaded; one IVFOPQ object is shared read-only).
static int cmd_bench(int argc, char** argv) {
    if (argc < 10) return 2;
    std::string model = argv[2], dbf = argv[3], qfile = argv[4];
    int nk = atoi(argv[5]), topk = atoi(argv[6]), nq = atoi(argv[7]), threads = atoi(argv[8]);
    std::string tmpdir = argv[9];
    int repeat = argc > 10 ? atoi(argv[10]) : 1;
    const char* dump = argc > 11 ? argv[11] : NULL;
    if (threads <= 0) threads = omp_get_max_threads();

    mute();
    IVFOPQ index(1 << 30);
    if (!index.LoadModel(model)) { unmute(); fprintf(stderr, "cannot load model\n"); return 1; }
    double t0 = now_s();
    {
        float** feat = NULL;
        int n = 0;
        index.LoadSingleFeatFile(dbf, feat, n);
        for (int r = 0; r < n; r++) {
            float* rowp[1] = {feat[r]};
            index.Add(rowp, 1);
            index.m_imgNum++;
        }
        Delete2DArray(feat);
    }
    set_ivf_sizes(index);
    double t_build = now_s() - t0;

    // split the first nq query rows into one small feature file per thread (QueryThrehold takes
    // a file path and allocates a dense [frames][m_imgNum] score matrix, so keep files small)
    const int D = index.m_featDim;
    std::vector<float> qraw((size_t)nq * D);
    {
        FILE* f = fopen(qfile.c_str(), "rb");
        if (!f || fread(qraw.data(), sizeof(float), (size_t)nq * D, f) != (size_t)nq * D) {
            unmute(); fprintf(stderr, "cannot read %d query rows\n", nq); return 1;
        }
        fclose(f);
    }
    const int per_file = 4;
    int nfiles = (nq + per_file - 1) / per_file;
    std::vector<std::string> qpaths(nfiles);
    for (int i = 0; i < nfiles; i++) {
        char buf[64];
        snprintf(buf, sizeof buf, "/refq_%d.bin", i);
        qpaths[i] = tmpdir + buf;
        FILE* f = fopen(qpaths[i].c_str(), "wb");
        int lo = i * per_file, hi = std::min(nq, lo + per_file);
        fwrite(qraw.data() + (size_t)lo * D, sizeof(float), (size_t)(hi - lo) * D, f);
        fclose(f);
    }
    std::vector<unsigned> first_ids(nq, 0);
    std::vector<float> first_scores(nq, 0.f);
    std::vector<unsigned> all_ids((size_t)nq * topk, 0);
    std::vector<float> all_scores((size_t)nq * topk, 0.f);
    double best = 1e30, total = 0;
    for (int rep = 0; rep < repeat; rep++) {
        double t1 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int i = 0; i < nfiles; i++) {
            std::vector<std::vector<float> > score;
            index.QueryThrehold(qpaths[i], score, nk);
            for (size_t f = 0; f < score.size(); f++) {
                std::vector<std::pair<float, uint> > r = get_sort_results(score[f], topk);
                first_ids[i * per_file + f] = r[0].second;
                first_scores[i * per_file + f] = r[0].first;
                for (int j = 0; j < topk; j++) {
                    all_ids[(size_t)(i * per_file + f) * topk + j] = r[j].second;
                    all_scores[(size_t)(i * per_file + f) * topk + j] = r[j].first;
                }
            }
        }
        double dt = now_s() - t1;
        total += dt;
        if (dt < best) best = dt;
    }
    for (int i = 0; i < nfiles; i++) unlink(qpaths[i].c_str());
    unmute();
    if (dump) {  // [nq][topk] scores (f32) then [nq][topk] ids (u32), for the parity check next to the timing
        FILE* fd = fopen(dump, "wb");
        if (fd) {
            fwrite(all_scores.data(), sizeof(float), all_scores.size(), fd);
            fwrite(all_ids.data(), sizeof(unsigned), all_ids.size(), fd);
            fclose(fd);
        }
    }
    unsigned long long chk = 0;
    for (int i = 0; i < nq; i++) chk = chk * 1315423911ull + first_ids[i];
    printf("{\"n_rows\": %d, \"n_queries\": %d, \"threads\": %d, \"repeat\": %d, \"build_s\": %.6f, "
           "\"query_s_best\": %.6f, \"query_s_mean\": %.6f, \"qps\": %.6f, \"checksum\": %llu}\n",
           index.m_imgNum, nq, threads, repeat, t_build, best, total / repeat, nq / (total / repeat), chk);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_opq run|bench ...\n"); return 2; }
    std::string c = argv[1];
    if (c == "run") return cmd_run(argc, argv);
    if (c == "bench") return cmd_bench(argc, argv);
    return 2;
}
