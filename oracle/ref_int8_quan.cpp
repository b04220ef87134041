// ref_int8_quan.cpp -- TEST INFRASTRUCTURE: driver around the UNMODIFIED reference class cvtk::quant::Int8Quan
// (scalar_quantization/scalar_quantization/int8_quan.{h,cc}, compiled in place from /root/reference against the
// stub faiss / nlohmann headers under oracle/stubs; see oracle/Makefile).  It feeds rows through the reference's
// own methods and dumps what they return, so that the restatement (cvt_oracle.c) and the CUDA path can be pinned
// on outputs of the reference itself:
//   Int8Encode (int8_quan.cc:72-94)            -> codes + the row as the call leaves it (normalised in place)
//   L2NormalizeVector (:46-56, private)        -> normalised row
//   Int8Decode(std::string&, float*) (:117-132)-> the reference's double-precision decode
//   Int8EncodeFaiss / Int8Decode(uint8_t*)     -> through the STUB codec (faiss itself is absent: unpinned)
//
//   ref_int8_quan <model | conf.json> <num_source (0 = single-model ctor)> <source> <rows.f32> <n> <d> <l2norm 0|1> <out.bin>
// out.bin: int32 status, n, d | u8 codes[n][d] | f32 x_after[n][d] | f32 normed[n][d] | f32 decode[n][d] |
//          u8 codes_faiss[n][d] | f32 decode_faiss[n][d] | int32 rc_bad_dims
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include <nlohmann/json.hpp>
#include "third_party/faiss/include/Index.h"
#include "third_party/faiss/include/IndexScalarQuantizer.h"
#include "third_party/faiss/include/index_io.h"
#define private public  // L2NormalizeVector is private; every std header int8_quan.h pulls in is already included
#include "int8_quan.h"
#undef private

int main(int argc, char** argv) {
    if (argc != 9) { fprintf(stderr, "usage: see the header of ref_int8_quan.cpp\n"); return 2; }
    const std::string model = argv[1];
    const int num_source = atoi(argv[2]), source = atoi(argv[3]);
    const size_t n = (size_t)atoll(argv[5]), d = (size_t)atoll(argv[6]);
    const bool l2norm = atoi(argv[7]) != 0;
    std::vector<float> rows(n * d);
    {
        FILE* f = fopen(argv[4], "rb");
        if (!f || fread(rows.data(), 4, n * d, f) != n * d) { fprintf(stderr, "cannot read rows\n"); return 2; }
        fclose(f);
    }
    cvtk::quant::Int8Quan* q = num_source > 0 ? new cvtk::quant::Int8Quan(model, num_source) : new cvtk::quant::Int8Quan(model);
    int32_t status = q->status() ? 1 : 0;
    std::vector<uint8_t> codes(n * d), codes_f(n * d);
    std::vector<float> x_after(n * d), normed(n * d), dec(n * d), dec_f(n * d);
    int32_t rc_bad = -1;
    if (status) {
        for (size_t r = 0; r < n; r++) {
            std::copy(rows.begin() + r * d, rows.begin() + (r + 1) * d, x_after.begin() + r * d);
            if (q->Int8Encode(&x_after[r * d], &codes[r * d], d, !l2norm, source) != 1) status = -1;
            std::copy(rows.begin() + r * d, rows.begin() + (r + 1) * d, normed.begin() + r * d);
            q->L2NormalizeVector(&normed[r * d], (int)d);
            std::string emb(codes.begin() + r * d, codes.begin() + (r + 1) * d);
            if (q->Int8Decode(emb, &dec[r * d], source) != 1) status = -2;
        }
        std::vector<float> tmp(rows);
        if (q->Int8EncodeFaiss(tmp.data(), codes_f.data(), n * d, !l2norm, source) != 1) status = -3;
        if (q->Int8Decode(codes.data(), dec_f.data(), n * d, source) != 1) status = -4;
        std::vector<float> bad(d + 1, 0.5f);
        std::vector<uint8_t> badc(d + 1);
        rc_bad = q->Int8Encode(bad.data(), badc.data(), d + 1, true, source);  // n_dims not a multiple of code_size -> 0
    }
    FILE* o = fopen(argv[8], "wb");
    if (!o) return 2;
    const int32_t h[3] = {status, (int32_t)n, (int32_t)d};
    fwrite(h, 4, 3, o);
    fwrite(codes.data(), 1, n * d, o);
    fwrite(x_after.data(), 4, n * d, o);
    fwrite(normed.data(), 4, n * d, o);
    fwrite(dec.data(), 4, n * d, o);
    fwrite(codes_f.data(), 1, n * d, o);
    fwrite(dec_f.data(), 4, n * d, o);
    fwrite(&rc_bad, 4, 1, o);
    fclose(o);
    return 0;
}
