// int8_quan_gpu.hpp -- header-only C++ shim: cvtk::quant::Int8Quan with the reference's public
// interface (scalar_quantization/scalar_quantization/int8_quan.h:17-37) over the b200nn C ABI.
//
// Model loading: the reference reads a faiss 1.5.3 IndexScalarQuantizer file through
// faiss::read_index (int8_quan.cc:14).  faiss is not a dependency here; the reader below parses the
// "IxSQ" layout documented in SURVEY.md App. A-8 (written from the published faiss 1.5.x
// index_io source, NOT verified against a real file -- none is shipped with the reference) and
// only needs trained[] = [vmin | vdiff].  Int8Quan(vmin, vdiff) builds the quantizer directly.
#pragma once
#include <stdint.h>
#include <string.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../b200nn.h"
#include "hnswlib_gpu.hpp"  // b200nn::default_ctx / check

namespace cvtk {
namespace quant {

class Int8Quan {
public:
    explicit Int8Quan(const std::string& model_path) { add_model(model_path); }
    // {"0": {"model_path": "..."}, "1": {...}} (int8_quan.cc:20-39); a minimal scanner, no JSON library
    explicit Int8Quan(const std::string& model_conf_path, int num_source) {
        std::ifstream fin(model_conf_path.c_str());
        if (!fin.good()) {
            std::cout << "model file is not exists" << std::endl;
            load_model_ok = false;
            return;
        }
        std::string txt((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
        size_t pos = 0;
        for (int i = 0; i < num_source || num_source <= 0; i++) {
            pos = txt.find("\"model_path\"", pos);
            if (pos == std::string::npos) break;
            const size_t a = txt.find('"', txt.find(':', pos) + 1), b = txt.find('"', a + 1);
            if (a == std::string::npos || b == std::string::npos) break;
            const std::string path = txt.substr(a + 1, b - a - 1);
            std::cout << "load model: " << path << std::endl;
            add_model(path);
            pos = b;
        }
    }
    // direct construction from the trained range (extension)
    Int8Quan(const std::vector<float>& vmin, const std::vector<float>& vdiff) { add_range(vmin, vdiff); }
    ~Int8Quan() {
        for (size_t i = 0; i < sq_.size(); i++) b200nn_sq_destroy(sq_[i]);
    }
    Int8Quan(const Int8Quan&) = delete;
    Int8Quan& operator=(const Int8Quan&) = delete;

    bool status() { return load_model_ok; }

    // int8_quan.cc:72-94: one vector (the first code_size dims), x normalised in place unless turned off
    int Int8Encode(float* x, uint8_t* bytes, size_t n_dims, bool turn_off_l2norm = false, int source = 0) {
        if (n_dims % d_[source] != 0) return 0;
        b200nn::check(b200nn_sq_encode(sq_[source], x, 1, turn_off_l2norm ? 0 : 1, bytes));
        return 1;
    }
    // int8_quan.cc:58-70: n = n_dims / code_size vectors
    int Int8EncodeFaiss(float* x, uint8_t* bytes, size_t n_dims, bool turn_off_l2norm = false, int source = 0) {
        if (n_dims % d_[source] != 0) return 0;
        b200nn::check(b200nn_sq_encode(sq_[source], x, n_dims / d_[source], turn_off_l2norm ? 0 : 1, bytes));
        return 1;
    }
    // int8_quan.cc:117-132: the reference's double-precision decode of ONE vector
    int Int8Decode(std::string& embedding, float* x, int source = 0) {
        if (embedding.empty()) return 0;
        if (embedding.size() % d_[source] != 0) return 0;
        b200nn::check(b200nn_sq_decode(sq_[source], (const uint8_t*)embedding.data(), 1, 0, x));
        return 1;
    }
    // int8_quan.cc:96-115: faiss's all-float decode of n vectors
    int Int8Decode(uint8_t* bytes, float* x, size_t n_dims, int source = 0) {
        if (n_dims % d_[source] != 0) return 0;
        b200nn::check(b200nn_sq_decode(sq_[source], bytes, n_dims / d_[source], 1, x));
        return 1;
    }
    int Int8DecodeFaiss(std::string& embedding, float* x, int source = 0) {
        return Int8Decode((uint8_t*)embedding.data(), x, embedding.size(), source);
    }

private:
    void add_range(const std::vector<float>& vmin, const std::vector<float>& vdiff) {
        b200nn_sq_t h = NULL;
        b200nn::check(b200nn_sq_create(b200nn::default_ctx(), (int)vmin.size(), vmin.data(), vdiff.data(), &h));
        sq_.push_back(h);
        d_.push_back(vmin.size());
    }
    template <typename T>
    static bool rd(std::ifstream& f, T& v) { return (bool)f.read((char*)&v, sizeof(T)); }
    void add_model(const std::string& path) {
        std::ifstream fin(path.c_str(), std::ios::binary);
        if (!fin.good()) {
            std::cout << "model file is not exists" << std::endl;  // int8_quan.cc:9
            load_model_ok = false;
            return;
        }
        char fourcc[4];
        int32_t d, metric, qtype, rangestat;
        int64_t ntotal, dummy;
        uint8_t trained;
        float rs_arg;
        uint64_t d64, code_size, nt;
        fin.read(fourcc, 4);
        bool ok = memcmp(fourcc, "IxSQ", 4) == 0 && rd(fin, d) && rd(fin, ntotal) && rd(fin, dummy) && rd(fin, dummy) && rd(fin, trained) &&
                  rd(fin, metric) && rd(fin, qtype) && rd(fin, rangestat) && rd(fin, rs_arg) && rd(fin, d64) && rd(fin, code_size) && rd(fin, nt);
        if (!ok || qtype != 0 /*QT_8bit*/ || nt != 2 * (uint64_t)d) {
            std::cout << "unsupported scalar-quantizer model file: " << path << std::endl;
            load_model_ok = false;
            return;
        }
        std::vector<float> tr(nt);
        fin.read((char*)tr.data(), 4 * nt);
        if (!fin) { load_model_ok = false; return; }
        add_range(std::vector<float>(tr.begin(), tr.begin() + d), std::vector<float>(tr.begin() + d, tr.end()));
    }
    bool load_model_ok = true;
    std::vector<b200nn_sq_t> sq_;
    std::vector<size_t> d_;
};

}  // namespace quant
}  // namespace cvtk
