// hnswlib_gpu.hpp -- header-only C++ shim with the reference's hnswlib interface over the b200nn
// C ABI, so that brute_force_search/src/brute_force.cpp and hnsw_sifts_retrieval/makeSearch.cpp
// compile against the GPU index with (at most) an #include change.
//
// What it mirrors (willard-yuan/cvt):
//   brute_force_search/src/hnswlib.hpp:22,35,38-58   labeltype, DISTFUNC, SpaceInterface, AlgorithmInterface
//   brute_force_search/src/space_ip.hpp:211-239      InnerProductSpace
//   hnsw_sifts_retrieval/hnswlib/space_l2.h:153-180  L2Space        :221-245  L2SpaceI
//   brute_force_search/src/brutoforce.hpp:8-136      BruteforceSearch<dist_t>
//
// Semantics kept: addPoint copies the vector, duplicate label / capacity throw std::runtime_error
// with the reference's messages, searchKnn returns a max-heap std::priority_queue whose top() is
// the farthest of the k (lexicographic (dist,label) selection), saveIndex writes the reference's
// byte format.  Added: searchKnnBatch (all queries of an image in one launch).
// The distances are computed on the GPU in the reference's accumulation order; the DISTFUNC a
// space returns is therefore never called by the index and exists for interface compatibility.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <queue>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../b200nn.h"

namespace b200nn {
// one lazily created context per process (device from $B200NN_DEVICE, default 0); the function-local static is
// initialised exactly once even when several host threads make their first call together (C++11)
inline b200nn_ctx_t default_ctx() {
    static b200nn_ctx_t ctx = []() {
        b200nn_ctx_t c = nullptr;
        const char* e = std::getenv("B200NN_DEVICE");
        if (b200nn_ctx_create(e ? std::atoi(e) : 0, &c) != 0) throw std::runtime_error(b200nn_last_error());
        return c;
    }();
    return ctx;
}
inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(b200nn_last_error());
}
// $B200NN_DEVICES = "0,1,2,3": the GPUs a sharded index spreads over (empty / one entry: the single default context)
inline std::vector<int> env_devices() {
    std::vector<int> d;
    const char* e = std::getenv("B200NN_DEVICES");
    if (!e) return d;
    for (const char* p = e; *p;) {
        char* end = nullptr;
        const long v = std::strtol(p, &end, 10);
        if (end == p) break;
        d.push_back((int)v);
        p = (*end == ',') ? end + 1 : end;
    }
    return d;
}
// how a space tells the GPU index what to compute
struct GpuMetric {
    virtual int b200nn_metric() const = 0;  // B200NN_METRIC_*
    virtual int b200nn_order() const = 0;   // 1 scalar, 4 SSE, 8 AVX accumulation order
    virtual size_t b200nn_dim() const = 0;
    virtual ~GpuMetric() {}
};
}  // namespace b200nn

namespace hnswlib {

typedef size_t labeltype;

template <typename MTYPE>
using DISTFUNC = MTYPE (*)(const void*, const void*, const void*);

template <typename MTYPE>
class SpaceInterface {
public:
    virtual size_t get_data_size() = 0;
    virtual DISTFUNC<MTYPE> get_dist_func() = 0;
    virtual void* get_dist_func_param() = 0;
};

template <typename dist_t>
class AlgorithmInterface {
public:
    virtual void addPoint(void* datapoint, labeltype label) = 0;
    virtual std::priority_queue<std::pair<dist_t, labeltype>> searchKnn(void*, size_t) = 0;
    virtual void saveIndex(const std::string& location) = 0;
    virtual ~AlgorithmInterface() {}
};

namespace detail {
template <typename T>
T device_only(const void*, const void*, const void*) {
    throw std::logic_error("b200nn: distances are evaluated on the GPU; this DISTFUNC is a placeholder");
}
template <typename T, int METRIC, size_t ELEM>
class GpuSpace : public SpaceInterface<T>, public b200nn::GpuMetric {
    size_t dim_, data_size_;
    int order_;

public:
    // order: accumulation order of the reference kernel that this build of the reference would pick.
    // Default 4 (SSE) = brute_force_search/src/CMakeLists.txt:4 (no -mavx); pass 8 for an -mavx build.
    explicit GpuSpace(size_t dim, int order = 4) : dim_(dim), data_size_(dim * ELEM), order_(order) {
        if (METRIC != B200NN_METRIC_L2_U8) {
            if (dim % 4 != 0) order_ = 1;                 // InnerProduct / L2Sqr scalar loops
            else if (dim % 16 != 0 && order_ == 8) order_ = 4;  // *SIMD4Ext
        }
    }
    size_t get_data_size() { return data_size_; }
    DISTFUNC<T> get_dist_func() { return &device_only<T>; }
    void* get_dist_func_param() { return &dim_; }
    int b200nn_metric() const { return METRIC; }
    int b200nn_order() const { return order_; }
    size_t b200nn_dim() const { return dim_; }
};
}  // namespace detail

typedef detail::GpuSpace<float, B200NN_METRIC_IP, sizeof(float)> InnerProductSpace;
typedef detail::GpuSpace<float, B200NN_METRIC_L2, sizeof(float)> L2Space;
typedef detail::GpuSpace<int, B200NN_METRIC_L2_U8, 1> L2SpaceI;

template <typename dist_t>
class BruteforceSearch : public AlgorithmInterface<dist_t> {
    b200nn_flat_t h_ = nullptr;
    int metric_ = 0, order_ = 4;
    size_t dim_ = 0;
    // addPoint is called once per vector by the reference's CLIs; rows are buffered on the host and
    // shipped in bulk, while the reference's error checks still fire at the offending addPoint.
    std::vector<unsigned char> pend_rows_;
    std::vector<uint64_t> pend_labels_;
    std::unordered_set<uint64_t> labels_;
    size_t count_ = 0;
    void flush() {
        if (pend_labels_.empty()) return;
        b200nn::check(b200nn_flat_add(h_, pend_rows_.data(), pend_labels_.data(), pend_labels_.size()));
        pend_rows_.clear();
        pend_labels_.clear();
    }

    void bind(SpaceInterface<dist_t>* s) {
        b200nn::GpuMetric* g = dynamic_cast<b200nn::GpuMetric*>(s);
        if (!g) throw std::runtime_error("b200nn: the space must be one of InnerProductSpace / L2Space / L2SpaceI of hnswlib_gpu.hpp");
        metric_ = g->b200nn_metric();
        order_ = g->b200nn_order();
        dim_ = g->b200nn_dim();
        data_size_ = s->get_data_size();
    }

public:
    size_t maxelements_ = 0;
    size_t data_size_ = 0;

    BruteforceSearch(SpaceInterface<dist_t>* s) { bind(s); }
    BruteforceSearch(SpaceInterface<dist_t>* s, const std::string& location) { loadIndex(location, s); }
    BruteforceSearch(SpaceInterface<dist_t>* s, size_t maxElements) {
        bind(s);
        maxelements_ = maxElements;
        b200nn::check(b200nn_flat_create(b200nn::default_ctx(), metric_, order_, dim_, maxElements, &h_));
    }
    ~BruteforceSearch() { b200nn_flat_destroy(h_); }
    BruteforceSearch(const BruteforceSearch&) = delete;
    BruteforceSearch& operator=(const BruteforceSearch&) = delete;

    size_t size() const { return count_; }

    void addPoint(void* datapoint, labeltype label) {
        if (labels_.count((uint64_t)label)) throw std::runtime_error("Ids have to be unique");  // brutoforce.hpp:44-45
        if (count_ >= maxelements_) throw std::runtime_error("The number of elements exceeds the specified limit\n");  // :48-50
        const unsigned char* p = static_cast<const unsigned char*>(datapoint);
        pend_rows_.insert(pend_rows_.end(), p, p + data_size_);
        pend_labels_.push_back((uint64_t)label);
        labels_.insert((uint64_t)label);
        count_++;
        if (pend_labels_.size() >= 8192) flush();
    }
    // Deviation, on purpose: for a label that was never added the reference's `dict_external_to_internal[cur_external]`
    // (brutoforce.hpp:59) default-inserts internal id 0 and so silently deletes the FIRST stored row; here it is an error.
    void removePoint(labeltype cur_external) {
        if (!labels_.count((uint64_t)cur_external)) throw std::runtime_error("removePoint: label not found");
        flush();
        b200nn::check(b200nn_flat_remove(h_, (uint64_t)cur_external));
        labels_.erase((uint64_t)cur_external);
        count_--;
    }

    // all queries of a batch in one launch; result[i] is a max-heap like searchKnn's
    std::vector<std::priority_queue<std::pair<dist_t, labeltype>>> searchKnnBatch(const void* queries, size_t nq, size_t k) {
        flush();
        std::vector<dist_t> d(nq * k);
        std::vector<uint64_t> l(nq * k);
        b200nn::check(b200nn_flat_search(h_, queries, nq, k, d.data(), l.data()));
        std::vector<std::priority_queue<std::pair<dist_t, labeltype>>> out(nq);
        for (size_t i = 0; i < nq; i++)
            for (size_t j = 0; j < k; j++)
                if (l[i * k + j] != UINT64_MAX) out[i].push(std::pair<dist_t, labeltype>(d[i * k + j], (labeltype)l[i * k + j]));
        return out;
    }
    std::priority_queue<std::pair<dist_t, labeltype>> searchKnn(void* query_data, size_t k) {
        return searchKnnBatch(query_data, 1, k)[0];
    }

    void saveIndex(const std::string& location) {
        flush();
        b200nn::check(b200nn_flat_save(h_, location.c_str()));
    }
    void loadIndex(const std::string& location, SpaceInterface<dist_t>* s) { load_file(location, s, false); }

protected:
    // hnsw_format: the file is a HierarchicalNSW::saveIndex file (hnswalg.h:491-519) whose vectors + labels are taken
    void load_file(const std::string& location, SpaceInterface<dist_t>* s, bool hnsw_format) {
        bind(s);
        if (h_) b200nn_flat_destroy(h_);
        h_ = nullptr;
        pend_rows_.clear();
        pend_labels_.clear();
        labels_.clear();
        b200nn::check(hnsw_format ? b200nn_flat_load_hnsw(b200nn::default_ctx(), metric_, order_, dim_, location.c_str(), &h_)
                                  : b200nn_flat_load(b200nn::default_ctx(), metric_, order_, dim_, location.c_str(), &h_));
        // the reference restores maxelements_ from the file (brutoforce.hpp:113) and keeps its label map implicit in data_;
        // here capacity and labels come back from the library so that addPoint after a load sees the same limit and the
        // same "Ids have to be unique" check as before the save
        b200nn::check(b200nn_flat_info(h_, &maxelements_, &count_, nullptr, 0));
        std::vector<uint64_t> labs(count_);
        b200nn::check(b200nn_flat_info(h_, nullptr, nullptr, labs.data(), labs.size()));
        labels_.insert(labs.begin(), labs.end());
    }
};

// makeIdx.cpp / siftsIndex.cpp construct `HierarchicalNSW<float>(space, max_elements, M, efConstruction)`
// and call setEf(); the exact GPU scan needs none of these knobs.  This subclass accepts the same
// constructor shape so those call sites compile unchanged once the member is retyped
// (hnsw_sifts_retrieval/siftsIndex.hpp:49, makeIdx.cpp:321-325, siftsIndex.cpp:51).
template <typename dist_t>
class ExactNSW : public BruteforceSearch<dist_t> {
public:
    ExactNSW(SpaceInterface<dist_t>* s, size_t max_elements, size_t /*M*/ = 16, size_t /*ef_construction*/ = 200)
        : BruteforceSearch<dist_t>(s, max_elements) {}
    // HierarchicalNSW(space, location, nmslib = false) (hnswalg.h:31-34): `location` is the HNSW index file the
    // reference's makeIdx wrote (or a file saved by this class, which is BruteforceSearch's format) -- told apart by the
    // header: a HierarchicalNSW file starts with offsetLevel0_ == 0, a brute-force file with maxelements_ > 0
    ExactNSW(SpaceInterface<dist_t>* s, const std::string& location, bool /*nmslib*/ = false) : BruteforceSearch<dist_t>(s) {
        std::ifstream in(location.c_str(), std::ios::binary);
        size_t first = 1;
        in.read((char*)&first, sizeof first);
        if (!in) throw std::runtime_error("Cannot open index file " + location);
        this->load_file(location, s, first == 0);
    }
    void setEf(size_t) {}  // exact search: nothing to tune
};

#ifdef B200NN_HNSW_DROP_IN
// For sources that name the class directly (hnsw_sifts_retrieval/siftsIndex.hpp:49 `hnswlib::HierarchicalNSW<float>* appr_alg`,
// makeIdx.cpp:321-325): with -DB200NN_HNSW_DROP_IN they compile UNMODIFIED and get the exact GPU scan.
template <typename dist_t>
using HierarchicalNSW = ExactNSW<dist_t>;
#endif

}  // namespace hnswlib
