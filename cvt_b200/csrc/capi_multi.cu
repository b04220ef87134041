// capi_multi.cu -- the row-sharded index across the GPUs of one box, behind the C ABI.
//
// Two ways to run it, both on top of the single-GPU entry points of capi_pq.cu:
//   * one process per GPU (torchrun): every rank owns a b200nn_pq shard and a b200nn_comm (ncclCommInitRank);
//     b200nn_pq_search_sharded_dev = local scan -> ONE ncclAllGather of the per-rank top-k records -> merge.
//   * one process, several GPUs (a C++ host such as the IVFOPQ shim): b200nn_mpq owns one context + shard per device and a
//     worker thread each.  Devices that can address each other (NVLink/NVSwitch peer access) skip the collective
//     altogether: device g merges ITS chunk of the queries by reading the other shards' sorted key lists straight out of
//     peer memory inside the merge kernel (the gather and the merge are one kernel; 1/G of the all-gather traffic) and
//     copies its slice of the result to the host.  Without peer access, or with B200NN_EXCHANGE=nccl, the shards exchange
//     with one ncclAllGather (ncclCommInitAll) as the multi-process path does.
// Row ids: shard-local rows map to global row ids through a per-shard table that ascends with the local row, so the
// (score, id) order of a shard's candidates is the global order and the merged result equals a single index's.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded when there is one), so the library
// itself has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

#include "capi_common.cuh"
#include "capi_pq_internal.cuh"
#include "topk.cuh"

using namespace b200nn;

namespace {

// ------------------------------------------------------------------------------------------------ NCCL binding
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("B200NN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("NCCL is not loadable (libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
            return;
        }
#define B2_SYM(field, name)                                                             \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));         \
    if (!api.field) api.error = std::string("NCCL symbol missing: ") + name;
        B2_SYM(GetUniqueId, "ncclGetUniqueId")
        B2_SYM(CommInitRank, "ncclCommInitRank")
        B2_SYM(CommInitAll, "ncclCommInitAll")
        B2_SYM(CommDestroy, "ncclCommDestroy")
        B2_SYM(AllGather, "ncclAllGather")
        B2_SYM(AllReduce, "ncclAllReduce")
        B2_SYM(GetErrorString, "ncclGetErrorString")
        B2_SYM(GetVersion, "ncclGetVersion")
#undef B2_SYM
    });
    return &api;
}

#define B2_NCCL(api, call)                                                                                             \
    do {                                                                                                               \
        ncclResult_t _r = (call);                                                                                      \
        if (_r != ncclSuccess) {                                                                                       \
            b200nn::set_last_error(std::string(#call) + ": " + (api)->GetErrorString(_r) + " (" + __FILE__ + ":" +     \
                                   std::to_string(__LINE__) + ")");                                                    \
            return B200NN_ERR_CUDA;                                                                                    \
        }                                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------ small kernels
// shard-local row id -> global row id inside packed (score, id) records; empty slots (KEY_MAX) stay empty
__global__ void remap_ids_kernel(unsigned long long* keys, long long n, const uint32_t* __restrict__ gid) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key != KEY_MAX) keys[i] = (key & 0xFFFFFFFF00000000ull) | gid[(uint32_t)key];
    }
}
__global__ void iota_u32_kernel(uint32_t* p, long long n, uint32_t start) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = start + (uint32_t)i;
}
// elementwise min of non-negative floats held by a peer (group scores of another shard), read through peer memory
__global__ void min_into_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = fminf(dst[i], src[i]);
}

inline unsigned grid1d(long long work) { return (unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, 148 * 8)); }

// ------------------------------------------------------------------------------------------------ worker threads
// one per device of an mpq: the per-device part of every call is issued concurrently (kernel launches and copies of
// 8 shards from one host thread would serialise ~40 us of launch latency per shard in front of a ~1.5 ms step)
class Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<int()> job;
    bool has_job = false, stop = false, done = true;
    int rc = 0;
    std::string err;

public:
    Worker() {
        th = std::thread([this] {
            for (;;) {
                std::function<int()> j;
                {
                    std::unique_lock<std::mutex> l(mu);
                    cv.wait(l, [this] { return has_job || stop; });
                    if (stop) return;
                    j = std::move(job);
                    has_job = false;
                }
                const int r = j();
                const std::string e = r ? std::string(b200nn_last_error()) : std::string();  // last_error is thread-local
                {
                    std::lock_guard<std::mutex> l(mu);
                    rc = r;
                    err = e;
                    done = true;
                }
                cv_done.notify_all();
            }
        });
    }
    ~Worker() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    void submit(std::function<int()> j) {
        {
            std::lock_guard<std::mutex> l(mu);
            job = std::move(j);
            has_job = true;
            done = false;
        }
        cv.notify_all();
    }
    int wait(std::string* e) {
        std::unique_lock<std::mutex> l(mu);
        cv_done.wait(l, [this] { return done; });
        if (rc && e) *e = err;
        return rc;
    }
};

}  // namespace

// ================================================================================================ communicator
struct b200nn_comm {
    b200nn_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    DevBuf<unsigned long long> keys, gathered;
};

extern "C" {

int b200nn_comm_get_unique_id(void* id128) {
    if (!id128) B2_FAIL(B200NN_ERR_INVALID, "comm_get_unique_id: NULL argument");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) B2_FAIL(B200NN_ERR_UNSUPPORTED, api->error);
    ncclUniqueId id;
    B2_NCCL(api, api->GetUniqueId(&id));
    static_assert(sizeof(id) == B200NN_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id128, &id, sizeof id);
    return 0;
}

int b200nn_comm_create(b200nn_ctx_t ctx, int rank, int nranks, const void* id128, b200nn_comm_t* out) {
    if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128))
        B2_FAIL(B200NN_ERR_INVALID, "comm_create: bad arguments");
    *out = nullptr;
    std::unique_ptr<b200nn_comm> c(new b200nn_comm());
    c->ctx = ctx; c->rank = rank; c->nranks = nranks;
    if (nranks > 1) {
        NcclApi* api = nccl_api();
        if (!api->error.empty()) B2_FAIL(B200NN_ERR_UNSUPPORTED, api->error);
        std::lock_guard<std::mutex> g(ctx->mu);
        B2_CUDA(cudaSetDevice(ctx->c.device));
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        B2_NCCL(api, api->CommInitRank(&c->comm, nranks, id, rank));
    }
    *out = c.release();
    return 0;
}

void b200nn_comm_destroy(b200nn_comm_t c) {
    if (!c) return;
    if (c->comm) {
        cudaSetDevice(c->ctx->c.device);
        cudaStreamSynchronize(c->ctx->c.stream);
        nccl_api()->CommDestroy(c->comm);
    }
    delete c;
}

int b200nn_comm_info(b200nn_comm_t c, int* rank, int* nranks, int* nccl_version) {
    if (!c) B2_FAIL(B200NN_ERR_INVALID, "comm is NULL");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (nccl_version) {
        *nccl_version = 0;
        if (c->comm) nccl_api()->GetVersion(nccl_version);
    }
    return 0;
}

// Host-only: how N ranks split (rows x queries).  Cost model of one step of a (row shards R x query chunks N/R) grid,
// constants measured on B200 (profiles/): streaming rate of the fused scan, the per-(CTA, segment) top-k warm-up
// (parked rows + selection + streamed candidates ~ k ln(rows/2048)), rotate + LUT build per query, and the exchange
// (all-gather launch + payload + merge).  R = N (plain row sharding: memory-minimal, what SURVEY.md 8(e) describes) is kept
// whenever it is within 2 % of the best grid; a database so small that a 1/N shard no longer amortises the warm-ups
// (cfg3: 16 MB of codes) is replicated in row blocks and the BATCH is split instead.
static double layout_cost(int world, int R, long long n_rows, long long batch, int M, int k, int sm) {
    const int Q = world / R;
    const double rows = (double)((n_rows + R - 1) / R), bq = (double)((batch + Q - 1) / Q);
    const int qw = std::max(1, 128 / M);
    const long long groups = ((long long)bq + qw - 1) / qw;
    auto segment = [&](double r) { return 30e-6 + qw * 1.5 * k * log(std::max(r, 4096.0) / 2048.0) * 0.015e-6; };
    const long long waves = groups / sm, rem = groups % sm;
    const double t_warm = waves * segment(rows) + (rem ? 1.5 * segment((double)rem * rows / sm) : 0.0);
    const double t_scan = bq * rows * M / 6.4e12, t_lut = bq * 0.012e-6;
    const double gathered = (double)world * bq * k * 8;
    const double t_x = world == 1 ? 0.0 : 40e-6 + gathered / 300e9 + (R > 1 ? gathered / 1.0e12 : 0.0);
    return t_scan + t_warm + t_lut + t_x;
}

int b200nn_plan_layout(int n_ranks, uint64_t n_rows, uint64_t batch, int M, int k, int sm_count, int* row_shards, int* query_chunks) {
    if (n_ranks < 1 || M < 1 || k < 1 || sm_count < 1 || !row_shards || !query_chunks) B2_FAIL(B200NN_ERR_INVALID, "plan_layout: bad arguments");
    double best = 1e300;
    std::vector<std::pair<int, double>> cand;
    for (int r = 1; r <= n_ranks; r++)
        if (n_ranks % r == 0) {
            cand.emplace_back(r, layout_cost(n_ranks, r, (long long)n_rows, (long long)batch, M, k, sm_count));
            best = std::min(best, cand.back().second);
        }
    int R = n_ranks;
    for (auto& c : cand)
        if (c.second <= best * 1.02) R = c.first;  // the largest R within tolerance (candidates ascend)
    *row_shards = R;
    *query_chunks = n_ranks / R;
    return 0;
}

// One step of the sharded search on this rank (collective: every rank of `comm` calls it with the same nq, k, row_shards).
// The ranks form a (query chunk x row shard) grid, rank = chunk * row_shards + shard (row_shards = nranks: plain row
// sharding, every rank scans its rows for the whole batch).  q_raw_dev holds the WHOLE batch; this rank scans its chunk
// of it over its shard, the [chunk_q, k] records of all ranks are exchanged with ONE ncclAllGather and every rank merges
// the row shards of every chunk: out_* receive the full [nq, k] result on every rank.
int b200nn_pq_search_sharded_dev(b200nn_pq_t shard, b200nn_comm_t c, int row_shards, const float* q_raw_dev, size_t nq, int nprobe,
                                 size_t k, uint64_t id_base, float* out_dist_dev, uint64_t* out_id_dev) {
    if (!shard || !c || (nq && !q_raw_dev)) B2_FAIL(B200NN_ERR_INVALID, "pq_search_sharded: NULL argument");
    const int W = c->nranks, R = row_shards;
    if (R < 1 || W % R) B2_FAIL(B200NN_ERR_INVALID, "pq_search_sharded: row_shards must divide the number of ranks");
    if (!nq) return 0;
    const int Q = W / R, chunk = c->rank / R;
    int D = 0;
    int rc;
    if ((rc = b200nn_pq_info(shard, &D, nullptr, nullptr, nullptr, nullptr, nullptr))) return rc;
    const size_t cq = (nq + Q - 1) / Q;
    const size_t lo = std::min(nq, (size_t)chunk * cq), hi = std::min(nq, lo + cq);
    if (W == 1) return b200nn_pq_search_dev(shard, q_raw_dev, nq, nprobe, k, out_dist_dev, out_id_dev, nullptr, id_base);
    Ctx* x = &c->ctx->c;
    {
        std::lock_guard<std::mutex> g(c->ctx->mu);
        B2_CUDA(cudaSetDevice(x->device));
        if ((rc = c->keys.ensure(cq * k)) || (rc = c->gathered.ensure((size_t)W * cq * k))) return rc;
        if (hi - lo < cq)  // ragged last chunk: the missing rows are empty records
            B2_CUDA(cudaMemsetAsync(c->keys.p + (hi - lo) * k, 0xFF, (cq - (hi - lo)) * k * sizeof(unsigned long long), x->stream));
    }
    if (hi > lo && (rc = b200nn_pq_search_dev(shard, q_raw_dev + lo * D, hi - lo, nprobe, k, nullptr, nullptr, (uint64_t*)c->keys.p, id_base)))
        return rc;
    NcclApi* api = nccl_api();
    std::lock_guard<std::mutex> g(c->ctx->mu);
    B2_CUDA(cudaSetDevice(x->device));
    B2_NCCL(api, api->AllGather(c->keys.p, c->gathered.p, cq * k, ncclUint64, c->comm, x->stream));  // the single collective of the path
    // gathered = [Q][R][cq][k]; query q of the batch is row q % cq of chunk q / cq
    return launch_topk_merge(x, c->gathered.p, R, (long long)nq, (int)k, (long long)(cq * k), out_dist_dev, nullptr,
                             (unsigned long long*)out_id_dev, nullptr);
}

}  // extern "C"

// ================================================================================================ single-process multi-GPU index
namespace {
struct Shard {
    int device = 0;
    b200nn_ctx_t ctx = nullptr;
    b200nn_pq_t pq = nullptr;
    ncclComm_t comm = nullptr;
    long long n = 0;                 // rows of this shard
    DevBuf<uint32_t> gid;            // local row -> global row id (ascending)
    DevBuf<float> q_dev, dist, scores;
    DevBuf<unsigned long long> keys, gathered, ids;
    DevBuf<const unsigned long long*> peer_ptrs;  // [G] key list of every shard (peer memory)
    cudaEvent_t ready = nullptr, done = nullptr;   // keys/scores of this shard complete | this shard has finished reading its peers
    std::unique_ptr<Worker> worker;
};
}  // namespace

struct b200nn_mpq {
    std::vector<std::unique_ptr<Shard>> s;
    std::mutex mu;
    int D = 0, K = 0, M = 0, ksub = 0;
    long long n_total = 0, n_groups = 0;
    bool p2p = false;   // merge straight out of peer memory (no collective)
    float timing_ms = 0;
};

namespace {

// run fn(g) on every shard's worker, wait for all, report the first failure
int for_shards(b200nn_mpq* m, const std::function<int(int)>& fn) {
    const int G = (int)m->s.size();
    if (G == 1) return fn(0);
    for (int g = 0; g < G; g++) m->s[g]->worker->submit([&fn, g] { return fn(g); });
    int rc = 0;
    std::string err;
    for (int g = 0; g < G; g++) {
        std::string e;
        const int r = m->s[g]->worker->wait(&e);
        if (r && !rc) { rc = r; err = e; }
    }
    if (rc) set_last_error(err);
    return rc;
}

int mpq_destroy_impl(b200nn_mpq* m) {
    if (!m) return 0;
    for (auto& sp : m->s) {
        Shard* s = sp.get();
        if (!s) continue;
        s->worker.reset();
        if (s->ctx) {
            cudaSetDevice(s->device);
            cudaDeviceSynchronize();
        }
        if (s->comm) nccl_api()->CommDestroy(s->comm);
        if (s->pq) b200nn_pq_destroy(s->pq);
        if (s->ready) cudaEventDestroy(s->ready);
        if (s->done) cudaEventDestroy(s->done);
        s->gid.release(); s->q_dev.release(); s->dist.release(); s->scores.release(); s->keys.release(); s->gathered.release();
        s->ids.release(); s->peer_ptrs.release();
        if (s->ctx) b200nn_ctx_destroy(s->ctx);
    }
    delete m;
    return 0;
}

// contexts, peer access or NCCL communicators; the per-shard indexes are created by `make_pq(ctx, &pq)`
int mpq_init(const int* devices, int n_devices, const std::function<int(b200nn_ctx_t, b200nn_pq_t*)>& make_pq, b200nn_mpq_t* out) {
    if (!devices || n_devices < 1 || !out) B2_FAIL(B200NN_ERR_INVALID, "mpq_create: bad arguments");
    *out = nullptr;
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j] && !getenv("B200NN_ALLOW_DUPLICATE_DEVICES"))
                B2_FAIL(B200NN_ERR_INVALID, "mpq_create: a device is listed twice (B200NN_ALLOW_DUPLICATE_DEVICES=1 permits it for single-GPU testing)");
    b200nn_mpq* m = new b200nn_mpq();
    auto fail = [&](int rc) { const std::string e = b200nn_last_error(); mpq_destroy_impl(m); set_last_error(e); return rc; };
    int rc;
    for (int g = 0; g < n_devices; g++) {
        m->s.emplace_back(new Shard());
        Shard* s = m->s.back().get();
        s->device = devices[g];
        if ((rc = b200nn_ctx_create(devices[g], &s->ctx))) return fail(rc);
        if ((rc = make_pq(s->ctx, &s->pq))) return fail(rc);
        if (cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming) != cudaSuccess) {
            set_last_error("mpq_create: cudaEventCreate failed");
            return fail(B200NN_ERR_CUDA);
        }
        if (n_devices > 1) s->worker.reset(new Worker());
    }
    if ((rc = b200nn_pq_info(m->s[0]->pq, &m->D, &m->K, &m->M, &m->ksub, nullptr, nullptr))) return fail(rc);
    // exchange: peer memory when every pair of devices can address each other, else (or on request) NCCL
    const char* ex = getenv("B200NN_EXCHANGE");
    bool p2p = !(ex && std::string(ex) == "nccl");
    for (int a = 0; a < n_devices && p2p; a++)
        for (int b = 0; b < n_devices && p2p; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) p2p = false;
        }
    if (ex && std::string(ex) == "p2p" && !p2p) { set_last_error("mpq_create: B200NN_EXCHANGE=p2p but the devices cannot address each other"); return fail(B200NN_ERR_UNSUPPORTED); }
    if (p2p) {
        for (int a = 0; a < n_devices; a++) {
            cudaSetDevice(devices[a]);
            for (int b = 0; b < n_devices; b++) {
                if (devices[a] == devices[b]) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_last_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                    return fail(B200NN_ERR_CUDA);
                }
                cudaGetLastError();
            }
        }
    } else if (n_devices > 1) {
        NcclApi* api = nccl_api();
        if (!api->error.empty()) { set_last_error(api->error); return fail(B200NN_ERR_UNSUPPORTED); }
        std::vector<ncclComm_t> comms(n_devices);
        const ncclResult_t r = api->CommInitAll(comms.data(), n_devices, devices);
        if (r != ncclSuccess) { set_last_error(std::string("ncclCommInitAll: ") + api->GetErrorString(r)); return fail(B200NN_ERR_CUDA); }
        for (int g = 0; g < n_devices; g++) m->s[g]->comm = comms[g];
    }
    m->p2p = p2p;
    *out = m;
    return 0;
}

// append n rows: contiguous pieces of the call's rows go to the shards, global ids = rows in arrival order
int mpq_add_impl(b200nn_mpq* m, const float* x, size_t n, const int32_t* group_ids, bool rotated) {
    if (!m || (n && !x)) B2_FAIL(B200NN_ERR_INVALID, "mpq_add: NULL argument");
    if (!n) return 0;
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->s.size();
    if ((unsigned long long)m->n_total + n > 0xFFFFFFFFull) B2_FAIL(B200NN_ERR_STATE, "mpq_add: more than 2^32-1 rows");
    const size_t per = (n + G - 1) / G;
    const long long base = m->n_total;
    std::vector<int32_t> iota;
    if (!group_ids) {  // one group per row, id = global row (as pq_add with NULL group ids on a single index)
        if ((unsigned long long)base + n > 0x7fffffffull) B2_FAIL(B200NN_ERR_STATE, "mpq_add: implicit group ids exceed 2^31-1");
        iota.resize(n);
        for (size_t i = 0; i < n; i++) iota[i] = (int32_t)(base + (long long)i);
        group_ids = iota.data();
    }
    int rc = for_shards(m, [&](int g) -> int {
        Shard* s = m->s[g].get();
        const size_t lo = std::min(n, (size_t)g * per), hi = std::min(n, lo + per);
        if (hi == lo) return 0;
        int r = rotated ? b200nn_pq_add_rotated(s->pq, x + lo * m->D, hi - lo, group_ids + lo)
                        : b200nn_pq_add(s->pq, x + lo * m->D, hi - lo, group_ids + lo);
        if (r) return r;
        std::lock_guard<std::mutex> g2(s->ctx->mu);
        B2_CUDA(cudaSetDevice(s->device));
        cudaStream_t st = s->ctx->c.stream;
        if ((r = s->gid.reserve((size_t)s->n + (hi - lo), (size_t)s->n, st))) return r;
        iota_u32_kernel<<<grid1d((long long)(hi - lo)), 256, 0, st>>>(s->gid.p + s->n, (long long)(hi - lo), (uint32_t)(base + (long long)lo));
        s->ctx->c.launches++;
        B2_CUDA(cudaStreamSynchronize(st));
        s->n += (long long)(hi - lo);
        return 0;
    });
    if (rc) return rc;
    m->n_total += (long long)n;
    for (size_t i = 0; i < n; i++) m->n_groups = std::max<long long>(m->n_groups, (long long)group_ids[i] + 1);
    return 0;
}

}  // namespace

extern "C" {

int b200nn_mpq_create(const int* devices, int n_devices, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                      const int32_t* perm, const float* R, float clamp_threshold, b200nn_mpq_t* out) {
    return mpq_init(devices, n_devices, [&](b200nn_ctx_t ctx, b200nn_pq_t* pq) {
        return b200nn_pq_create(ctx, D, K, M, ksub, coarse, codebooks, perm, R, clamp_threshold, pq);
    }, out);
}

int b200nn_mpq_load_model(const int* devices, int n_devices, const char* model_path, b200nn_mpq_t* out) {
    return mpq_init(devices, n_devices, [&](b200nn_ctx_t ctx, b200nn_pq_t* pq) { return b200nn_pq_load_model(ctx, model_path, pq); }, out);
}

void b200nn_mpq_destroy(b200nn_mpq_t m) { mpq_destroy_impl(m); }

int b200nn_mpq_info(b200nn_mpq_t m, int* D, int* K, int* M, int* ksub, uint64_t* n_rows, uint64_t* n_groups, int* n_devices, int* peer_exchange) {
    if (!m) B2_FAIL(B200NN_ERR_INVALID, "mpq is NULL");
    if (D) *D = m->D;
    if (K) *K = m->K;
    if (M) *M = m->M;
    if (ksub) *ksub = m->ksub;
    if (n_rows) *n_rows = (uint64_t)m->n_total;
    if (n_groups) *n_groups = (uint64_t)m->n_groups;
    if (n_devices) *n_devices = (int)m->s.size();
    if (peer_exchange) *peer_exchange = m->p2p ? 1 : 0;
    return 0;
}

int b200nn_mpq_shard_rows(b200nn_mpq_t m, uint64_t* rows /*[n_devices]*/) {
    if (!m || !rows) B2_FAIL(B200NN_ERR_INVALID, "mpq_shard_rows: NULL argument");
    for (size_t g = 0; g < m->s.size(); g++) rows[g] = (uint64_t)m->s[g]->n;
    return 0;
}

int b200nn_mpq_set_clamp(b200nn_mpq_t m, float clamp) {
    if (!m) B2_FAIL(B200NN_ERR_INVALID, "mpq is NULL");
    for (auto& s : m->s) {
        const int rc = b200nn_pq_set_clamp(s->pq, clamp);
        if (rc) return rc;
    }
    return 0;
}

int b200nn_mpq_rotate(b200nn_mpq_t m, const float* x, size_t n, float* y) {
    if (!m) B2_FAIL(B200NN_ERR_INVALID, "mpq is NULL");
    return b200nn_pq_rotate(m->s[0]->pq, x, n, y);
}

int b200nn_mpq_add(b200nn_mpq_t m, const float* x_raw, size_t n, const int32_t* group_ids) { return mpq_add_impl(m, x_raw, n, group_ids, false); }
int b200nn_mpq_add_rotated(b200nn_mpq_t m, const float* x_rot, size_t n, const int32_t* group_ids) { return mpq_add_impl(m, x_rot, n, group_ids, true); }

// IVFOPQ search across the shards: host buffers in and out, every device works on the whole batch over its rows,
// then device g merges query chunk g of all shards' candidates and writes that slice of the result.
int b200nn_mpq_search(b200nn_mpq_t m, const float* q_raw, size_t nq, int nprobe, size_t k, float* out_dist, uint64_t* out_id) {
    if (!m || (nq && (!q_raw || !out_dist || !out_id))) B2_FAIL(B200NN_ERR_INVALID, "mpq_search: NULL argument");
    if (!nq) return 0;
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->s.size();
    const size_t cq = (nq + G - 1) / G;
    NcclApi* api = m->p2p ? nullptr : nccl_api();
    // phase 1 (all devices concurrently): H2D, local scan -> sorted keys with global ids; mark them ready
    int rc = for_shards(m, [&](int g) -> int {
        Shard* s = m->s[g].get();
        int r;
        {
            std::lock_guard<std::mutex> g2(s->ctx->mu);
            B2_CUDA(cudaSetDevice(s->device));
            if ((r = s->q_dev.ensure(nq * m->D)) || (r = s->keys.ensure(nq * k)) || (r = s->dist.ensure(cq * k)) || (r = s->ids.ensure(cq * k))) return r;
            if (!m->p2p && G > 1 && (r = s->gathered.ensure((size_t)G * nq * k))) return r;
            if (m->p2p && (r = s->peer_ptrs.ensure(G))) return r;
            cudaStream_t st = s->ctx->c.stream;
            // the keys buffer is about to be overwritten: every peer must have finished reading the previous call's keys
            for (int p = 0; p < G; p++)
                if (p != g) B2_CUDA(cudaStreamWaitEvent(st, m->s[p]->done, 0));
            B2_CUDA(cudaMemcpyAsync(s->q_dev.p, q_raw, sizeof(float) * nq * m->D, cudaMemcpyHostToDevice, st));
        }
        if (s->n > 0) {
            if ((r = b200nn_pq_search_dev(s->pq, s->q_dev.p, nq, nprobe, k, nullptr, nullptr, (uint64_t*)s->keys.p, 0))) return r;
        }
        std::lock_guard<std::mutex> g2(s->ctx->mu);
        B2_CUDA(cudaSetDevice(s->device));
        cudaStream_t st = s->ctx->c.stream;
        if (s->n > 0) {
            remap_ids_kernel<<<grid1d((long long)(nq * k)), 256, 0, st>>>(s->keys.p, (long long)(nq * k), s->gid.p);
            s->ctx->c.launches++;
        } else {
            B2_CUDA(cudaMemsetAsync(s->keys.p, 0xFF, nq * k * sizeof(unsigned long long), st));
        }
        B2_CUDA(cudaEventRecord(s->ready, st));
        B2_CUDA(cudaGetLastError());
        return 0;
    });
    if (rc) return rc;
    // phase 2: exchange + merge of this device's query chunk, result slice -> host
    rc = for_shards(m, [&](int g) -> int {
        Shard* s = m->s[g].get();
        std::lock_guard<std::mutex> g2(s->ctx->mu);
        B2_CUDA(cudaSetDevice(s->device));
        Ctx* x = &s->ctx->c;
        cudaStream_t st = x->stream;
        const size_t lo = std::min(nq, (size_t)g * cq), hi = std::min(nq, lo + cq);
        int r = 0;
        if (m->p2p) {
            for (int p = 0; p < G; p++)
                if (p != g) B2_CUDA(cudaStreamWaitEvent(st, m->s[p]->ready, 0));
            if (hi > lo) {
                std::vector<const unsigned long long*> ptrs(G);
                for (int p = 0; p < G; p++) ptrs[p] = m->s[p]->keys.p + lo * k;
                B2_CUDA(cudaMemcpyAsync(s->peer_ptrs.p, ptrs.data(), sizeof(void*) * G, cudaMemcpyHostToDevice, st));
                B2_CUDA(cudaStreamSynchronize(st));  // ptrs is a host temporary (8 * G bytes)
                r = launch_topk_merge_ptrs(x, s->peer_ptrs.p, G, (long long)(hi - lo), (int)k, s->dist.p, nullptr, s->ids.p, nullptr);
            }
            B2_CUDA(cudaEventRecord(s->done, st));
        } else {
            const unsigned long long* src = s->keys.p;
            if (G > 1) {
                B2_NCCL(api, api->AllGather(s->keys.p, s->gathered.p, nq * k, ncclUint64, s->comm, st));
                src = s->gathered.p;
            }
            if (hi > lo)
                r = launch_topk_merge(x, src + lo * k, G, (long long)(hi - lo), (int)k, (long long)(nq * k), s->dist.p, nullptr, s->ids.p, nullptr);
            B2_CUDA(cudaEventRecord(s->done, st));
        }
        if (r) return r;
        if (hi > lo) {
            B2_CUDA(cudaMemcpyAsync(out_dist + lo * k, s->dist.p, sizeof(float) * (hi - lo) * k, cudaMemcpyDeviceToHost, st));
            B2_CUDA(cudaMemcpyAsync(out_id + lo * k, s->ids.p, sizeof(uint64_t) * (hi - lo) * k, cudaMemcpyDeviceToHost, st));
        }
        B2_CUDA(cudaStreamSynchronize(st));
        return check_dev_err(x, "mpq_search");
    });
    return rc;
}

// IVFOPQ::QueryThrehold across the shards: every device min-aggregates its rows into [nq, n_groups] (clamp-initialised);
// the shards' matrices are combined with an elementwise min (min is associative: SURVEY.md 8(e)) -- device 0 folds its
// peers' matrices in out of peer memory, or one ncclAllReduce(min) without peer access -- and device 0 returns the result.
int b200nn_mpq_scores(b200nn_mpq_t m, const float* q_raw, size_t nq, int nprobe, float* out_scores) {
    if (!m || (nq && (!q_raw || !out_scores))) B2_FAIL(B200NN_ERR_INVALID, "mpq_scores: NULL argument");
    if (!nq || !m->n_groups) return 0;
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->s.size();
    const long long ng = m->n_groups;
    const long long qc = std::max<long long>(1, std::min<long long>((long long)nq, std::min<long long>(1024, (1LL << 26) / ng)));
    NcclApi* api = m->p2p ? nullptr : nccl_api();
    for (long long q0 = 0; q0 < (long long)nq; q0 += qc) {
        const long long cqn = std::min<long long>(qc, (long long)nq - q0);
        int rc = for_shards(m, [&](int g) -> int {
            Shard* s = m->s[g].get();
            int r;
            {
                std::lock_guard<std::mutex> g2(s->ctx->mu);
                B2_CUDA(cudaSetDevice(s->device));
                if ((r = s->q_dev.ensure((size_t)qc * m->D)) || (r = s->scores.ensure((size_t)qc * ng))) return r;
                B2_CUDA(cudaStreamWaitEvent(s->ctx->c.stream, m->s[0]->done, 0));  // device 0 has finished reading the previous chunk
                B2_CUDA(cudaMemcpyAsync(s->q_dev.p, q_raw + q0 * m->D, sizeof(float) * cqn * m->D, cudaMemcpyHostToDevice, s->ctx->c.stream));
            }
            if ((r = b200nn_pq_scores_dev(s->pq, s->q_dev.p, (size_t)cqn, nprobe, (size_t)ng, s->scores.p))) return r;
            std::lock_guard<std::mutex> g2(s->ctx->mu);
            B2_CUDA(cudaSetDevice(s->device));
            if (!m->p2p && G > 1)
                B2_NCCL(api, api->AllReduce(s->scores.p, s->scores.p, (size_t)(cqn * ng), ncclFloat32, ncclMin, s->comm, s->ctx->c.stream));
            B2_CUDA(cudaEventRecord(s->ready, s->ctx->c.stream));
            if (g != 0) B2_CUDA(cudaStreamSynchronize(s->ctx->c.stream));
            return 0;
        });
        if (rc) return rc;
        Shard* s0 = m->s[0].get();
        std::lock_guard<std::mutex> g2(s0->ctx->mu);
        B2_CUDA(cudaSetDevice(s0->device));
        cudaStream_t st = s0->ctx->c.stream;
        if (m->p2p)
            for (int p = 1; p < G; p++) {
                B2_CUDA(cudaStreamWaitEvent(st, m->s[p]->ready, 0));
                min_into_kernel<<<grid1d(cqn * ng), 256, 0, st>>>(s0->scores.p, m->s[p]->scores.p, cqn * ng);
                s0->ctx->c.launches++;
            }
        B2_CUDA(cudaEventRecord(s0->done, st));
        B2_CUDA(cudaMemcpyAsync(out_scores + q0 * ng, s0->scores.p, sizeof(float) * cqn * ng, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        B2_CUDA(cudaGetLastError());
    }
    return 0;
}

// SaveIndex of the sharded index: rows of all shards in global order -> the reference's file (App. A-3)
int b200nn_mpq_save_index(b200nn_mpq_t m, const char* dir_or_path, const char* const* group_paths, size_t n_paths) {
    if (!m || !dir_or_path) B2_FAIL(B200NN_ERR_INVALID, "mpq_save_index: NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const long long n = m->n_total;
    std::vector<int> lists((size_t)n), groups((size_t)n);
    std::vector<unsigned char> codes((size_t)n * m->M);
    std::vector<float> coarse, cb;
    int rc;
    if ((rc = pq_model_to_host(m->s[0]->pq, &coarse, &cb, nullptr))) return rc;
    for (auto& sp : m->s) {
        Shard* s = sp.get();
        if (!s->n) continue;
        PQHostRows rows;
        if ((rc = pq_rows_to_host(s->pq, &rows))) return rc;
        std::vector<uint32_t> gid((size_t)s->n);
        cudaSetDevice(s->device);
        B2_CUDA(cudaMemcpy(gid.data(), s->gid.p, sizeof(uint32_t) * s->n, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < s->n; i++) {
            lists[gid[i]] = rows.lists[i];
            groups[gid[i]] = rows.groups[i];
            memcpy(&codes[(size_t)gid[i] * m->M], &rows.codes[(size_t)i * m->M], m->M);
        }
    }
    return pq_write_index_file(pq_index_file_name(dir_or_path, m->n_groups, m->D, m->K, m->M, m->ksub), m->D, m->K, m->M, m->ksub, m->n_groups,
                               coarse.data(), cb.data(), n, lists.data(), groups.data(), codes.data(), group_paths, n_paths);
}

// LoadIndex onto several devices: the file's rows (list by list, as stored) are dealt out in contiguous blocks
int b200nn_mpq_load_index(const int* devices, int n_devices, const char* path, const int32_t* perm, float clamp, b200nn_mpq_t* out) {
    if (!path || !out) B2_FAIL(B200NN_ERR_INVALID, "mpq_load_index: NULL argument");
    *out = nullptr;
    int D, K, M, ksub;
    long long ng;
    std::vector<float> coarse, cb;
    PQHostRows rows;
    int rc;
    if ((rc = pq_parse_index_file(path, &D, &K, &M, &ksub, &ng, &coarse, &cb, &rows))) return rc;
    b200nn_mpq* m = nullptr;
    if ((rc = b200nn_mpq_create(devices, n_devices, D, K, M, ksub, coarse.data(), cb.data(), perm, nullptr, clamp, &m))) return rc;
    const size_t n = rows.lists.size(), G = m->s.size(), per = (n + G - 1) / G;
    rc = for_shards(m, [&](int g) -> int {
        Shard* s = m->s[g].get();
        const size_t lo = std::min(n, (size_t)g * per), hi = std::min(n, lo + per);
        if (hi == lo) return 0;
        int r = b200nn_pq_append_coded(s->pq, rows.lists.data() + lo, rows.groups.data() + lo, rows.codes.data() + lo * M, hi - lo);
        if (r) return r;
        std::lock_guard<std::mutex> g2(s->ctx->mu);
        B2_CUDA(cudaSetDevice(s->device));
        if ((r = s->gid.reserve(hi - lo, 0, s->ctx->c.stream))) return r;
        iota_u32_kernel<<<grid1d((long long)(hi - lo)), 256, 0, s->ctx->c.stream>>>(s->gid.p, (long long)(hi - lo), (uint32_t)lo);
        B2_CUDA(cudaStreamSynchronize(s->ctx->c.stream));
        s->n = (long long)(hi - lo);
        return 0;
    });
    if (rc) {
        const std::string e = b200nn_last_error();
        mpq_destroy_impl(m);
        set_last_error(e);
        return rc;
    }
    m->n_total = (long long)n;
    m->n_groups = ng;
    for (size_t i = 0; i < n; i++) m->n_groups = std::max<long long>(m->n_groups, (long long)rows.groups[i] + 1);
    *out = m;
    return 0;
}

}  // extern "C"
