/*
 * b200nn.h -- C ABI of the B200-native quantized nearest-neighbour path (drop-in boundary).
 *
 * Each entry point replaces one reference interface (paths relative to willard-yuan/cvt):
 *   flat  : hnswlib::BruteforceSearch<dist_t>::addPoint/searchKnn/saveIndex/loadIndex
 *           brute_force_search/src/brutoforce.hpp:43-56,73-93,95-134 with the distance functions of
 *           space_ip.hpp:25-207 and hnsw_sifts_retrieval/hnswlib/space_l2.h:26-245
 *   pq    : IVFOPQ::LoadModel/Add/IndexDatabase/Query/QueryThrehold/SaveIndex
 *           opq/src/IVFOPQ.h:31-47, opq/src/IVFOPQ.cpp:64-583, get_sort_results opq/src/common.h:25-37
 *   sq    : cvtk::quant::Int8Quan::Int8Encode/Int8Decode/L2NormalizeVector
 *           scalar_quantization/scalar_quantization/int8_quan.h:17-37, int8_quan.cc:46-132
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on success,
 * <0 on error (message via b200nn_last_error(), thread-local).  Pointers are HOST pointers unless
 * the name ends in _dev (device pointers on the context's GPU; work is enqueued on the context
 * stream and NOT synchronised -- call b200nn_ctx_synchronize or sync the stream you supplied).
 * There is no CPU fallback: without a CUDA device ctx_create fails.
 */
#ifndef B200NN_H
#define B200NN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NN_OK 0
#define B200NN_ERR_INVALID -1   /* bad argument */
#define B200NN_ERR_CUDA -2      /* CUDA runtime / launch failure */
#define B200NN_ERR_IO -3        /* file format / open failure */
#define B200NN_ERR_UNSUPPORTED -4
#define B200NN_ERR_STATE -5     /* e.g. capacity exceeded, duplicate label */

typedef struct b200nn_ctx* b200nn_ctx_t;
typedef struct b200nn_flat* b200nn_flat_t;
typedef struct b200nn_pq* b200nn_pq_t;
typedef struct b200nn_sq* b200nn_sq_t;
typedef struct b200nn_proj* b200nn_proj_t;
typedef struct b200nn_comm* b200nn_comm_t; /* one rank of a multi-GPU job (NCCL communicator + exchange buffers) */
typedef struct b200nn_mpq* b200nn_mpq_t;   /* (O)PQ index row-sharded over several GPUs of one process */

/* ---- context ---------------------------------------------------------------------------- */
const char* b200nn_last_error(void);
const char* b200nn_version(void);
int b200nn_ctx_create(int device, b200nn_ctx_t* out);
void b200nn_ctx_destroy(b200nn_ctx_t ctx);
/* Adopt an external cudaStream_t (e.g. torch's current stream); NULL restores the private stream. */
int b200nn_ctx_set_stream(b200nn_ctx_t ctx, void* cuda_stream);
int b200nn_ctx_synchronize(b200nn_ctx_t ctx);
/* Number of kernels this library has launched through ctx so far (bench.py's gpu_launches). */
int b200nn_ctx_launch_count(b200nn_ctx_t ctx, uint64_t* out);
/* CUDA-event timing on the context stream: mark(slot) records an event, elapsed returns ms. */
int b200nn_ctx_event_record(b200nn_ctx_t ctx, int slot);
int b200nn_ctx_event_elapsed_ms(b200nn_ctx_t ctx, int slot_a, int slot_b, float* ms);

/* ---- flat exact index: BruteforceSearch<float|int> drop-in ---------------------------------
 * metric: 0 = 1 - <a,b>   (InnerProductSpace, space_ip.hpp)        elements f32, dist f32
 *         1 = sum (a-b)^2 (L2Space, space_l2.h:26-180)             elements f32, dist f32
 *         2 = sum (a-b)^2 (L2SpaceI, space_l2.h:186-245)           elements u8,  dist s32
 * order : accumulation order of the reference's SIMD kernels (bit-exact fp32 distances):
 *         1 = scalar loop, 4 = SSE (the brute-force CLI's own build flags), 8 = AVX (dim%16==0).
 * Result set = the k smallest under lexicographic (dist, label), ascending -- the order
 * brutoforce.hpp:73-93 yields once its max-heap is drained and reversed (brute_force.cpp:89-101). */
#define B200NN_METRIC_IP 0
#define B200NN_METRIC_L2 1
#define B200NN_METRIC_L2_U8 2
int b200nn_flat_create(b200nn_ctx_t ctx, int metric, int order, size_t dim, size_t max_elements, b200nn_flat_t* out);
void b200nn_flat_destroy(b200nn_flat_t idx);
/* addPoint for n rows; duplicate label -> ERR_STATE "Ids have to be unique"; capacity -> ERR_STATE. */
int b200nn_flat_add(b200nn_flat_t idx, const void* vectors, const uint64_t* labels, size_t n);
int b200nn_flat_remove(b200nn_flat_t idx, uint64_t label); /* removePoint: last row moves into the hole */
int b200nn_flat_size(b200nn_flat_t idx, size_t* out);
/* searchKnn for nq queries.  out_dist is f32 (metric 0,1) or s32 (metric 2), [nq,k]; out_label [nq,k].
 * If fewer than k rows are indexed the tail is filled with +inf / INT32_MAX and label UINT64_MAX
 * (the reference reads uninitialised memory there). */
int b200nn_flat_search(b200nn_flat_t idx, const void* queries, size_t nq, size_t k, void* out_dist, uint64_t* out_label);
int b200nn_flat_search_dev(b200nn_flat_t idx, const void* queries_dev, size_t nq, size_t k, void* out_dist_dev,
                           uint64_t* out_label_dev);
int b200nn_flat_save(b200nn_flat_t idx, const char* path);  /* byte format of brutoforce.hpp:95-106 */
int b200nn_flat_load(b200nn_ctx_t ctx, int metric, int order, size_t dim, const char* path, b200nn_flat_t* out);
/* the vectors and labels of a hnswlib::HierarchicalNSW<dist_t>::saveIndex file (hnsw_sifts_retrieval/hnswlib/hnswalg.h:491-519;
 * what makeSearch.cpp:19-22 opens) as a flat exact index: the graph is not used, searchKnn becomes exact. */
int b200nn_flat_load_hnsw(b200nn_ctx_t ctx, int metric, int order, size_t dim, const char* path, b200nn_flat_t* out);
/* capacity, row count and (optionally) the labels in row order */
int b200nn_flat_info(b200nn_flat_t idx, size_t* max_elements, size_t* n, uint64_t* labels_out, size_t labels_capacity);

/* ---- (O)PQ / IVFOPQ ----------------------------------------------------------------------
 * Model = D, K coarse centroids, M sub-quantizers x ksub(=256) codewords, and the "rotation":
 * a permutation perm[D] (y[i] = x[perm[i]], IVFOPQ::reorder, IVFOPQ.cpp:424-439) and/or a dense
 * row-major R[D,D] (y = R x, tcgen05 split-TF32 GEMM).  clamp_threshold is the reference's
 * `threhold = 1.0` (IVFOPQ.cpp:5,262,308); pass INFINITY to disable it. */
int b200nn_pq_create(b200nn_ctx_t ctx, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                     const int32_t* perm, const float* R, float clamp_threshold, b200nn_pq_t* out);
int b200nn_pq_load_model(b200nn_ctx_t ctx, const char* model_path, b200nn_pq_t* out); /* IVFOPQ::LoadModel format */
void b200nn_pq_destroy(b200nn_pq_t idx);
int b200nn_pq_set_clamp(b200nn_pq_t idx, float clamp_threshold);
int b200nn_pq_info(b200nn_pq_t idx, int* D, int* K, int* M, int* ksub, uint64_t* n_rows, uint64_t* n_groups);
/* a1: rotation of n raw rows. */
int b200nn_pq_rotate(b200nn_pq_t idx, const float* x, size_t n, float* y);
/* a2+a3: coarse argmin + residual PQ argmin of already-rotated rows (IVFOPQ::Add, :105-174). */
int b200nn_pq_encode(b200nn_pq_t idx, const float* x_rotated, size_t n, int32_t* out_list, uint8_t* out_codes);
/* rotate + encode + append.  group_ids = videoId per row (NULL: one new group per row, id = row). */
int b200nn_pq_add(b200nn_pq_t idx, const float* x_raw, size_t n, const int32_t* group_ids);
int b200nn_pq_add_dev(b200nn_pq_t idx, const float* x_raw_dev, size_t n, const int32_t* group_ids_dev);
/* IVFOPQ::Add proper: rows that LoadSingleFeatFile already reordered (IVFOPQ.cpp:459-461) -> encode + append. */
int b200nn_pq_add_rotated(b200nn_pq_t idx, const float* x_rotated, size_t n, const int32_t* group_ids);
/* read back what Add stored, rows [start, start+n) in insertion order. */
int b200nn_pq_get_rows(b200nn_pq_t idx, uint64_t start, size_t n, int32_t* out_list, int32_t* out_group,
                       uint8_t* out_codes);
/* a4+a5: coarse top-nprobe (in the reference's pop order) + residual LUTs [nq,nprobe,M,ksub]. */
int b200nn_pq_build_lut(b200nn_pq_t idx, const float* q_rotated, size_t nq, int nprobe, int32_t* out_lists,
                        float* out_lut);
/* IVFOPQ::QueryThrehold: out_scores [nq, n_groups], clamp-initialised, min-aggregated per group. */
int b200nn_pq_scores(b200nn_pq_t idx, const float* q_raw, size_t nq, int nprobe, float* out_scores);
/* device variant: rows of out_scores_dev are out_stride >= n_groups floats apart (a shard of a larger index writes into
 * the matrix of ALL groups; columns it holds no rows of keep the clamp value). */
int b200nn_pq_scores_dev(b200nn_pq_t idx, const float* q_raw_dev, size_t nq, int nprobe, size_t out_stride, float* out_scores_dev);
/* append n rows that are already coded (host pointers; what an index file holds): coarse list id, videoId
 * (NULL: id = row) and M code bytes per row.  Ids and code bytes are validated. */
int b200nn_pq_append_coded(b200nn_pq_t idx, const int32_t* lists, const int32_t* group_ids, const uint8_t* codes, size_t n);
/* What the reference's query main does with Query's output (opq/src/multi_frame_index_test.cpp:54-68): query
 * video v = frames [frame_off[v], frame_off[v+1]) of q_raw; per frame the QueryThrehold scores (above), summed
 * over the video's frames per indexed videoId in frame order (fp32, from 0.0f), then get_sort_results
 * (opq/src/common.h:25-37): the k smallest (score, videoId), ascending.  out_* are [n_videos, k]; when fewer than k
 * videos are indexed the tail is (+inf, UINT64_MAX).  Only the results leave the device. */
int b200nn_pq_query_groups(b200nn_pq_t idx, const float* q_raw, const int64_t* frame_off, size_t n_videos, int nprobe, size_t k,
                           float* out_score, uint64_t* out_group);
/* a1,a4,a5,a6,a7 fused: per-row ADC scores (clamped) -> k smallest under (score,row), ascending.
 * out_id = id_base + row index.  K==1 runs the TMA-staged conflict-free scan kernel. */
int b200nn_pq_search(b200nn_pq_t idx, const float* q_raw, size_t nq, int nprobe, size_t k, float* out_dist,
                     uint64_t* out_id);
/* device variant.  out_key_dev (may be NULL) additionally receives the packed sortable keys
 * ((orderable(score)<<32) | (uint32)(id_base+row)), the record exchanged between shards. */
int b200nn_pq_search_dev(b200nn_pq_t idx, const float* q_raw_dev, size_t nq, int nprobe, size_t k, float* out_dist_dev,
                         uint64_t* out_id_dev, uint64_t* out_key_dev, uint64_t id_base);
/* merge L sorted key lists per query ([L][nq][k] u64, e.g. the all-gathered shard results). */
int b200nn_topk_merge_dev(b200nn_ctx_t ctx, const uint64_t* keys_dev, int L, size_t nq, size_t k, float* out_dist_dev,
                          uint64_t* out_id_dev);
/* the same merge over the all-gathered records of a (query chunk x row shard) grid of ranks:
 * keys [n_chunks][L][chunk_q][k], rank = chunk * L + shard; batch query q = row q % chunk_q of chunk q / chunk_q. */
int b200nn_topk_merge_grid_dev(b200nn_ctx_t ctx, const uint64_t* keys_dev, int n_chunks, int L, size_t chunk_q, size_t nq,
                               size_t k, float* out_dist_dev, uint64_t* out_id_dev);
/* host-only: the work plan of the fused flat scan for nq queries x n_rows rows on sm_count SMs -- n_full
 * whole-shard CTAs, n_tail tail CTAs whose equal pieces of the last wave are described by 8 ints each
 * (two {query group, output slice, granule lo, granule hi} segments; granule = 64 rows), slices = lists per query. */
int b200nn_pq_scan_plan(int sm_count, int M, size_t nq, size_t n_rows, int* n_full, int* n_tail, int* slices, int32_t* desc,
                        size_t desc_capacity);
int b200nn_pq_save_index(b200nn_pq_t idx, const char* dir_or_path, const char* const* group_paths); /* App. A-3; group_paths holds n_groups entries */
/* the same with the length of group_paths stated: groups past n_paths get an empty name (IVFOPQ::Add called directly
 * leaves m_imgNum behind the highest videoId, so the shim's path table can be shorter than the group count) */
int b200nn_pq_save_index_n(b200nn_pq_t idx, const char* dir_or_path, const char* const* group_paths, size_t n_paths);
int b200nn_pq_load_index(b200nn_ctx_t ctx, const char* path, const int32_t* perm, float clamp, b200nn_pq_t* out);
/* ---- several GPUs (the coded database shards by rows, queries are replicated, ONE exchange of the
 * per-shard top-k records, a per-query merge).  Global ids order ties exactly as a single index does.
 *
 * (1) one process per GPU (torchrun, MPI, ...): rank 0 draws an id, every rank builds a communicator on its context's
 *     device (ncclCommInitRank; NCCL is bound at run time, libnccl.so.2) and calls pq_search_sharded_dev collectively. */
#define B200NN_COMM_ID_BYTES 128
int b200nn_comm_get_unique_id(void* id128);
int b200nn_comm_create(b200nn_ctx_t ctx, int rank, int nranks, const void* id128, b200nn_comm_t* out);
void b200nn_comm_destroy(b200nn_comm_t comm);
int b200nn_comm_info(b200nn_comm_t comm, int* rank, int* nranks, int* nccl_version);
/* Ranks form a (query chunk x row shard) grid, rank = chunk * row_shards + shard; row_shards = nranks is the plain
 * row-sharded layout.  q_raw_dev = the WHOLE batch [nq, D] on this rank's device, id_base = global id of the shard's
 * first row.  Local scan of this rank's query chunk -> one ncclAllGather of [chunk_q, k] records -> merge: out_*_dev
 * receive the full [nq, k] result on every rank.  Enqueued on the context stream, not synchronised. */
int b200nn_pq_search_sharded_dev(b200nn_pq_t shard, b200nn_comm_t comm, int row_shards, const float* q_raw_dev, size_t nq, int nprobe,
                                 size_t k, uint64_t id_base, float* out_dist_dev, uint64_t* out_id_dev);
/* host-only: the (row shards x query chunks) grid for n_ranks GPUs from a measured cost model; row_shards = n_ranks (plain
 * row sharding) unless a 1/n_ranks shard is too small to amortise the per-query top-k warm-up of the scan. */
int b200nn_plan_layout(int n_ranks, uint64_t n_rows, uint64_t batch, int M, int k, int sm_count, int* row_shards, int* query_chunks);
/* (2) one process, several GPUs: the IVFOPQ drop-in over devices[] (replaces the same IVFOPQ methods as b200nn_pq_*;
 *     the C++ shim picks it when $B200NN_DEVICES lists more than one device).  Every add call's rows are dealt to the
 *     shards in contiguous blocks.  Exchange: devices with peer access (NVLink) merge straight out of each other's
 *     memory -- device g merges query chunk g, the gather is inside the merge kernel; otherwise, or with
 *     $B200NN_EXCHANGE=nccl, one ncclAllGather (ncclCommInitAll).  Host buffers in and out. */
int b200nn_mpq_create(const int* devices, int n_devices, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                      const int32_t* perm, const float* R, float clamp_threshold, b200nn_mpq_t* out);
int b200nn_mpq_load_model(const int* devices, int n_devices, const char* model_path, b200nn_mpq_t* out);
void b200nn_mpq_destroy(b200nn_mpq_t idx);
int b200nn_mpq_info(b200nn_mpq_t idx, int* D, int* K, int* M, int* ksub, uint64_t* n_rows, uint64_t* n_groups, int* n_devices,
                    int* peer_exchange);
int b200nn_mpq_shard_rows(b200nn_mpq_t idx, uint64_t* rows /*[n_devices]*/);
int b200nn_mpq_set_clamp(b200nn_mpq_t idx, float clamp_threshold);
int b200nn_mpq_rotate(b200nn_mpq_t idx, const float* x, size_t n, float* y);
int b200nn_mpq_add(b200nn_mpq_t idx, const float* x_raw, size_t n, const int32_t* group_ids);
int b200nn_mpq_add_rotated(b200nn_mpq_t idx, const float* x_rotated, size_t n, const int32_t* group_ids);
int b200nn_mpq_search(b200nn_mpq_t idx, const float* q_raw, size_t nq, int nprobe, size_t k, float* out_dist, uint64_t* out_id);
int b200nn_mpq_scores(b200nn_mpq_t idx, const float* q_raw, size_t nq, int nprobe, float* out_scores); /* elementwise min over shards */
int b200nn_mpq_save_index(b200nn_mpq_t idx, const char* dir_or_path, const char* const* group_paths, size_t n_paths);
int b200nn_mpq_load_index(const int* devices, int n_devices, const char* path, const int32_t* perm, float clamp, b200nn_mpq_t* out);

/* timing of the last pq_search[_dev] stages in ms: [rotate, lut, scan, merge] (CUDA events). */
int b200nn_pq_last_timing(b200nn_pq_t idx, float* ms4);
/* bytes of device memory held by the coded database (codes as scanned). */
int b200nn_pq_scan_bytes(b200nn_pq_t idx, uint64_t* code_bytes);

/* ---- scalar quantizer: Int8Quan ------------------------------------------------------------ */
int b200nn_sq_create(b200nn_ctx_t ctx, int d, const float* vmin, const float* vdiff, b200nn_sq_t* out);
void b200nn_sq_destroy(b200nn_sq_t sq);
/* faiss RS_minmax (rs_arg = 0) over rows that are ALREADY L2-normalised (sq_train.cpp:84,100-132). */
int b200nn_sq_train_minmax(b200nn_ctx_t ctx, int d, const float* x, size_t n, float* vmin, float* vdiff);
/* Int8Encode for n rows (the reference encodes one; n rows = n calls).  l2norm!=0 normalises x IN
 * PLACE first, exactly as the reference mutates its argument (int8_quan.cc:76-78). */
int b200nn_sq_encode(b200nn_sq_t sq, float* x, size_t n, int l2norm, uint8_t* codes);
/* Int8Decode: variant 0 = the reference's double-precision formula (int8_quan.cc:126-130),
 * variant 1 = faiss 1.5.3's all-float decode (reached via Int8Decode(uint8_t*)/Int8DecodeFaiss). */
int b200nn_sq_decode(b200nn_sq_t sq, const uint8_t* codes, size_t n, int faiss_float_variant, float* x);
int b200nn_sq_encode_dev(b200nn_sq_t sq, float* x_dev, size_t n, int l2norm, uint8_t* codes_dev);

/* ---- front end (the steps that produce the vectors fed to the path) --------------------------------
 * cvtk::PCAUtils (pca_train_project/pca_online/pca_utils.h:10-27, pca_utils.cc:16-35): loadModel reads cv::PCA's
 * `mean` [K_in] and `vectors` (eigenvectors, row-major [N_out, K_in]); reduceDim = cv::PCA::project, i.e.
 * y = (x - mean) * vectors^T, followed (l2norm != 0) by the per-row L2 normalisation of pca_utils.cc:28-34.
 * Runs as a tcgen05 split-TF32 GEMM (fp32-level accuracy; the reference's OpenCV gemm accumulates in double):
 * K_in % 32 == 0, N_out % 64 == 0, and N_out in {64, 128, 256} when l2norm is requested. */
int b200nn_proj_create(b200nn_ctx_t ctx, int K_in, int N_out, const float* mean /*[K_in] or NULL*/, const float* vectors,
                       b200nn_proj_t* out);
/* PCAUtils::loadModel (pca_utils.cc:16-23): the model is cv::PCA's YAML as cv::FileStorage writes it (`vectors`, `values`,
 * `mean` opencv-matrix nodes; shipped: pca_train_project/model/*.yml).  pca_read_model is host-only: with NULL buffers it
 * returns the sizes, otherwise it fills mean [K_in], vectors [N_out, K_in] and/or values [N_out] (numbers parsed as double
 * and narrowed to float, as OpenCV does).  proj_load_model = read + proj_create. */
int b200nn_pca_read_model(const char* path, int* K_in, int* N_out, float* mean, float* vectors, float* values);
int b200nn_proj_load_model(b200nn_ctx_t ctx, const char* path, b200nn_proj_t* out);
void b200nn_proj_destroy(b200nn_proj_t p);
int b200nn_proj_apply(b200nn_proj_t p, const float* x /*[n,K_in]*/, size_t n, int l2norm, float* y /*[n,N_out]*/);
int b200nn_proj_apply_dev(b200nn_proj_t p, const float* x_dev, size_t n, int l2norm, float* y_dev);
/* siftsIDX::rootSift (hnsw_sifts_retrieval/siftsIndex.cpp:54-71, eps = 1e-7 there), in place over n rows of d floats:
 * abs -> / (L1 + eps) -> sqrt -> cv::normalize(NORM_L2); both sums accumulate in double, in column order. */
int b200nn_rootsift(b200nn_ctx_t ctx, float* x, size_t n, int d, float eps);
int b200nn_rootsift_dev(b200nn_ctx_t ctx, float* x_dev, size_t n, int d, float eps);

/* ---- training ----------------------------------------------------------------
 * TrainPQ (opq/train_codebook/train_PQ_codebook.h:23-54): CoarseQuan (train_PQ_codebook.cpp:150-199) = k-means with
 * coarseK centroids over the reordered training rows + residue of every row; ProdQuan (:201-244) = one k-means with
 * pq_k centroids per sub-space of the residue; SaveCodebook (:247-288) = the model file IVFOPQ::LoadModel reads.
 * The reference calls yael's kmeans (un-vendored, random initialisation: numerically unpinned); here it is a
 * deterministic Lloyd iteration (init, assignment arithmetic, update summation order, empty-cluster rule and stop rule
 * are fixed: cvt_b200/csrc/capi_train.cu), all floating-point work on the device.
 * max_iter = 0 means "until convergence" (capped at 10000 updates), as niter = 0 does in the reference. */
int b200nn_kmeans(b200nn_ctx_t ctx, const float* x /*[n,d]*/, size_t n, int d, int k, int max_iter, uint64_t seed,
                  float* centroids /*[k,d]*/, int32_t* assign /*[n] or NULL*/, float* dist /*[n] or NULL*/, int* iters_done /*or NULL*/,
                  double* mse /*or NULL*/);
/* host-only: the integer bookkeeping of that iteration, exported so that it can be checked without a device.
 * init_rows: the k rows the initial centroids are copied from (splitmix64-driven partial Fisher-Yates).
 * plan_update: one update's work on an assignment pass's result -- cluster sizes; every empty cluster (ascending) takes
 * the farthest row (ties: lowest row) whose cluster keeps another row, `assign` is modified accordingly; stable counting
 * sort: row_sorted[cluster_off[j] .. cluster_off[j+1]) = rows of cluster j, ascending. */
int b200nn_kmeans_init_rows(size_t n, int k, uint64_t seed, int32_t* rows_out /*[k]*/);
int b200nn_kmeans_plan_update(size_t n, int k, int32_t* assign /*[n] in/out*/, const float* dist /*[n]*/, int32_t* count /*[k]*/,
                              int32_t* row_sorted /*[n]*/, int64_t* cluster_off /*[k+1]*/);
/* x_raw [n,D] un-reordered rows (perm, if given, is applied first: TrainPQ::LoadFeatureSample :80,98,112).
 * K >= 1: coarse [K,D] trained; K == 0: no coarse quantizer, coarse receives ONE all-zero centroid (the flat-ADC model).
 * codebooks [M,ksub,D/M].  mse_out (NULL or 1+M doubles): mean squared quantisation error of the coarse stage and of
 * each sub-space.  Sub-space m trains with seed + 1 + m. */
int b200nn_pq_train(b200nn_ctx_t ctx, const float* x_raw, size_t n, int D, int K, int M, int ksub, const int32_t* perm /*or NULL*/,
                    int max_iter, uint64_t seed, float* coarse, float* codebooks, double* mse_out);
/* host-only: the OPQ model file, reorder tail written as int32[D] (identity when perm is NULL);
 * the reference writer emits sizeof(int)*D bytes of a long-int array there (train_PQ_codebook.cpp:286). */
int b200nn_pq_write_model(const char* path, int D, int K, int M, int ksub, const float* coarse, const float* codebooks,
                          const int32_t* perm);

#ifdef __cplusplus
}
#endif
#endif /* B200NN_H */
