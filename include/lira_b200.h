/*
 * liblira_b200 -- C ABI of the B200-native LIRA query phase and ground-truth path.
 *
 * The reference (qfshen23/LIRA-ANN-search) has no plugin / FFI layer: its hot path is reached
 * through Python call shapes (LIRA_smallscale.py, LIRA_largescale.py, utils.py, model_probing.py)
 * that bottom out in faiss-cpu, and through two C++ programs (search.cpp, compute_knn.cpp).
 * Every entry point below names the reference interface it stands in for (file:line into the
 * reference tree). INTEGRATION.md shows the ctypes / C++ stubs a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer it passes;
 *   - functions return 0 on success, non-zero on failure, and lira_last_error() then holds a
 *     message (the reference throws std::runtime_error -> "[Error] ..." + exit 1, search.cpp:552-555);
 *   - vectors are fp32 row-major; ids are int32 inside the index (index.py:166-167, search.cpp:274)
 *     and int64 in results (faiss idx_t; numpy default int, LIRA_smallscale.py:154);
 *   - metric: LIRA_METRIC_L2 = squared L2, no sqrt (faiss IndexFlatL2; search.cpp:253-260),
 *             LIRA_METRIC_IP = inner product, larger is better (IndexFlatIP; search.cpp:263-269);
 *   - one caller thread per handle; each handle owns one CUDA stream on its device;
 *   - names ending in _dev take DEVICE pointers (and a cudaStream_t passed as void*; NULL = the
 *     handle's own stream) and do not synchronise; all other functions take HOST pointers, copy
 *     in/out inside the call and return after the results have landed.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef LIRA_B200_H_
#define LIRA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIRA_METRIC_L2 0
#define LIRA_METRIC_IP 1

/* partition-selection modes (a5) */
#define LIRA_SELECT_GT 0        /* score >  value          LIRA_smallscale.py:206, LIRA_largescale.py:163 */
#define LIRA_SELECT_GE_ARGMAX 1 /* score >= value, argmax if none   search.cpp:448-466                    */
#define LIRA_SELECT_TOPN 2      /* the (int)value best scores       utils.py:512 (all_outputs[q].topk)    */

typedef struct lira_index lira_index_t; /* inverted lists resident in HBM       */
typedef struct lira_model lira_model_t; /* probing model + centroids + scaler   */
typedef struct lira_knn_index lira_knn_t;    /* base vectors resident for exact kNN  */

/* ---- library -------------------------------------------------------------------------- */
const char* lira_last_error(void);
int lira_version(void);
int lira_device_count(void); /* 0 when no CUDA device / driver is present */

/* ---- a6: inverted lists ------------------------------------------------------------------
 * utils.create_flat_indexes / create_inner_indexes (utils.py:407-429): list b holds
 * x_d[cluster_ids[b]] in that order, i.e. local index i of list b <-> list_ids[list_offsets[b]+i].
 * base[N,d] host; list_offsets[B+1] (int64), list_ids[E] (int32, values in [0,N)). The vectors are
 * gathered into list order on the device; `base` is not referenced after the call. */
int lira_index_create(const float* base, int64_t N, int d, const int64_t* list_offsets,
                      const int32_t* list_ids, int B, int metric, int device, lira_index_t** out);
/* search.cpp:368-403: build the lists from data_2_bkt[N,n_mul] (int32, -1 = empty slot); ids of a
 * bucket are sorted and unique. Errors with "bucket id out of range." like search.cpp:375-377. */
int lira_index_create_from_assign(const float* base, int64_t N, int d, const int32_t* data_2_bkt,
                                  int n_mul, int B, int metric, int device, lira_index_t** out);
/* adopt lists already laid out on the device: vecs[E, ld] fp32 (ld % 4 == 0, 16-byte aligned),
 * ids[E] int32, offsets host int64[B+1]. The index does NOT own or copy vecs/ids. */
int lira_index_create_dev(const float* d_vecs, int64_t ld, int d, const int64_t* list_offsets,
                          const int32_t* d_ids, int B, int metric, int device, lira_index_t** out);
int lira_index_free(lira_index_t* h);
int64_t lira_index_ntotal(const lira_index_t* h, int list); /* faiss Index.ntotal of list b (LIRA_smallscale.py:171); list < 0: all entries */
int lira_index_nlist(const lira_index_t* h);
int lira_index_dim(const lira_index_t* h);

/* ---- a7: per-list flat search (the faiss IndexFlat surface) ------------------------------
 * inner_indexes[b].search(q[nq,d], k) -> (D[nq,k] fp32, I[nq,k] int64 LOCAL positions), best
 * first, -1 / +-inf padded when the list holds fewer than k vectors (LIRA_smallscale.py:168). */
int lira_index_list_search(lira_index_t* h, int list, const float* q, int64_t nq, int k, float* D,
                           int64_t* I);
/* get_cmp_recall (LIRA_smallscale.py:145-174, LIRA_largescale.py:120-149) in one call: for EVERY
 * (query, list) pair the top-k inside the list mapped to global ids.
 * found[Q,B,k] int64 (shorter non-empty lists repeat their last id, empty lists stay -1, exactly as
 * the reference's xd_id_bid[idx] indexing does); cmp[Q,B] int64 = list size. */
int lira_scan_all_pairs(lira_index_t* h, const float* q, int64_t Q, int k, int64_t* found,
                        int64_t* cmp);

/* ---- a10: online search with explicit probe sets (search.cpp:468-514) --------------------
 * probe_offsets[Q+1] / probe_ids[P]: CSR of the probed lists per query.
 * dedup = 1: an id stored in several probed lists counts once before selection (Python recall
 *            semantics, LIRA_smallscale.py:210-214); dedup = 0: search.cpp:499-513 as shipped.
 * D[Q,k] metric value, I[Q,k] global ids (-1 padded), cmp[Q] (may be NULL) = sum of probed sizes.
 * A probed list id outside [0, B) is an error in lira_search (checked on the host). lira_search_dev cannot look at
 * device memory before the launch: it skips such ids (they count neither in cmp nor in the results) and reports
 * "probed list id out of range" from batches of >= 256 queries on the tensor-core scan, whose completion it waits for.
 * Device inputs: the row stride may exceed d only if the padding columns d..ld-1 are zero (vectors and queries). */
int lira_search(lira_index_t* h, const float* q, int64_t Q, const int64_t* probe_offsets,
                const int32_t* probe_ids, int k, int dedup, float* D, int64_t* I, int64_t* cmp);
int lira_search_dev(lira_index_t* h, const float* d_q, int64_t ldq, int64_t Q,
                    const int64_t* d_probe_offsets, const int32_t* d_probe_ids, int64_t P, int k,
                    int dedup, float* d_D, int64_t* d_I, int64_t* d_cmp, void* stream);

/* ---- a1-a5: probing model --------------------------------------------------------------
 * centroids[B,d]; scaler mean/scale[B] (utils.py:171-175 -> search.cpp:322-329; scale==0 acts as 1);
 * weights: the 12 tensors of MLP_2_Input in state_dict order (model_probing.py:12-31):
 *   distance_net.0 W[128,B] b[128], distance_net.2 W[64,128] b[64], vector_net.0 W[128,d] b[128],
 *   vector_net.2 W[64,128] b[64], fc.0 W[128,128] b[128], fc.2 W[Bout,128] b[Bout]; Bout == B. */
int lira_model_create(const float* centroids, const float* scaler_mean, const float* scaler_scale,
                      int B, int d, const float* const weights[12], int device, lira_model_t** out);
int lira_model_free(lira_model_t* m);
/* The forward pass (centroid features + the six Linear layers) has two implementations: fp32 CUDA cores and
 * tcgen05 tensor cores with error-compensated TF32 (every operand carried as an exact hi + lo pair, three MMAs
 * per product: fp32-level accuracy). Default: tensor cores. 0 pins the CUDA-core kernels. */
int lira_model_set_use_tensor_cores(lira_model_t* m, int enable);
/* get_dist_cid + StandardScaler.transform (utils.py:98-118, 142-167) == compute_l2_to_centroids +
 * standardize_distances (search.cpp:220-250): out[Q,B] fp32. mean/scale NULL -> raw distances.
 * Always Euclidean, also for inner-product datasets (utils.py:115). */
int lira_centroid_features(const float* q, int64_t Q, const float* centroids, int B, int d,
                           const float* mean, const float* scale, int device, float* out);
/* model(x_dist, x_vec) for a batch of raw queries: all_outputs[Q,B] of model_evaluate /
 * model_infer (model_probing.py:86-156); feats (may be NULL) receives the scaled distances. */
int lira_model_scores(lira_model_t* m, const float* q, int64_t Q, float* scores, float* feats);

/* ---- the whole query phase (search.cpp:421-517 per batch) --------------------------------
 * features -> MLP -> select(mode,value) -> grouped scan -> merge.
 * nprobe[Q] / cmp[Q] may be NULL. */
int lira_probe_search(lira_index_t* h, lira_model_t* m, const float* q, int64_t Q, int mode,
                      double value, int k, int dedup, float* D, int64_t* I, int32_t* nprobe,
                      int64_t* cmp);
int lira_probe_search_dev(lira_index_t* h, lira_model_t* m, const float* d_q, int64_t ldq, int64_t Q,
                          int mode, double value, int k, int dedup, float* d_D, int64_t* d_I,
                          int32_t* d_nprobe, int64_t* d_cmp, void* stream);
/* The same call split in two so that consecutive batches pipeline (search.cpp answers its queries one after the other,
 * search.cpp:421-517; a server answers a stream of batches):
 *   lira_probe_search_enqueue_dev  launches the whole query phase of the batch on the stream and returns without waiting;
 *   lira_index_finish              waits for every batch enqueued on the handle and checks their status words. The fused
 *                                  tensor-core flow is optimistic (fp16-exact batch, at most nprobe_cap partitions per query,
 *                                  no candidate-region overflow); a batch for which that did not hold is answered again here
 *                                  through the checked path, into the same output buffers. Results are valid after finish.
 * Host-buffer form with four staging slots (slot in [0, 4)): submit(slot) copies the queries in (pinned or pageable source) on a copy
 * stream, enqueues the batch and the copy of its results into pinned memory on a second copy stream; wait(slot) returns
 * the results in the caller's arrays. With submit(i+1) issued before wait(i), the copies of one batch overlap the
 * kernels of the other. A pageable query array is staged through a pinned buffer of the handle; a pinned one is uploaded
 * from directly unless the first timed batches show its upload taking more than 0.7 of the time of the batch's kernels (pages pinned by
 * another allocator can upload at half the rate), in which case it is staged too (LIRA_STAGE_PINNED=0 / 1 fixes the choice). */
int lira_probe_search_enqueue_dev(lira_index_t* h, lira_model_t* m, const float* d_q, int64_t ldq, int64_t Q,
                                  int mode, double value, int k, int dedup, float* d_D, int64_t* d_I,
                                  int32_t* d_nprobe, int64_t* d_cmp, void* stream);
int lira_index_finish(lira_index_t* h);
int lira_probe_search_submit(lira_index_t* h, lira_model_t* m, const float* q, int64_t Q, int mode,
                             double value, int k, int dedup, int slot);
int lira_probe_search_wait(lira_index_t* h, int slot, float* D, int64_t* I, int32_t* nprobe, int64_t* cmp);
/* search with scores already on the device (threshold replay of query_tuning, LIRA_smallscale.py:199-220) */
int lira_select_search_dev(lira_index_t* h, const float* d_scores, int64_t lds, const float* d_q,
                           int64_t ldq, int64_t Q, int mode, double value, int k, int dedup, float* d_D,
                           int64_t* d_I, int32_t* d_nprobe, int64_t* d_cmp, void* stream);

/* ---- a11: exact k-nearest neighbours -----------------------------------------------------
 * compute_knn.cpp:208-259 / utils.compute_data_knn fallback (utils.py:286-319) /
 * LIRA_largescale.py:221-231: brute-force top-k of query[Q,d] against base[N,d]; ties to the lower
 * base id. The caller drops column 0 for self-kNN as the reference does (compute_knn.cpp:254-259). */
int lira_knn(const float* base, int64_t N, const float* query, int64_t Q, int d, int k, int metric,
             int device, float* D, int64_t* I);
/* The same with the base resident behind a handle (faiss: index.add(x) once, index.search(batch, k + 1) per batch of
 * 10 000 rows -- compute_knn.cpp:208-244, utils.py:293-310, LIRA_largescale.py:225-229): the upload, the row norms and the
 * fp16 shadow copy are made once at create, every search only moves its query batch. */
int lira_knn_create(const float* base, int64_t N, int d, int metric, int device, lira_knn_t** out);
int lira_knn_search(lira_knn_t* h, const float* query, int64_t Q, int k, float* D, int64_t* I);
/* The same over DEVICE memory: the base d_base[N, ld] (ld % 4 == 0, padding columns zero, 16-byte aligned) is adopted, not
 * copied, and must outlive the handle; queries and results are device pointers (the 10 000-row batches of
 * compute_knn.cpp:228-244 without a host round trip; also the K-Means assignment step of utils.py:325, where the base is
 * the centroid table and k = 1). */
int lira_knn_create_dev(const float* d_base, int64_t ld, int64_t N, int d, int metric, int device, lira_knn_t** out);
int lira_knn_search_dev(lira_knn_t* h, const float* d_query, int64_t ldq, int64_t Q, int k, float* d_D, int64_t* d_I,
                        void* stream);
int lira_knn_free(lira_knn_t* h);
int64_t lira_knn_ntotal(const lira_knn_t* h);
int lira_knn_set_use_tensor_cores(lira_knn_t* h, int enable);
int lira_knn_last_path(const lira_knn_t* h); /* last search: 0 CUDA cores, 1 tensor cores, 2 some batches on each */
int lira_knn_last_redo(const lira_knn_t* h); /* queries of the last search answered again by the exact CUDA-core scan */
int lira_knn_last_scan_kind(const lira_knn_t* h); /* kernel family of the last batch, see lira_index_last_scan_kind */

/* ---- a12: IVF-approximate self-kNN (compute_knn.cpp:158-203) ------------------------------------------------------
 * faiss::IndexIVFFlat(quantizer = IndexFlatL2(d), d, nlist): train (K-Means, lira_kmeans_train), add (every vector to the list
 * of its nearest centroid), search with `nprobe`: the k best of each base vector inside its nprobe nearest lists, exact
 * distances, best first (-1 padded). D[N,k], I[N,k] host. The caller drops column 0 (the vector itself) as
 * compute_knn.cpp:254-259 does. */
int lira_knn_ivf(const float* base, int64_t N, int d, int k, int nlist, int nprobe, uint64_t seed, int device, float* D,
                 int64_t* I);

/* ---- f2 / f3: index build (the callers in front of the query path) ----------------------------------------------
 * faiss.Kmeans(d, B, niter=20).train(x) as build_kmeans_index uses it (utils.py:321-324): at most 256 points per centroid
 * take part (random subset, `seed`), centroids start from B random points (or init_centroids[B,d] when given), `niter` Lloyd
 * iterations, an empty cluster is re-seeded by splitting the largest one. Assignment = exact nearest centroid on the kNN path
 * (ties to the lower id), update = segmented mean on the device. centroids_out[B,d] host. The assignment of the full data
 * (kmeans.index.search(x, 1): utils.py:325, LIRA_largescale.py:294) is lira_knn_* with the centroid table as base, k = 1.
 * _dev: rows d_x[n, ld] on the device (ld % 4 == 0, padding zero); d_centroids[B, ldc] is the result and, with init_given,
 * the start; otherwise d_init_rows[B] (int64) names the rows of d_x the centroids start from. */
int lira_kmeans_train(const float* x, int64_t n, int d, int B, int niter, uint64_t seed, const float* init_centroids,
                      int device, float* centroids_out);
int lira_kmeans_train_dev(const float* d_x, int64_t ld, int64_t n, int d, int B, int niter, int init_given,
                          float* d_centroids, int64_t ldc, const int64_t* d_init_rows, int device, void* stream);
/* get_dist_cid (+ StandardScaler.transform) with everything on the device: d_out[Q, ldo] (ldo % 4 == 0). */
int lira_centroid_features_dev(const float* d_q, int64_t ldq, int64_t Q, const float* d_centroids, int64_t ldc, int B,
                               int d, const float* d_mean, const float* d_scale, float* d_out, int64_t ldo, int device,
                               void* stream);
/* StandardScaler.fit / partial_fit on the centroid distances of n rows (utils.py:133-168, and per 1 000 000-row batch in
 * get_scaled_dist_data, utils.py:182-215): fp64 mean and variance per partition; the [n, B] matrix never leaves the device.
 * mean_out[B], var_out[B] host. Waits for the stream. */
int lira_feature_stats_dev(const float* d_x, int64_t ld, int64_t n, const float* d_centroids, int64_t ldc, int B, int d,
                           double* mean_out, double* var_out, int device, void* stream);
/* mul_partition_by_model (LIRA_smallscale.py:77-97, LIRA_largescale.py:51-72), one warp per point: row i of d_score[n_rows,
 * lds] (the probing model's outputs) belongs to point d_points[i] (NULL: first + i). With the partitions ranked by score
 * (equal scores: lower id first), n_eff = #(score > sigma), n_act = min(n_mul - 1, n_eff) and loc = the rank of the point's
 * current partition d_data_2_bkt[t, 0]:  loc >= n_act -> the n_act best go to columns 1..n_act;  else n_eff == n_act -> they
 * replace columns 0..n_act-1;  else the n_act + 1 best replace columns 0..n_act. d_data_2_bkt[N, n_mul] int32 is updated in
 * place; d_added[N, n_mul] (caller presets -1) receives the partitions that newly hold the point, for the caller's
 * cluster_ids / cluster_cnts bookkeeping. 2 <= n_mul <= 8. */
int lira_mul_partition_dev(const float* d_score, int64_t lds, int64_t n_rows, int B, float sigma, const int64_t* d_points,
                           int64_t first, int n_mul, int32_t* d_data_2_bkt, int32_t* d_added, int device, void* stream);

/* ---- multi-GPU merge (e): per-rank top-k lists -> global top-k with id de-duplication ----
 * d_keys_in[R, Q, k]: rank-major gathered (score,id) lists as produced by lira_*_topk_keys_dev /
 * lira_pack_keys_dev; output as lira_search. */
int lira_pack_keys_dev(const float* d_D, const int64_t* d_I, int64_t n, int metric, uint64_t* d_keys,
                       int device, void* stream);
int lira_merge_ranks_dev(const uint64_t* d_keys_in, int R, int64_t Q, int k, int metric, int dedup,
                         float* d_D, int64_t* d_I, int device, void* stream);

/* ---- instrumentation -------------------------------------------------------------------
 * kernels launched by this library since load (bench.py's gpu_launches), and the device time of
 * the last scan kernel / last whole search in milliseconds (CUDA events on the handle's stream). */
int64_t lira_launch_count(void);
int lira_index_last_timing(const lira_index_t* h, float* scan_ms, float* total_ms, int64_t* scan_bytes,
                           int64_t* scan_pairs);
int lira_index_set_timing(lira_index_t* h, int enable);
/* device time of ALL kernels of the last scan -- seed pass + filter pass + refine on the tensor-core path (the scan as
 * SURVEY.md 8d defines it: probed entries in, de-duplicated top-k out); equals scan_ms on the CUDA-core path */
int lira_index_last_scan_total_ms(const lira_index_t* h, float* scan_total_ms);
/* The online search (lira_search*, lira_probe_search*, lira_select_search*) has two implementations of the list
 * scan with the same results: fp32 CUDA cores (always valid) and tcgen05 tensor cores, taken for batches of
 * >= 256 queries and d <= 1024 (256 < d: both operands stream) in one of two modes decided when the index is created:
 *   mode 1 (exact): every stored value and every query value is an integer of at most 11 bits and |x|^2 < 2^22;
 *           the kernel streams an fp16 copy of the rows, products and fp32 sums are exact, both paths return
 *           identical bits;
 *   mode 2 (approximate filter + exact re-rank, k <= 16): any other finite data; the kernel streams fp16(sigma v),
 *           keeps every candidate whose rounded score is within a rigorous error margin of the running bound, and
 *           every survivor is scored again exactly in fp32 from the original rows, so ids and distances agree with
 *           the CUDA-core scan up to fp32 summation order (ties / 1e-6 relative).
 * set_use_tensor_cores(0) pins the CUDA-core scan; last_path reports which one ran (0 CUDA cores, 1 tensor
 * cores); tensor_core_eligible / tensor_core_mode tell whether and how the stored vectors qualify (0, 1, 2). */
int lira_index_set_use_tensor_cores(lira_index_t* h, int enable);
int lira_index_last_path(const lira_index_t* h);
int lira_index_last_redo(const lira_index_t* h); /* queries of the last tensor-core batch whose candidate buffer
                                                    overflowed and that were answered by the CUDA-core scan */
int lira_index_tensor_core_eligible(const lira_index_t* h);
int lira_index_tensor_core_mode(const lira_index_t* h);
/* Byte-valued data (mode 1 with every value an integer in [0, 255], d <= 128) additionally qualifies for the integer
 * tensor-core scan (tcgen05.mma kind::i8 over a one-byte-per-component copy of the rows, int32 accumulators: exact).
 * It is what exhaustive probe sets (lira_knn*, compute_knn.cpp:208-259) run by default; threshold / top-n / explicit probe
 * sets use it when LIRA_U8_SEARCH=1 is set in the environment. last_scan_kind: 0 CUDA cores, 1 fp16 tensor-core scan,
 * 2 byte tensor-core scan. */
int lira_index_byte_scan_eligible(const lira_index_t* h);
int lira_index_last_scan_kind(const lira_index_t* h);

#ifdef __cplusplus
}
#endif
#endif /* LIRA_B200_H_ */
