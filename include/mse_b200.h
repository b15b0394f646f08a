/*
 * mse_b200.h -- C ABI of libmse_b200.so: the B200-native embed-and-search hot path of
 * osmarks/meme-search-engine.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository root).  Host code in the reference is Rust; INTEGRATION.md shows the `extern "C"`
 * block + safe wrappers a maintainer would add.
 *
 * Conventions
 *   - return 0 on success, a negative MSE_ERR_* otherwise; mse_last_error() gives a thread-local message
 *   - the caller owns every buffer it passes; the library owns what is behind its opaque handles
 *   - one in-flight call per handle (the reference has one inference thread / one &mut Scratch per thread);
 *     handles on different devices are independent
 *   - fixed-point scores are i64 = trunc(f32 * 2^32) exactly as diskann/src/vector.rs:46,249-250,408-416
 *   - "_dev" variants take device pointers and a cudaStream_t (as void*); they return with the work queued on that stream.
 *     Exceptions are stated per entry (mse_search_flat_dev reads one status word back before returning; *_check synchronise)
 *   - there is NO CPU fallback: every compute entry fails with MSE_ERR_CUDA when no sm_100 device is usable
 */
#ifndef MSE_B200_H
#define MSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSE_OK 0
#define MSE_ERR_INVALID (-1)     /* bad argument */
#define MSE_ERR_CUDA (-2)        /* CUDA runtime/driver failure, or no usable device */
#define MSE_ERR_OOM (-3)         /* device or host allocation failed */
#define MSE_ERR_UNSUPPORTED (-4) /* shape outside what the kernels were built for */
#define MSE_ERR_STATE (-5)       /* handle is missing something the call needs (graph, codes, ...) */

#define MSE_ID_NONE 0xFFFFFFFFu /* padding id when k > ntotal (FAISS returns -1: src/main.rs:908) */

const char *mse_last_error(void);
/* library / device facts: writes up to cap bytes of a JSON object (sm count, device name, kernels built) */
int mse_device_info(int device, char *json_out, size_t cap);
/* number of kernel launches this library has issued in this process (bench.py's gpu_launches) */
uint64_t mse_launch_count(void);

/* =====================================================================================
 * Distance kernel -- diskann/src/vector.rs:192-306 (fast_dot / fast_dot_noprefetch)
 * ===================================================================================== */

/* scores[i] = fast_dot(query, rows[row_ids[i]]) bit-identically to the reference's AVX2 routine
 * (32 partial sums in fp32 FMA, its reduction tree, trunc(x*2^32)).  d % 64 == 0.  Host pointers.
 * row_ids == NULL means rows 0..n_ids-1. */
int mse_fast_dot_batch(int device, const uint16_t *query_f16, const uint16_t *rows_f16, uint64_t n_rows, uint32_t d,
                       const uint32_t *row_ids, uint64_t n_ids, int64_t *scores);

/* =====================================================================================
 * Index handle -- VectorList (diskann/src/vector.rs:118-186) + IndexGraph (diskann/src/lib.rs:16-39)
 *                 + FAISS IndexScalarQuantizer(QT_fp16, InnerProduct) (src/main.rs:822)
 * One handle = one id-range shard resident in one GPU's HBM.
 * ===================================================================================== */

typedef struct mse_index mse_index;

/* VectorList::from_f16s / ScalarQuantizerIndexImpl::new.  x_f16 may be NULL with n == 0 (empty index).
 * id_base is added to every returned id (the shard's first global vector id). */
int mse_index_create(const uint16_t *x_f16, uint64_t n, uint32_t d, int device, uint32_t id_base, mse_index **out);
/* faiss Index::add(&[f32]) (src/main.rs:858,892): rows are rounded to fp16 (RNE) as QT_fp16 stores them */
int mse_index_add(mse_index *ix, const float *x_f32, uint64_t n);
/* VectorList::push for already-fp16 rows */
int mse_index_add_f16(mse_index *ix, const uint16_t *x_f16, uint64_t n);
/* same, source rows already in device memory */
int mse_index_add_f16_dev(mse_index *ix, const uint16_t *d_x_f16, uint64_t n, void *stream);
/* Vec::with_capacity for the row store (generate_index_shard.rs:52-53 pre-sizes with -N): no reallocation below `rows` */
int mse_index_reserve(mse_index *ix, uint64_t rows);
/* faiss Index::ntotal (src/main.rs:1015,1053) / VectorList::len */
uint64_t mse_index_ntotal(const mse_index *ix);
uint32_t mse_index_dim(const mse_index *ix);
int mse_index_device(const mse_index *ix);
/* device address of the fp16 rows (for zero-copy interop with the caller's own CUDA code) */
const uint16_t *mse_index_vectors_dev(const mse_index *ix);
void mse_index_destroy(mse_index *ix);

/* =====================================================================================
 * Flat search -- faiss Index::search on IndexScalarQuantizer(QT_fp16, IP)  (src/main.rs:900; mse.py:79)
 * score_i = sum_d q_d * f32(x_id)  ranked by (score desc, id asc); q stays f32 (never rounded to fp16).
 * ids/scores are [nq][k]; slots past ntotal hold MSE_ID_NONE / -inf.
 * Results are exact: candidates found by the tensor-core pass are re-scored in fp64 and the cut is
 * certified against the rounding bound of the fp16 pass; uncertified queries are re-run on the exact scan.
 * ===================================================================================== */
int mse_search_flat(mse_index *ix, const float *q, uint32_t nq, uint32_t k, uint32_t *ids, float *scores);
/* Device pointers; queues the whole search on `stream` and returns without synchronising.  The results in d_ids / d_scores
 * are final once mse_search_flat_check has returned: it synchronises the stream, reads the search's status word and re-runs
 * the (rare) queries whose cut could not be certified, or whose candidate buffer overflowed, on the exact scan.
 * *repaired (optional) = number of queries it had to re-run. */
int mse_search_flat_dev(mse_index *ix, const float *d_q, uint32_t nq, uint32_t k, uint32_t *d_ids, float *d_scores,
                        void *stream);
int mse_search_flat_check(mse_index *ix, uint32_t *repaired);
/* Query assembly on the device -- get_total_embedding (src/common.rs:215-274): d_q[i][:] += weight * f32(d_e_f16[i][:]) for nq rows of
 * d values (`*total += *value * weight`, :270).  The query stays f32 and is NOT renormalised.  Zero d_q first. */
int mse_query_accumulate_f16_dev(int device, const uint16_t *d_e_f16, float weight, float *d_q, uint32_t nq, uint32_t d, void *stream);
/* statistics of the last mse_search_flat* call on this handle:
 * out[0]=tensor-path queries, out[1]=exact-scan queries, out[2]=queries escalated after a failed certificate,
 * out[3]=candidate-buffer overflows, out[4]=kernel launches, out[5]=chunks,
 * out[6]=nanoseconds spent inside the scoring kernels (CUDA events; only with profiling on), out[7]=scoring launches */
int mse_search_flat_stats(const mse_index *ix, uint64_t out[8]);
/* bracket every scoring-kernel launch with CUDA events on the launching stream (feeds out[6], out[7]) */
int mse_search_flat_profile(mse_index *ix, int enable);
/* force a path for testing: 0 = auto, 1 = exact fp64 scan only, 2 = tensor-core pass (+ certified rerank) only */
int mse_search_flat_set_mode(mse_index *ix, int mode);

/* k-way merge of per-shard top-k lists (the step after the all-gather of SURVEY 8e): lists are
 * [n_shards][nq][k] (ids, scores), output [nq][k] by (score desc, id asc). Device pointers. */
int mse_merge_topk_dev(int device, const uint32_t *d_ids, const float *d_scores, uint32_t n_shards, uint32_t nq,
                       uint32_t k, uint32_t *d_out_ids, float *d_out_scores, void *stream);

/* =====================================================================================
 * Id-range sharding over the GPUs of one box (BASELINE north_star: "the index partitions by vector-id range across
 * the 8 GPUs of one box with a single NCCL all-gather of per-shard top-k over NVLink").  One process (or thread) per GPU:
 * every rank creates its shard with mse_index_create(..., device, id_base = first global id of the shard), joins the group,
 * and calls the *_sharded_dev searches with the same replicated queries.  Each call queues, on the caller's stream:
 * local search -> top-k written as packed (score, global id) entries straight into this rank's slot of the NCCL-registered
 * gather buffer -> ONE ncclAllGather (in place) -> k-way merge by (score desc, id asc) on every rank.  Ids are global and
 * the order is total, so the merged result is independent of the number of shards.  No host synchronisation.
 * NCCL is bound at run time (the libnccl.so.2 already loaded in the process, else $MSE_NCCL_LIB, else the system's); a
 * one-rank group needs no NCCL at all.  The reference has no counterpart (it shards on disk: dump_processor.rs:134,438-461).
 * ===================================================================================== */
#define MSE_SHARD_ID_BYTES 128
typedef struct mse_shard_group mse_shard_group;

/* [lo, hi) = the global ids of shard `rank` of `n_ranks` over n_total vectors */
int mse_shard_range(uint64_t n_total, int n_ranks, int rank, uint64_t *lo, uint64_t *hi);
/* rank 0 draws the group id (ncclGetUniqueId) and hands the bytes to the other ranks by any means (file, socket, MPI, ...) */
int mse_shard_group_unique_id(uint8_t out[MSE_SHARD_ID_BYTES]);
/* collective over the n_ranks callers (ncclCommInitRank).  unique_id may be NULL when n_ranks == 1. */
int mse_shard_group_create(const uint8_t *unique_id, int n_ranks, int rank, int device, mse_shard_group **out);
/* out = n_ranks, rank, device, NCCL version (0 for a one-rank group) */
int mse_shard_group_info(const mse_shard_group *g, int32_t out[4]);
/* number of all-gathers this group has issued (bench.py's collective count) */
uint64_t mse_shard_group_gathers(const mse_shard_group *g);
void mse_shard_group_destroy(mse_shard_group *g);

/* mse_search_flat_dev over all shards: d_ids (global) / d_scores [nq][k], identical on every rank.  The certificate status
 * of every shard travels with its list, so mse_search_sharded_check -- which every rank must call, like the search itself --
 * decides without a further exchange whether a repair round (exact re-run on the flagged shards, second gather + merge) is needed. */
int mse_search_flat_sharded_dev(mse_shard_group *g, mse_index *shard, const float *d_q, uint32_t nq, uint32_t k, uint32_t *d_ids,
                                float *d_scores, void *stream);
int mse_search_sharded_check(mse_shard_group *g, uint32_t *repaired);
/* greedy_search (lib.rs:183-211) on every shard's own Vamana graph from its entry point `start` (a local row of this shard),
 * best k of each shard's candidate list merged: d_ids (global) / d_scores (i64 fixed point) [nq][k].  d_distances (optional)
 * [nq] = this shard's GreedySearchCounters.distances.  mse_search_graph_check(shard, nq) reports overflows as usual. */
int mse_search_graph_sharded_dev(mse_shard_group *g, mse_index *shard, const uint16_t *d_q_f16, uint32_t nq, uint32_t L, uint32_t start,
                                 uint32_t k, uint32_t *d_ids, int64_t *d_scores, uint64_t *d_distances, void *stream);
/* mse_search_beam_dev (query_disk_index.rs:144-212: compressed scores for the frontier, exact scores for expanded nodes)
 * on every shard, best k expanded nodes of each shard merged.  d_cmps / d_pq_cmps (optional) [nq] = this shard's counters. */
int mse_search_beam_sharded_dev(mse_shard_group *g, mse_index *shard, const uint16_t *d_q_f16, const float *d_luts, const float *d_qtm,
                                uint32_t rabitq_output_dims, uint32_t rabitq_n_dims, const float *d_desc_scales, uint32_t nq, uint32_t L, uint32_t W,
                                uint32_t start, uint32_t n_centroids, uint32_t k, uint32_t *d_ids, int64_t *d_scores, uint64_t *d_cmps,
                                uint64_t *d_pq_cmps, void *stream);

/* =====================================================================================
 * Graph index -- the diskann crate's build/search entry points (diskann/src/lib.rs) and the packed-index search of
 * src/query_disk_index.rs, batched over queries.  All of it is bit-exact against the CPU restatement given the same
 * graph: ids, i64 scores and distance counters.
 * ===================================================================================== */

/* IndexBuildConfig (diskann/src/lib.rs:41-51); alpha / query_alpha are fixed-point x 2^16 */
typedef struct mse_build_config {
    uint64_t r, l, maxc;
    int64_t alpha;
    int32_t saturate_graph;
    uint32_t query_breakpoint; /* ids >= this are query nodes (OOD-DiskANN); 0xFFFFFFFF = none */
    uint64_t max_add_per_stitch_iter;
    int64_t query_alpha;
} mse_build_config;

/* IndexGraph (lib.rs:16-39) as fixed-stride adjacency adj[n][stride] + deg[n]; `N.shard.bin` + offsets
 * (generate_index_shard.rs:143-153) is the CSR form of the same lists */
int mse_index_set_graph(mse_index *ix, const uint32_t *adj, const uint32_t *deg, uint32_t stride);
int mse_index_get_graph(const mse_index *ix, uint32_t *adj, uint32_t *deg, uint32_t *stride_out);
/* index.pq-codes.bin / index.descriptor-codes.bin (query_disk_index.rs:686-709) and `url.len() > 0` flags (:172) */
int mse_index_set_pq_codes(mse_index *ix, const uint8_t *codes, uint32_t code_size);
int mse_index_set_descriptors(mse_index *ix, const uint8_t *desc, uint32_t n_desc, const uint8_t *has_url);

/* greedy_search (lib.rs:183-211) for nq queries: results are scratch.neighbour_buffer.ids (+ scores) [nq][L], best first,
 * MSE_ID_NONE past len[q]; distances[q] = GreedySearchCounters.distances.  starts may be NULL (all start at `start`).
 * visited_* (optional, [nq][visited_cap]) return scratch.visited_list in evaluation order. */
int mse_search_graph(mse_index *ix, const uint16_t *q_f16, uint32_t nq, uint32_t L, const uint32_t *starts, uint32_t start,
                     int base_vectors_only, uint32_t query_breakpoint, uint32_t *ids, int64_t *scores, uint32_t *len,
                     uint64_t *distances, uint32_t *visited_ids, int64_t *visited_scores, uint32_t *visited_len,
                     uint32_t visited_cap);
/* the same with every buffer already in HBM (device pointers), asynchronous on `stream`; no visited lists.  The
 * per-query visited-set overflow status is read back by mse_search_graph_check (which synchronises the device). */
int mse_search_graph_dev(mse_index *ix, const uint16_t *d_q_f16, uint32_t nq, uint32_t L, const uint32_t *d_starts, uint32_t start,
                         int base_vectors_only, uint32_t query_breakpoint, uint32_t *d_ids, int64_t *d_scores, uint32_t *d_len,
                         uint64_t *d_distances, void *stream);
int mse_search_graph_check(mse_index *ix, uint32_t nq);
/* schedule of greedy_search on the GPU: 0 automatic, 1 one CTA per query (small batches), 2 one warp per query (large
 * batches).  Both replay lib.rs:183-211 insert for insert; results are identical. */
int mse_search_graph_set_mode(int mode);
/* greedy_search of query_disk_index.rs:144-212 (beam W, PQ ADC for candidates, exact score + descriptor bias for expanded
 * nodes).  luts [nq][M*n_centroids] from mse_pq_preprocess_query; desc_scales [nq][n_desc] or NULL.  out_* [nq][out_cap]:
 * expanded nodes in visit order with their exact scores (the caller sorts, :529); cmps as :148.  pq_cmps counts every candidate
 * once; the reference clears its pre-buffer per beam iteration (:157), not per expanded node, and so re-scores (and counts, :149)
 * the earlier nodes' candidates again when W > 1 -- same results, larger counter.  A query that expands more than out_cap nodes
 * fails the call (MSE_ERR_UNSUPPORTED) instead of returning a truncated list.  start / starts[] are LOCAL row numbers of this
 * shard (IndexHeader.shards[i].medioid is global: subtract the shard's id_base); out-of-range values are MSE_ERR_INVALID. */
int mse_search_beam(mse_index *ix, const uint16_t *q_f16, const float *luts, const float *desc_scales, uint32_t nq, uint32_t L,
                    uint32_t W, const uint32_t *starts, uint32_t start, int disable_pq, uint32_t n_centroids, uint32_t *out_ids,
                    int64_t *out_scores, uint32_t *out_len, uint32_t out_cap, uint64_t *cmps, uint64_t *pq_cmps);
/* The same traversal with candidates ranked by  f32(sum of LUT entries) * code_scale[id] + code_bias[q]: the RabitQ
 * estimator (diskann/rabitq.py:42-48) with luts / code_bias from mse_rabitq_preprocess_query, codes from mse_rabitq_encode
 * (mse_index_set_pq_codes) and code_scale[id] = norms[id] * dots[id] (mse_index_set_code_scales). */
int mse_search_beam_scaled(mse_index *ix, const uint16_t *q_f16, const float *luts, const float *code_bias, const float *desc_scales,
                           uint32_t nq, uint32_t L, uint32_t W, const uint32_t *starts, uint32_t start, uint32_t n_centroids,
                           uint32_t *out_ids, int64_t *out_scores, uint32_t *out_len, uint32_t out_cap, uint64_t *cmps, uint64_t *pq_cmps);
/* Runtime de-duplication + top-k over the visit lists the last mse_search_beam_dev call on this handle left in HBM
 * (query_disk_index.rs:99,486-529): walking the expanded nodes in visit order, a node is dropped when an earlier KEPT node has
 * dot product (cosine of the unit fp16 rows, f32) > threshold (the reference's DUPLICATES_THRESHOLD is 0.95); the kept nodes are
 * ranked by (score desc, visit order).  d_kept (optional): kept nodes per query.  Same stream as the search. */
int mse_dedup_topk_dev(mse_index *ix, uint32_t nq, float threshold, uint32_t topk, uint32_t *d_top_ids, int64_t *d_top_scores,
                       uint32_t *d_top_len, uint32_t *d_kept, void *stream);
int mse_index_set_code_scales(mse_index *ix, const float *scales);
/* Beam search with every buffer in HBM, asynchronous on `stream`; the best `topk` expanded nodes are selected on the device
 * ((score desc, visit order asc): the stable sort of query_disk_index.rs:303).  Candidate scores: d_luts ([nq][M*n_centroids],
 * PQ ADC) or, when d_qtm is given ([nq][output_dims+1] from mse_rabitq_query_dev), RabitQ byte tables built in shared memory.
 * mse_search_graph_check(ix, nq) afterwards reports visited-set / visit-list overflows. */
int mse_search_beam_dev(mse_index *ix, const uint16_t *d_q_f16, const float *d_luts, const float *d_qtm, uint32_t rabitq_output_dims,
                        uint32_t rabitq_n_dims, const float *d_desc_scales, uint32_t nq, uint32_t L, uint32_t W, const uint32_t *d_starts,
                        uint32_t start, uint32_t n_centroids, uint32_t topk, uint32_t *d_top_ids, int64_t *d_top_scores,
                        uint32_t *d_top_len, uint64_t *d_cmps, uint64_t *d_pq_cmps, void *stream);
/* evaluator brute force (query_disk_index.rs:262-273): exact i64 score of one fp16 query against every row */
int mse_scores_i64(mse_index *ix, const uint16_t *q_f16, int64_t *scores);

/* robust_prune (lib.rs:227-285) of point p over explicit candidates; out holds <= cfg->r ids */
int mse_robust_prune(mse_index *ix, uint32_t p, const uint32_t *cand_ids, const int64_t *cand_scores, uint32_t n_cand,
                     const mse_build_config *cfg, uint32_t *out, uint32_t *out_len);
/* IndexGraph::empty + random_fill_graph (lib.rs:22-31,376-387) */
int mse_index_random_fill_graph(mse_index *ix, uint32_t r, uint64_t seed);
/* medioid (lib.rs:54-68) */
int mse_index_medioid(mse_index *ix, uint32_t *out);
/* robust_stitch (lib.rs:326-374): nodes >= cfg->query_breakpoint are query nodes; base -> query edges are dropped and each
 * base node that had one receives the query's best out-neighbours (<= max_add_per_stitch_iter per query, up to r).
 * query_order: the query node ids in the order of the reference's shuffled loop (:333-334), or NULL for a seeded shuffle. */
int mse_index_robust_stitch(mse_index *ix, const mse_build_config *cfg, const uint32_t *query_order, uint64_t seed);
/* build_graph (lib.rs:287-324), batch-synchronous (see csrc/build.cu).  max_batch 0 = default.
 * stats (optional, 6 values): batches, point searches, back-edge merges, distance evaluations of the searches, searches whose
 * visited_list (lib.rs:205, unbounded there) was cut at max(8192, 48 l) entries, searches stopped by a visited-set overflow
 * (any of those fails the build with MSE_ERR_UNSUPPORTED) */
int mse_index_build_vamana(mse_index *ix, uint32_t medioid, const mse_build_config *cfg, uint64_t seed, uint32_t max_batch,
                           uint64_t *stats);

/* =====================================================================================
 * Codecs -- ProductQuantizer (diskann/src/vector.rs:308-406, opq.msgpack from diskann/aopq_train.py:87-93) and the
 * RabitQ experiment (diskann/rabitq.py:8-48, rabitq.msgpack :62-68)
 * ===================================================================================== */
typedef struct mse_pq mse_pq;
int mse_pq_create(const float *centroids, const float *transform, uint32_t n_dims, uint32_t n_dims_per_code, uint32_t n_centroids,
                  int device, mse_pq **out);
int mse_pq_load(const uint8_t *msgpack, size_t len, int device, mse_pq **out); /* rmp_serde::from_slice::<ProductQuantizer> */
int mse_pq_info(const mse_pq *pq, uint32_t out[4]);                            /* n_dims, n_dims_per_code, chunks, centroids */
int mse_pq_apply_transform(mse_pq *pq, const float *x, uint64_t n, float *y);  /* vector.rs:319-329 */
int mse_pq_encode(mse_pq *pq, const float *x, uint64_t n, uint8_t *codes);     /* quantize_batch :331-364 */
int mse_pq_preprocess_query(mse_pq *pq, const float *q, uint32_t nq, float *lut); /* :367-384, lut [nq][chunks*centroids] */
int mse_pq_adc(mse_pq *pq, const float *lut, const uint8_t *codes, uint64_t n, int64_t *scores); /* asymmetric_dot_product :387-405 */
void mse_pq_destroy(mse_pq *pq);

/* =====================================================================================
 * Shard centroids (kmeans.py) and shard assignment (src/dump_processor.rs:426-457) over the rows of a handle.
 * centroids: host f32 [k][d] row-major (centroids.bin decoded, dump_processor.rs:196-207).  k <= 256, spill <= 4.
 * ===================================================================================== */
/* kmeans.py:78-95 `fitness`: top-`spill` centroids of every row by inner product (normalize != 0: against the L2-normalised
 * centroids, :82), counts [spill][k] = the histogram of rank-j assignments (`cluster_sizes`), assign (optional, host) [n][spill].
 * Equal scores: the lower centroid index ranks first. */
int mse_kmeans_assign(mse_index *ix, const float *centroids, uint32_t k, uint32_t spill, int normalize, uint32_t *counts, uint32_t *assign);
/* kmeans.py:73-131 `simulated_annealing` (the script's loop; Gaussian steps from a seeded host generator, fitness on the device).
 * centroids_out [k][d] L2-normalised (:131); fitness_out = max |cluster size - n / k| of the accepted state; iters_out = iterations run. */
int mse_kmeans_anneal(mse_index *ix, uint32_t k, uint32_t spill, uint32_t max_iter, uint64_t seed, float *centroids_out, float *fitness_out,
                      uint32_t *iters_out);
/* dump_processor.rs:438-457: record i goes to the `spill` shards that come first when the shards are ordered (stable, in place) by
 * -scale_dot_result_f64(dot(centroid, row) - balance_fudge * shard_count / bal_count).  shard_counts [k] and bal_count carry the state
 * across calls (the reference starts them at 0 and 1); assign: host [n][spill] shard indices.  Dots are f32 on the device. */
int mse_shard_assign(mse_index *ix, const float *centroids, uint32_t k, uint32_t spill, double balance_fudge, uint64_t *shard_counts,
                     uint64_t *bal_count, uint32_t *assign);

typedef struct mse_rabitq mse_rabitq;
int mse_rabitq_create(const float *mean, const float *transform, uint32_t n_dims, uint32_t output_dims, int device, mse_rabitq **out);
int mse_rabitq_load(const uint8_t *msgpack, size_t len, int device, mse_rabitq **out);
/* the script's "training" (rabitq.py:11-28) over rows already in HBM: mean of the first sample_rows rows (0 = all; the script
 * takes 100 000) and the first output_dims rows of a random orthogonal matrix (seeded Gaussian rows orthonormalised on the device) */
int mse_rabitq_train(mse_index *ix, uint64_t sample_rows, uint32_t output_dims, uint64_t seed, mse_rabitq **out);
int mse_rabitq_info(const mse_rabitq *r, uint32_t out[2]); /* n_dims, output_dims */
/* mean [n_dims], transform [output_dims][n_dims] row-major -- what rabitq.msgpack stores (rabitq.py:62-68) */
int mse_rabitq_export(const mse_rabitq *r, float *mean, float *transform);
/* quantize (rabitq.py:30-36): codes [n][output_dims/8] sign bits, norms [n] = |o - mean|, dots [n] = <o_bar, P o_hat> */
int mse_rabitq_encode(mse_rabitq *r, const uint16_t *x_f16, uint64_t n, uint8_t *codes, float *norms, float *dots);
/* approx_dot (rabitq.py:42-48) of one f32 query against n encoded vectors */
int mse_rabitq_estimate(mse_rabitq *r, const float *q, const uint8_t *codes, const float *norms, const float *dots, uint64_t n,
                        float *estimates);
/* query side of approx_dot (rabitq.py:42-46) as byte tables for mse_search_beam_scaled: luts [nq][output_dims/8][256]
 * (entry v of table b = (1/sqrt(n_dims)) * sum_j (+-) (P q)[8b+j], sign from bit j of v), bias [nq] = <mean, q> */
int mse_rabitq_preprocess_query(mse_rabitq *r, const float *q, uint32_t nq, float *luts, float *bias);
/* encode the index's own rows in HBM: codes -> the index's code store, norms*dots (estimator 0, rabitq.py:48) or norms/dots
 * (estimator 1, the RabitQ paper) -> its code scales; equivalent to mse_rabitq_encode + mse_index_set_pq_codes + _set_code_scales */
int mse_index_encode_rabitq(mse_index *ix, mse_rabitq *r, int estimator);
/* the same query side kept in HBM: d_qtm [nq][output_dims + 1] = (P q, <mean, q>), asynchronous on `stream` */
int mse_rabitq_query_dev(mse_rabitq *r, const float *d_q, uint32_t nq, float *d_qtm, void *stream);
void mse_rabitq_destroy(mse_rabitq *r);

/* =====================================================================================
 * Embedding towers -- clip_server.py (OpenCLIP ViT-SO400M-14-SigLIP-384, precision fp16)
 *   model creation            clip_server.py:23      -> mse_encoder_create
 *   preprocess + encode_image clip_server.py:140,114 -> mse_encode_images_u8   (u8 RGB HWC in, x/127.5-1 on device)
 *   tokenizer + encode_text   clip_server.py:137,98  -> mse_encode_text_ids    (token ids in; SentencePiece stays host-side)
 *   features /= norm; fp16    clip_server.py:99,115,166 -> fused; outputs are unit-norm fp16 rows [batch][dim]
 *   batch > max_batch_size    clip_server.py:136,139 -> MSE_ERR_INVALID with the same message
 * Weights: an "MSEW0001" container holding the OpenCLIP/timm state_dict tensors under their own names
 * (visual.trunk.blocks.N.attn.qkv.weight, text.transformer.resblocks.N.attn.in_proj_weight, ... as
 * clip_server.py:46-62 lists them) plus an i32 "config" tensor; written by meme-search-engine_b200/weights.py.
 * One in-flight call per handle (the reference has a single inference thread, clip_server.py:126-128).
 * ===================================================================================== */
typedef struct mse_encoder mse_encoder;

int mse_encoder_create(const char *weights_path, int device, int max_batch, mse_encoder **out);
/* out[0..12] = image_size, patch, dim, vision depth, heads, mlp dim, vocab, context length, activation (1 erf-GELU,
 * 2 tanh-GELU), has_vision, has_text, text depth, padded patch K   (feeds GET /config: clip_server.py:176-183) */
int mse_encoder_config(const mse_encoder *e, int32_t out[16]);
int mse_encode_images_u8(mse_encoder *e, const uint8_t *rgb_hwc, int batch, uint16_t *out_f16);
int mse_encode_images_u8_dev(mse_encoder *e, const uint8_t *d_rgb_hwc, int batch, uint16_t *d_out_f16, void *stream);
/* The files the reference's clients actually send (src/common.rs:42-53 resize_for_embed_sync: image_size x image_size 24-bit BI_RGB BMPs
 * written by the image crate's BmpEncoder; clip_server.py:140 opens them with PIL): bmps[i] / lens[i] are host pointers to whole BMP
 * files, the headers are read on the host and the pixel arrays (BGR, bottom-up) are unpacked on the device.  Any other size or pixel
 * format is MSE_ERR_UNSUPPORTED -- decode it on the host and call mse_encode_images_u8. */
int mse_encode_images_bmp(mse_encoder *e, const uint8_t *const *bmps, const size_t *lens, int batch, uint16_t *out_f16);
/* resize_for_embed_sync (src/common.rs:31-54) on the device: an RGB8 image [h][w][3] of any size -> out_w x out_h by a separable
 * convolution in Pillow's fixed-point arithmetic (the fast_image_resize crate the reference calls descends from it); filter 0 = the
 * reference's rule (Hamming when both dimensions shrink, else Lanczos3, :43-44), 1 = Hamming, 2 = Lanczos3.  Host pointers. */
int mse_resize_rgb_u8(int device, const uint8_t *rgb, uint32_t w, uint32_t h, uint32_t out_w, uint32_t out_h, int filter, uint8_t *out);
/* decoded images of any size (rgb[i]: host RGB8 [heights[i]][widths[i]][3]) resized on the device into the tower's input: what the
 * ingest client does with resize_for_embed_sync + a BMP + the HTTP hop (src/main.rs:392,445,946) in one call */
int mse_encode_images_resized(mse_encoder *e, const uint8_t *const *rgb, const uint32_t *widths, const uint32_t *heights, int batch,
                              uint16_t *out_f16);
int mse_encode_text_ids(mse_encoder *e, const int32_t *ids, int batch, uint16_t *out_f16);
int mse_encode_text_ids_dev(mse_encoder *e, const int32_t *d_ids, int batch, uint16_t *d_out_f16, void *stream);
/* per-layer parity hooks: token activations [batch*S][dim] fp16 after the embedding and the first n_blocks blocks */
int mse_encode_images_hidden(mse_encoder *e, const uint8_t *rgb_hwc, int batch, int n_blocks, uint16_t *out_tokens_f16);
int mse_encode_text_hidden(mse_encoder *e, const int32_t *ids, int batch, int n_blocks, uint16_t *out_tokens_f16);
/* bracket every GEMM / attention launch with CUDA events on the launching stream (bench.py roofline) */
int mse_encoder_profile(mse_encoder *e, int enable);
/* last encode call (synchronises): out[0]=ns in GEMM kernels, out[1]=GEMM launches, out[2]=ns in attention kernels,
 * out[3]=attention launches, out[4]=kernel launches, out[5]=algorithmic GEMM MFLOP (sum of 2*M*N*K) */
int mse_encoder_stats(mse_encoder *e, uint64_t out[8]);
void mse_encoder_destroy(mse_encoder *e);
/* profiling aid, not part of the reference surface: average time (ms) of the attention kernel alone on random data */
int mse_debug_attention(int device, int batch, int seq, int mode, int iters, float *ms_out);
int mse_debug_gemm(int device, uint32_t M, uint32_t N, uint32_t K, int bn, int mode, int iters, float *ms_out);

/* =====================================================================================
 * Dense GEMM building block (tcgen05 + TMA): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]), fp16 in, fp32 accumulate
 * and output.  act: 0 none, 1 erf-GELU, 2 tanh-GELU.  This is the contraction behind every linear layer of the
 * towers and the OPQ rotation (diskann/src/vector.rs:320-329, matrixmultiply::sgemm); exposed with host
 * pointers so the tensor path can be validated on its own.  K % 8 == 0.
 * ===================================================================================== */
int mse_gemm_f16_tn(int device, const uint16_t *a_f16, const uint16_t *b_f16, uint32_t M, uint32_t N, uint32_t K,
                    const float *bias, int act, float *c_f32);

#ifdef __cplusplus
}
#endif
#endif /* MSE_B200_H */
