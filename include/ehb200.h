/* ehb200 — C ABI of the B200-native ANN backend for embeddinghub.
 *
 * This is the drop-in boundary: everything the reference's
 * hnswlib-backed ANNIndex does for the k-NN hot path, as plain C entry points a
 * cgo / ctypes / C++ caller can bind.  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns an ehb_status (0 = OK); no exceptions cross the ABI;
 *     ehb_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every buffer; vectors are row-major contiguous fp32.
 *   - "host" entry points take host pointers and do the H2D/D2H copies
 *     themselves (what a Go slice / std::vector caller binds);  "_dev" entry
 *     points take device pointers + a cudaStream_t (passed as void*) and never
 *     synchronise, for callers that keep queries/results resident in HBM.
 *   - results are nearest-first; rows with fewer than k hits are padded with
 *     EHB_NO_LABEL / +inf and the true count is written to out_counts.
 *   - searches are re-entrant: several host threads may search one index at the
 *     same time (each in-flight search has its own stream and scratch; concurrent
 *     small ehb_index_search calls are coalesced into one batched launch by a
 *     combining queue — the goroutine-per-request pattern of
 *     serving/serving.go:744-771); mutations (add / remove / build / import) are
 *     exclusive.  The reference serialises everything under one service mutex
 *     (embeddinghub/embeddingstore/server.cc:175).
 *   - there is no CPU fallback: every call fails with EHB_ERR_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef EHB200_H
#define EHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHB_NO_LABEL UINT64_MAX

typedef enum ehb_status {
  EHB_OK = 0,
  EHB_ERR_INVALID = 1,   /* bad argument                                      */
  EHB_ERR_CUDA = 2,      /* CUDA runtime / launch failure, or no device        */
  EHB_ERR_OOM = 3,       /* device or host allocation failed                   */
  EHB_ERR_STATE = 4,     /* call not valid in the index's current state        */
  EHB_ERR_NOT_FOUND = 5, /* unknown label                                      */
  EHB_ERR_IO = 6         /* save / load failed                                 */
} ehb_status;

/* Distance "space".  Replaces hnswlib::L2Space (embeddinghub/embeddingstore/
 * index.cc:12-13: squared L2, no sqrt), hnswlib::InnerProductSpace (1 - dot) and
 * hnswlib's cosine convention (normalise on insert and on query, then IP) that
 * the Go providers use (provider/redis.go:253, provider/pinecone.go:252). */
typedef enum ehb_metric { EHB_L2 = 0, EHB_IP = 1, EHB_COSINE = 2 } ehb_metric;

typedef enum ehb_precision { EHB_FP32 = 0, EHB_BF16 = 1 } ehb_precision;

typedef struct ehb_index ehb_index; /* opaque */

/* Construction parameters.  Zero-initialise, then set what you need;
 * ehb_params_default() fills the reference's implicit hnswlib defaults
 * (index.cc:14-15: M=16, ef_construction=200, random_seed=100; ef=10 because the
 * reference never calls setEf; init capacity 128, index.h:21). */
typedef struct ehb_params {
  uint32_t dim;
  int32_t metric;           /* ehb_metric                                      */
  uint64_t capacity;        /* initial capacity in vectors; grows by doubling  */
  uint32_t M;               /* 2..16 (level-0 rows hold 2*M ids)               */
  uint32_t ef_construction;
  uint32_t ef_search;       /* default ef of searches (hnswlib ef_)            */
  uint64_t seed;            /* level generator seed                            */
  int32_t device;           /* CUDA device ordinal                             */
  uint32_t build_batch;     /* max points linked per build wave (0 = default)  */
  uint32_t reserved[6];
} ehb_params;

typedef struct ehb_stats {
  /* counters of the most recent graph search (hnswlib metric_hops /
   * metric_distance_computations semantics: one hop per expanded node, one eval
   * per unvisited neighbour + 1 for the entry point) */
  uint64_t queries;
  uint64_t hops_upper;
  uint64_t hops_base;
  uint64_t dist_evals;
  uint64_t visited_overflow; /* queries whose visited table filled up          */
  uint64_t algorithmic_bytes; /* hops_upper*4M + hops_base*8M + evals*4d + Q*4d */
  /* index shape */
  uint64_t size, capacity, upper_rows;
  uint32_t dim, M, max_level, entry_point;
  uint64_t device_bytes;
  uint64_t deleted;           /* tombstones (ehb_index_remove); `size` counts them, like hnswlib   */
  uint64_t combined_batches;  /* batched launches issued by the combining queue of ehb_index_search */
  uint64_t combined_queries;  /* queries those launches carried                                     */
  uint32_t metric;            /* ehb_metric of the index                                            */
  uint32_t reserved_;
} ehb_stats;

const char* ehb_last_error(void);
uint32_t ehb_abi_version(void);

void ehb_params_default(ehb_params* p, uint32_t dim);
/* Number of usable CUDA devices (what a caller sizes device_ids[] from); EHB_ERR_CUDA when there is none. */
int ehb_device_count(int32_t* out);

/* ANNIndex::ANNIndex(dims, init_cap) — index.cc:10-18 (allocates the hnswlib
 * arena); here: device arrays for vectors, labels, levels and adjacency. */
int ehb_index_create(const ehb_params* p, ehb_index** out);
int ehb_index_destroy(ehb_index* ix);

/* ANNIndex::set -> hnswlib addPoint / resizeIndex — index.cc:20-37.  Insert or
 * update-in-place (existing label).  labels == NULL assigns labels
 * size()..size()+n-1.  Points are linked into the graph lazily by the next
 * ehb_index_build / search (batched GPU construction replaces the reference's
 * one-addPoint-per-row loop, version.cc:64-74). */
int ehb_index_add(ehb_index* ix, uint64_t n, const float* vecs_host, const uint64_t* labels_host);
int ehb_index_add_dev(ehb_index* ix, uint64_t n, const float* vecs_dev, const uint64_t* labels_host);
int ehb_index_build(ehb_index* ix);

/* Delete (embeddinghub/docs/reading_and_writing_embeddings.md:49-66 promises delete / multidelete; hnswlib
 * markDelete semantics): the points stay in the graph as tombstones — traversed by searches, never returned,
 * ehb_index_get answers EHB_ERR_NOT_FOUND, size() still counts them.  Unknown label: EHB_ERR_NOT_FOUND;
 * already deleted: EHB_ERR_STATE.  Adding a deleted label again un-deletes it and updates it in place
 * (hnswlib addPoint). */
int ehb_index_remove(ehb_index* ix, uint64_t n, const uint64_t* labels_host);

/* hnswlib setEf (never called by the reference; named by BASELINE configs). */
int ehb_index_set_ef(ehb_index* ix, uint32_t ef);
int ehb_index_size(ehb_index* ix, uint64_t* out);

/* Version::get of the stored vector (version.cc / storage.cc:32-36 serve this
 * from RocksDB; the index keeps the fp32 rows resident so Get needs no KV). */
int ehb_index_get(ehb_index* ix, uint64_t label, float* out_vec_host);

/* ANNIndex::approx_nearest -> hnswlib searchKnn — index.cc:39-52, batched over
 * nq queries.  ef == 0 uses the index default; the walk runs with max(ef, k) like
 * searchKnn.  out_dists / out_counts may be NULL. */
int ehb_index_search(ehb_index* ix, uint64_t nq, const float* queries_host, uint32_t k, uint32_t ef,
                     uint64_t* out_labels_host, float* out_dists_host, uint32_t* out_counts_host);
int ehb_index_search_dev(ehb_index* ix, uint64_t nq, const float* queries_dev, uint32_t k, uint32_t ef,
                         uint64_t* out_labels_dev, float* out_dists_dev, uint32_t* out_counts_dev, void* stream);

/* hnswlib BruteforceSearch (not used by the reference; the exact path named by
 * the north star).  EHB_FP32 is exact with a defined total order (distance asc,
 * insertion index asc) and canonical arithmetic (one fp32 FMA chain, k
 * ascending) so ids are reproducible bit-for-bit.  EHB_BF16 is the tensor-core
 * path (bf16 GEMM + fp32 re-rank of an oversampled candidate set). */
int ehb_index_search_bruteforce(ehb_index* ix, uint64_t nq, const float* queries_host, uint32_t k, int precision,
                                uint64_t* out_labels_host, float* out_dists_host, uint32_t* out_counts_host);
int ehb_index_search_bruteforce_dev(ehb_index* ix, uint64_t nq, const float* queries_dev, uint32_t k, int precision,
                                    uint64_t* out_labels_dev, float* out_dists_dev, uint32_t* out_counts_dev,
                                    void* stream);

int ehb_index_stats(ehb_index* ix, ehb_stats* out);

/* Timing of the most recent search on the index's launch stream, measured with
 * CUDA events recorded around the kernel(s): milliseconds of the graph-walk (or
 * brute-force) kernels alone.  Synchronises on those events. */
int ehb_index_last_kernel_ms(ehb_index* ix, float* out_ms);
/* Name (with template arguments) of the graph-walk kernel that search launched. */
int ehb_index_last_kernel_name(ehb_index* ix, char* out, uint32_t out_bytes);

/* Graph exchange (hnswlib saveIndex/loadIndex are never called by the
 * reference; persistence is a "next" row).  Layout:
 *   levels[n] u8; links0[n][2M] u32 padded with UINT32_MAX; up_off[n] u32 = first
 *   upper row of node i (UINT32_MAX if level 0); links_up[rows][M] u32.
 * import replaces the whole index content (vectors are taken as stored, i.e.
 * already normalised for cosine). */
int ehb_index_export_graph(ehb_index* ix, float* vectors, uint64_t* labels, uint8_t* levels, uint32_t* links0,
                           uint32_t* up_off, uint32_t* links_up, uint32_t* entry, int32_t* max_level);
int ehb_index_import_graph(ehb_index* ix, uint64_t n, const float* vectors, const uint64_t* labels,
                           const uint8_t* levels, const uint32_t* links0, const uint32_t* up_off,
                           uint64_t upper_rows, const uint32_t* links_up, uint32_t entry, int32_t max_level);
int ehb_index_save(ehb_index* ix, const char* path);
int ehb_index_load(const char* path, int32_t device, ehb_index** out);

/* Final merge of G per-shard top-k lists (each [nq][k], nearest-first, padded
 * with EHB_NO_LABEL/+inf) gathered contiguously as [G][nq][k] — the step after
 * the single NCCL all-gather of the range-sharded index. */
int ehb_merge_topk_dev(uint32_t G, uint64_t nq, uint32_t k, const float* dists_dev, const uint64_t* labels_dev,
                       float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev, int32_t device,
                       void* stream);

/* Same merge over ONE packed gather buffer: rank g's block starts at packed + g * rank_stride_bytes and
 * holds [nq*k u64 labels | nq*k f32 distances] — what a single all-gather of each rank's packed
 * (labels, distances) result delivers, merged in place without unpacking. */
int ehb_merge_topk_packed_dev(uint32_t G, uint64_t nq, uint32_t k, const void* packed_dev, uint64_t rank_stride_bytes,
                              float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev, int32_t device,
                              void* stream);

/* Warps cooperating on one query: 1 = the exact hnswlib expansion order (one warp
 * per query); 2 or 4 = that many of the closest unexpanded candidates are expanded
 * concurrently per round (recall >= the sequential walk's at the same ef; used when
 * a batch is too small to fill the GPU with one warp per query); 0 = automatic. */
int ehb_index_set_search_width(ehb_index* ix, uint32_t warps_per_query);

/* Search tuning knobs (advanced; 0 = automatic).  stage_slots: vectors staged
 * per TMA group; stage_groups: groups in flight per warp; hash_bits: log2 of the
 * per-warp visited table. */
int ehb_index_set_tuning(ehb_index* ix, uint32_t stage_slots, uint32_t stage_groups, uint32_t hash_bits,
                         uint32_t warps_per_block);

/* ---------------------------------------------------------------------------------------------------
 * Range-sharded index over several GPUs of one box (SURVEY.md §8b B4 "device_ids[], n_dev"; §8e).
 * One process drives n_dev devices: labels [i*span, (i+1)*span) live on shard i % n_dev (span 0: label %
 * n_dev); every shard owns an independent graph; a search runs on every shard, whose kernels store their
 * top-k straight into device_ids[0]'s gather buffer over NVLink (peer access), and one merge kernel there
 * produces the result.  No collective on either path.  Same conventions as the single-index calls; calls on one
 * ehb_sharded handle are serialised (one gather buffer per handle). */
typedef struct ehb_sharded ehb_sharded; /* opaque */
int ehb_sharded_create(const ehb_params* p /* device ignored; capacity per shard */, const int32_t* device_ids,
                       uint32_t n_dev, uint64_t shard_span, ehb_sharded** out);
int ehb_sharded_destroy(ehb_sharded* sh);
int ehb_sharded_n_shards(ehb_sharded* sh, uint32_t* out);
int ehb_sharded_shard(ehb_sharded* sh, uint32_t i, ehb_index** out /* borrowed */);
int ehb_sharded_add(ehb_sharded* sh, uint64_t n, const float* vecs_host, const uint64_t* labels_host);
int ehb_sharded_remove(ehb_sharded* sh, uint64_t n, const uint64_t* labels_host);
int ehb_sharded_get(ehb_sharded* sh, uint64_t label, float* out_vec_host);
int ehb_sharded_size(ehb_sharded* sh, uint64_t* out);
int ehb_sharded_build(ehb_sharded* sh); /* shards build concurrently */
int ehb_sharded_set_ef(ehb_sharded* sh, uint32_t ef);
int ehb_sharded_search(ehb_sharded* sh, uint64_t nq, const float* queries_host, uint32_t k, uint32_t ef,
                       uint64_t* out_labels_host, float* out_dists_host, uint32_t* out_counts_host);
int ehb_sharded_search_bruteforce(ehb_sharded* sh, uint64_t nq, const float* queries_host, uint32_t k, int precision,
                                  uint64_t* out_labels_host, float* out_dists_host, uint32_t* out_counts_host);

/* Shard exchange for one-process-per-GPU deployments (torchrun / MPI): replaces "one ncclAllGather of the
 * per-shard top-k + merge kernel"  with ONE kernel per rank that pushes this rank's lists
 * into every peer's receive buffer with stores over NVLink (CUDA IPC mappings), flags them per slice and
 * merges each slice as soon as all peers' flags are up.  Protocol per rank:
 *   create -> ipc_handle -> (exchange the 64-byte handles out of band, e.g. one torch.distributed
 *   all_gather at start-up) -> open;  then per step, in lock step on every rank:
 *   begin(nq, k) -> run the local search with the returned output pointers -> merge_dev(stream). */
#define EHB_IPC_HANDLE_BYTES 64
typedef struct ehb_exchange ehb_exchange; /* opaque */
int ehb_exchange_create(int32_t device, uint32_t world, uint32_t rank, uint64_t max_nq, uint32_t max_k,
                        ehb_exchange** out);
int ehb_exchange_destroy(ehb_exchange* ex);
int ehb_exchange_ipc_handle(ehb_exchange* ex, void* out_handle /* EHB_IPC_HANDLE_BYTES */);
int ehb_exchange_open(ehb_exchange* ex, const void* handles /* [world][EHB_IPC_HANDLE_BYTES], rank order */);
int ehb_exchange_attach_local(ehb_exchange* ex, uint32_t peer_rank, ehb_exchange* peer /* same process */);
int ehb_exchange_begin(ehb_exchange* ex, uint64_t nq, uint32_t k, uint64_t** labels_dev, float** dists_dev);
int ehb_exchange_merge_dev(ehb_exchange* ex, float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev,
                           void* stream);
/* The fused step for graph searches (replaces begin + ehb_index_search_dev + merge_dev): the walk kernel's
 * epilogue stores each query's top-k into every peer's receive buffer (coalesced stores over NVLink, overlapping
 * the rest of the walk) and raises per-slice flags; one kernel then waits for the peers' flags and merges.
 * shard_counts_dev ([nq], this shard's hit counts) may be NULL. */
int ehb_exchange_search_dev(ehb_exchange* ex, ehb_index* ix, uint64_t nq, const float* queries_dev, uint32_t k,
                            uint32_t ef, float* out_dists_dev, uint64_t* out_labels_dev, uint32_t* out_counts_dev,
                            uint32_t* shard_counts_dev, void* stream);
int ehb_exchange_timed_out(ehb_exchange* ex, uint32_t* out /* 1: a wait for a peer gave up (~20 s) */);

/* Named integer options (A/B switches and construction knobs that are not part of
 * the reference's surface).  Unknown names fail with EHB_ERR_INVALID.
 *   "build_frac"    a construction wave links at most size/build_frac points (0 = default 64)
 *   "bf16_unfused"  bf16 brute force keeps the distance tiles in HBM (A/B of the fused epilogue)
 *   "gemm_2cta"     bf16 brute force uses the cta_group::2 cluster form of the GEMM
 *   "combine"       1 (default): concurrent host searches of <= 256 queries share batched launches
 *   "walk_prefetch" 1: L2-prefetch the speculated next hop's vectors (default 0: it cost 27 % extra DRAM traffic) */
int ehb_index_set_option(ehb_index* ix, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* EHB200_H */
