/* include/mcgpu.h -- C-ABI of libmcgpu.so: the B200-native (sm_100a) build/query
 * hot path of MetaCache-MPI.
 *
 * The reference (jmabuin/metacache-mpi) has no plugin / FFI layer: the path is
 * compiled into one executable.  This header therefore mirrors the internal C++
 * seams at which the reference's own drivers (mode_build.cpp, querying.h,
 * classification.cpp) call into the path; each entry point names the seam it
 * replaces as reference file:line (paths relative to the reference's src/).
 * INTEGRATION.md shows the call-site patch a maintainer would apply.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns MCG_OK (0) or a negative
 *     mcg_status; mcg_last_error() gives the message of the calling thread.
 *   - one mcg_ctx per GPU (device + stream + scratch arena); one mcg_db per
 *     database shard (= reference MPI rank, shard = target_id % nshards).
 *     Calls on one ctx / db are not re-entrant.
 *   - "host" entry points take caller-owned host buffers (pinned memory is
 *     detected and copied asynchronously) and return when the results are
 *     there; "_dev" entry points take device pointers and enqueue on the ctx
 *     stream.  Sketching, build-side kernels and the merge never wait for the
 *     device; the query "_dev" calls wait once, for one 4-byte count (how many
 *     reads the fused kernel left to the staged path), and -- only when that
 *     count is not zero -- for the sizes the staged path allocates from.
 *   - sequences: one byte buffer + offsets[n+1] (int64).  Sketches:
 *     uint32[nwin][sketchlen], ascending, padded with MCG_NO_FEATURE.
 *   - taxa are dense indices: index t < ntargets is the sequence-level taxon
 *     of target t; lineage[tax][21] holds indices by rank, -1 = none
 *     (taxonomy.h:61-84 numbering, 0 = sequence ... 20 = root, 21 = none).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with MCG_ERR_CUDA.
 */
#ifndef MCGPU_H
#define MCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCG_VERSION 100
#define MCG_NO_FEATURE 0xFFFFFFFFu
#define MCG_NUM_RANKS 21
#define MCG_RANK_SEQUENCE 0
#define MCG_RANK_NONE 21
#define MCG_MAX_K 16          /* kmer_type = uint32_t, config.h:44-48 */
#define MCG_MAX_SKETCH 32     /* device bottom-s network width */
#define MCG_MAX_CAND 32       /* top-list width: one list entry per lane of a warp.  Beyond 16 entries the reference's
                               * std::sort (candidates.h:270) leaves the order of equal-hit taxa unspecified; here it
                               * stays stable. */
#define MCG_MAX_LOCS 254      /* sketch_database.h:375-378 */

typedef enum {
    MCG_OK = 0,
    MCG_ERR_ARG = -1,         /* std::invalid_argument in the reference */
    MCG_ERR_CUDA = -2,        /* device / driver failure, or no device */
    MCG_ERR_NOMEM = -3,
    MCG_ERR_STATE = -4,       /* call order (e.g. query before finalize) */
    MCG_ERR_LIMIT = -5        /* target_limit_exceeded_error, sketch_database.h:138 */
} mcg_status;

typedef struct mcg_ctx mcg_ctx;
typedef struct mcg_db mcg_db;

/* sketcher + window parameters: hash_dna.h:60-95 (kmer_size, sketch_size),
 * sketch_database.h:310-345 (window size / stride) */
typedef struct {
    int32_t k;                /* -kmerlen    1..16 */
    int32_t sketchlen;        /* -sketchlen  1..32 */
    int64_t winlen;           /* -winlen     */
    int64_t winstride;        /* -winstride  */
} mcg_sketch_cfg;

/* query_options.h:123-135 classification_options + query window parameters
 * (query_options.cpp:40-75) */
typedef struct {
    int32_t sketchlen;        /* query sketch size (default: database's) */
    int64_t winlen;           /* query window size */
    int64_t winstride;
    int32_t maxcand;          /* -maxcand 1..MCG_MAX_CAND (the host driver maps "unlimited", i.e. < 1, to MCG_MAX_CAND) */
    int32_t lowest;           /* -lowest  (rank index) */
    int32_t highest;          /* -highest (rank index) */
    int32_t hitmin;           /* -hitmin  (already defaulted, mode_query.cpp:247) */
    float   hitdiff;          /* -hitdiff as a fraction (query_options.cpp:167-169) */
    int64_t insertsize;       /* -insertsize */
    int32_t reference_compat; /* 0: every shard's candidates take part in the cross-shard merge (restated semantics).
                               * 1: reproduce the reference's merge bit for bit, including its defect Q1: candidate taxon
                               *    ids travel as uint32 (querying.h:984,1001) and are widened unsigned on receipt
                               *    (:958), so sequence-level taxa (ids -(tgt)-1) of every rank but the receiver itself
                               *    are dropped -- along the reduction tree of querying.h:892-1066, P a power of two */
} mcg_query_opts;

typedef struct mcg_comm mcg_comm;
#define MCG_COMM_ID_BYTES 128

/* ---------------------------------------------------------------- context */
const char* mcg_last_error(void);
int  mcg_version(void);
int  mcg_device_count(void);
/* stream: a cudaStream_t (as void*) to enqueue on, or NULL to create one (pass
 * cudaStreamLegacy = (void*)1 for the legacy default stream) */
int  mcg_ctx_create(int device, void* stream, mcg_ctx** out);
int  mcg_ctx_destroy(mcg_ctx*);
int  mcg_ctx_sync(mcg_ctx*);
void* mcg_ctx_stream(mcg_ctx*);
/* scratch blocks the library released are parked for reuse (cudaFree of multi-GB blocks is slow and synchronises
 * the device), at most a quarter of the device memory (MCG_CACHE_CAP_MB); mcg_ctx_trim returns all of them to the driver */
int  mcg_ctx_trim(mcg_ctx*);
/* number of kernels this library launched on the ctx since creation */
int64_t mcg_ctx_launches(mcg_ctx*);
/* measurement hook: while enabled, the library brackets every launch of its three dominant kernels with CUDA
 * events on the ctx stream (no synchronisation: an event record costs the stream nothing).  mcg_ctx_kernel_times
 * waits for the stream and returns the summed device time [ms] and the launch count per kernel since the
 * last call: which = 0 sketch (K1), 1 fused query kernel (K3-K5), 2 radix-sort scatter (K2); phases of the
 * sharded query: 3 window bookkeeping + K1 of the slice, 4 exchange of sketches, 5 left-over reads (staged
 * path), 6 exchange of top lists, 7 merge + decision. */
int  mcg_ctx_profile(mcg_ctx*, int enable);
int  mcg_ctx_kernel_times(mcg_ctx*, int which, double* ms, int64_t* launches);
/* device memory helpers for callers that keep data resident (bench, tests) */
int  mcg_dev_alloc(mcg_ctx*, int64_t bytes, void** dptr);
int  mcg_dev_free(mcg_ctx*, void* dptr);
/* page-locked host memory for callers without a CUDA runtime of their own (the host driver's read and result
 * buffers): copies from / to it run asynchronously and at full PCIe rate */
int  mcg_host_alloc(int64_t bytes, void** hptr);
int  mcg_host_free(void* hptr);
int  mcg_memcpy_h2d(mcg_ctx*, void* dptr, const void* hptr, int64_t bytes);
int  mcg_memcpy_d2h(mcg_ctx*, void* hptr, const void* dptr, int64_t bytes);

/* ------------------------------------------------------------- sketching */
/* replaces dna_encoding.h:261 for_each_window (window count only) */
int64_t mcg_count_windows(const mcg_sketch_cfg*, const int64_t* seq_offsets, int64_t nseq);

/* replaces sketcher::operator() applied to every window of every sequence:
 * hash_dna.h:115-152 via dna_encoding.h:261,459,305,189,114 and hash_int.h:40.
 * out_features[nwin][sketchlen], out_counts[nwin] (may be NULL). */
int  mcg_sketch_windows(mcg_ctx*, const mcg_sketch_cfg*,
                        const char* seq_bytes, const int64_t* seq_offsets, int64_t nseq,
                        uint32_t* out_features, uint32_t* out_counts, int64_t* out_nwin);
/* same on resident data; d_seq must be 16-byte aligned and readable up to
 * total_bytes + 64 (the TMA span copies are 16-byte granular);
 * d_seq_offsets[nseq+1] on the device; nwin = mcg_count_windows(...) */
int  mcg_sketch_windows_dev(mcg_ctx*, const mcg_sketch_cfg*,
                            const char* d_seq, int64_t total_bytes,
                            const int64_t* d_seq_offsets, int64_t nseq, int64_t nwin,
                            uint32_t* d_out_features, uint32_t* d_out_counts);

/* -------------------------------------------------------------- database */
/* sketch_database ctor + max_locations_per_feature (sketch_database.h:248-268,
 * :356-368; maxlocs is narrowed to uint8 like mode_build.cpp:659) */
int  mcg_db_create(mcg_ctx*, const mcg_sketch_cfg*, int32_t maxlocs,
                   int32_t shard, int32_t nshards, mcg_db** out);
int  mcg_db_destroy(mcg_db*);
/* database::add_target_distributed (sketch_database.h:519-563) for n targets in
 * ascending target id; targets with id % nshards != shard are skipped (:540).
 * out_windows[n] receives source.windows (0 for foreign targets); may be NULL. */
int  mcg_db_add_targets(mcg_db*, const char* seq_bytes, const int64_t* seq_offsets,
                        const uint32_t* target_ids, int64_t n, uint32_t* out_windows);
/* same with the sequence bytes resident on the device (seq_offsets and target_ids stay host arrays); the
 * bytes must be 16-byte aligned and readable up to the last offset + 64 */
int  mcg_db_add_targets_dev(mcg_db*, const char* d_seq_bytes, const int64_t* seq_offsets,
                            const uint32_t* target_ids, int64_t n, uint32_t* out_windows);
/* ends the insert phase: builds the feature -> location-list multimap with the
 * reference's capping semantics (hash_multimap.h:653,706; sketch_database.h:
 * 1089-1092: first maxlocs locations in ascending (tgt,win)) */
int  mcg_db_finalize(mcg_db*);
/* feature_count / dead_feature_count / location_count (sketch_database.h:1007-1017) */
int  mcg_db_counts(mcg_db*, uint64_t* nkeys, uint64_t* nkeys_nonempty, uint64_t* nlocs);
/* shard contents in the order of hash_multimap::serialize's payload
 * (hash_multimap.h:975-1029) sorted by key: non-empty keys ascending, sizes,
 * and the concatenated lists */
int  mcg_db_export(mcg_db*, uint32_t* keys, uint32_t* sizes, uint32_t* tgts, uint32_t* wins);
/* hash_multimap::deserialize (hash_multimap.h:923-968): load a finalized shard */
int  mcg_db_import(mcg_db*, const uint32_t* keys, const uint32_t* sizes,
                   const uint32_t* tgts, const uint32_t* wins, uint64_t nkeys, uint64_t nlocs);
/* get_keys_num_items (sketch_database.h:1060-1075): (key,size) of non-empty buckets */
int  mcg_db_key_sizes(mcg_db*, uint32_t* keys, uint32_t* sizes);
/* remove_features_with_more_locations_than_distributed (sketch_database.h:398-417) */
int  mcg_db_remove_keys(mcg_db*, const uint32_t* keys, uint64_t n, uint64_t* removed);
/* the whole -remove-overpopulated-features reduction of mode_build.cpp:847-1074
 * for shards living in this process (single-process multi-GPU or tests) */
int  mcg_remove_overpopulated(mcg_db** shards, int32_t nshards, uint64_t* removed);

/* same as mcg_db_key_sizes with device output arrays (sized by mcg_db_counts: nkeys_nonempty) */
int  mcg_db_key_sizes_dev(mcg_db*, uint32_t* d_keys, uint32_t* d_sizes);

/* ---------------------------------------------------------- communicator */
/* One rank = one GPU = one mcg_ctx; ranks are processes (torchrun / mpirun) or threads of one process.
 * Replaces MPI_Init / MPI_COMM_WORLD of main.cpp:41 for this path: NCCL over NVLink / NVSwitch.
 * mcg_comm_get_id on one rank, the 128 bytes travel to the others by whatever the launcher offers
 * (torch.distributed store, MPI_Bcast, shared memory of the process), then every rank calls
 * mcg_comm_create at the same time (collective). */
int  mcg_comm_get_id(uint8_t* id /* [MCG_COMM_ID_BYTES] */);
int  mcg_comm_create(mcg_ctx*, const uint8_t* id, int32_t rank, int32_t nranks, mcg_comm** out);
int  mcg_comm_destroy(mcg_comm*);
int  mcg_comm_rank(const mcg_comm*);
int  mcg_comm_size(const mcg_comm*);

/* the -remove-overpopulated-features reduction across ranks (mode_build.cpp:847-1074: tree reduce of
 * (key, bucket size) pairs to rank 0, MPI_Bcast of the removal list, sketch_database.h:398-417 clear):
 * (key, size) pairs go to rank key % P (all-to-all), the owner sums them, keys whose summed size exceeds
 * maxlocs - 1 are all-gathered and cleared on every rank.  Collective over the communicator; the shard
 * must be finalized.  removed_here: buckets cleared in this shard; removal_list_len: keys removed overall. */
int  mcg_db_remove_overpopulated_dist(mcg_db*, mcg_comm*, uint64_t* removed_here, uint64_t* removal_list_len);

/* -------------------------------------------------------------- taxonomy */
/* flattened ranked_lineages_cache (taxonomy.h:702-806, :576-597) */
int  mcg_tax_upload(mcg_ctx*, const int32_t* lineage, const uint8_t* rank,
                    int64_t ntaxa, int64_t ntargets);

/* ----------------------------------------------------------------- query */
/* per-shard part of query_batched_parallel2 (querying.h:792-814): for each read
 * (pair): accumulate_matches (sketch_database.h:806), merge_sort (querying.h:90),
 * make_classification_candidates (classification.cpp:210, candidates.h:119,236).
 * reads2/offs2 may be NULL (single-end).  Outputs [n][maxcand]: taxon index
 * (-1 = unused slot), hits; out_ncand[n]. */
/* (host buffers: the batch is cut into sub-batches of 256 k reads and H2D copy, query and D2H copy of
 * consecutive sub-batches overlap; pinned buffers make the copies asynchronous) */
int  mcg_query_batch(mcg_db*, const mcg_query_opts*,
                     const char* reads1, const int64_t* offs1,
                     const char* reads2, const int64_t* offs2, int64_t n,
                     int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand);
/* same with device-resident reads and outputs; offsets on the device; d_reads1 / d_reads2 must be
 * readable up to bytes + 64 (the TMA span copies of the sketch kernel are 16-byte granular) */
int  mcg_query_batch_dev(mcg_db*, const mcg_query_opts*,
                         const char* d_reads1, const int64_t* d_offs1, int64_t bytes1,
                         const char* d_reads2, const int64_t* d_offs2, int64_t bytes2,
                         int64_t n, int64_t max_read_len /* 0 = unknown: tightens the scratch bounds only */,
                         int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand);
/* debugging / parity seam: the sorted match list of every read
 * (querying.h:798-811); out arrays sized by a first call with NULL arrays that
 * fills out_counts[n] */
int  mcg_query_matches(mcg_db*, const mcg_query_opts*,
                       const char* reads1, const int64_t* offs1,
                       const char* reads2, const int64_t* offs2, int64_t n,
                       int64_t* out_counts, uint32_t* out_tgts, uint32_t* out_wins);

/* query_batched_parallel2 (querying.h:721-1173) as ONE collective call over all shards: every rank passes
 * ITS slice of the batch (slices in rank order form the batch) and receives the merged top lists, the
 * decision and the statistics counters of that slice.  Per part of the call:
 *   H2D of the slice over the rank's own PCIe link -> K1 on the slice only -> all-gather of the sketches
 *   (64 B per window) over NVLink -> every rank probes ALL reads against its shard (rows 10-15) ->
 *   all-to-all of the per-read top lists (rank r receives the lists of its slice from every shard) ->
 *   merge in rank order + classify + statistics of the slice (rows 16-19) -> D2H of the slice's results.
 * Replaces the reference's replicated sketching of every read on every rank and its blocking tree reduce
 * to rank 0 (querying.h:867-1071).  comm may be NULL or of size 1 (single shard: no exchange).
 * out_tax/out_hits [n_local][maxcand], out_ncand/out_best/out_best_rank [n_local]; any of them may be
 * NULL; out_stats22 (may be NULL): the slice's counters are ADDED.  offs1[0] / offs2[0] need not be 0 (read i =
 * reads1[offs1[i] .. offs1[i+1])): a slice of a larger batch is passed by pointing into its offset arrays. */
int  mcg_query_batch_sharded(mcg_db*, mcg_comm*, const mcg_query_opts*,
                             const char* reads1, const int64_t* offs1,
                             const char* reads2, const int64_t* offs2, int64_t n_local,
                             int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand,
                             int32_t* out_best, int32_t* out_best_rank, uint64_t* out_stats22);
/* same with the slice and the outputs resident on the device (one part; enqueues on the ctx stream and
 * waits for the fused kernel's left-over count only); d_stats22: counters are ADDED on the device */
int  mcg_query_batch_sharded_dev(mcg_db*, mcg_comm*, const mcg_query_opts*,
                                 const char* d_reads1, const int64_t* d_offs1, int64_t bytes1,
                                 const char* d_reads2, const int64_t* d_offs2, int64_t bytes2,
                                 int64_t n_local, int64_t max_read_len,
                                 int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                                 int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22);

/* cross-shard merge + decision + statistics on rank 0: querying.h:905-971
 * (insert in rank order), classification.cpp:236-265 (classify),
 * classification_statistics.h:70-78 (assign).
 * cand_*: [nshards][n][maxcand] / [nshards][n] in rank order.
 * out_stats22 may be NULL; counters are ADDED to it. */
int  mcg_merge_classify(mcg_ctx*, const mcg_query_opts*,
                        const int32_t* cand_tax, const uint32_t* cand_hits,
                        const int32_t* cand_n, int32_t nshards, int64_t n,
                        int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand,
                        int32_t* out_best, int32_t* out_best_rank, uint64_t* out_stats22);
int  mcg_merge_classify_dev(mcg_ctx*, const mcg_query_opts*,
                            const int32_t* d_cand_tax, const uint32_t* d_cand_hits,
                            const int32_t* d_cand_n, int32_t nshards, int64_t n,
                            int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                            int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22);

#ifdef __cplusplus
}
#endif
#endif
