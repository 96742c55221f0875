/*
 * svdb_b200.h -- native C-ABI of the B200 similarity engine.
 *
 * One engine = one shard of the vector store resident in one GPU's HBM.
 * Plain pointers and sizes only; every function returns 0 on success or a
 * negative svdb_status, and svdb_last_error() describes the last failure of the
 * calling thread.  There is NO CPU fallback: without a usable CUDA device
 * svdb_engine_create fails.
 *
 * Which reference interface each entry point stands in for (paths relative to the
 * reference tree):
 *   svdb_insert_batch            vector_db_insert      src/vector_database.c:81-119
 *   svdb_update_batch            vector_db_update      src/vector_database.c:169-177
 *   svdb_delete_batch            vector_db_delete      src/vector_database.c:185-195
 *   svdb_append_kdpoints         kdtree_insert         src/kdtree.c:87-91
 *   svdb_nearest_batch (k=1)     kdtree_nearest        src/kdtree.c:171-178 (loop :131-162)
 *   svdb_compare_batch           cosine_similarity / euclidean_distance / dot_product
 *                                                      src/vector_database.c:301-352
 *   svdb_read_row                vector_db_read        src/vector_database.c:128-136
 * The reference's own function names (vector_db_*, kdtree_*, the three metrics)
 * are exported by the same library; they are declared in svdb_dropin.h.
 *
 * Semantics kept from the reference:
 *   - the searchable set is an APPEND-ONLY log of (kd-point, index) entries: one per
 *     insert and one per update; deletes remove nothing from it;
 *   - distance = squared L2 over the first kd_dim coordinates, summed sequentially
 *     in index order with a rounded multiply and a rounded add per term (no FMA);
 *   - nearest = smallest (distance, log sequence number); non-finite distances
 *     never win; empty log -> SVDB_NONE;
 *   - metrics accumulate in float exactly as the reference does; bit-identical.
 */
#ifndef SVDB_B200_H
#define SVDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVDB_NONE ((size_t)-1)
#define SVDB_MAX_K 24            /* largest k of a top-k query */

typedef enum svdb_status {
    SVDB_OK = 0,
    SVDB_ERR_ARG = -1,           /* bad argument (NULL, kd_dim > dimension, k > SVDB_MAX_K ...) */
    SVDB_ERR_CUDA = -2,          /* a CUDA call failed; text in svdb_last_error() */
    SVDB_ERR_OOM = -3,           /* device or host allocation failed */
    SVDB_ERR_RANGE = -4,         /* index out of range where the call cannot be a silent no-op */
    SVDB_ERR_STATE = -5          /* the shards of a collective call disagree (different inputs on different ranks) */
} svdb_status;

typedef enum svdb_metric {
    SVDB_COSINE = 0,             /* vector_database.c:301 */
    SVDB_EUCLIDEAN = 1,          /* vector_database.c:322 */
    SVDB_DOT = 2                 /* vector_database.c:342 */
} svdb_metric;

typedef struct svdb_engine svdb_engine;

typedef struct svdb_config {
    size_t dimension;            /* D: doubles per stored row (db_vector_size) */
    size_t kd_dim;               /* K <= D: coordinates nearest() measures (kd_tree_dimension) */
    int    device;               /* CUDA device ordinal */
    uint64_t seq_base;           /* global sequence number of this shard's first log entry */
    size_t reserve_rows;         /* map HBM for this many rows up front (0: grow on demand) */
    uint32_t flags;              /* SVDB_FLAG_* */
} svdb_config;

#define SVDB_FLAG_LOG_ONLY   1u  /* rows ARE kd-points (dimension == kd_dim), no index map: a bare KDTree */
#define SVDB_FLAG_NO_LOG     2u  /* rows only (compare / read); nearest is not available */
#define SVDB_FLAG_SHARD      4u  /* this engine holds a slice of a larger log: results come in plain (dist, seq)
                                    order and exact ties at the minimum are FLAGGED (SVDB_CAND_TIE); which tied
                                    entry the reference's tree reaches first is decided across the shards
                                    (svdb_resolve_ties_sharded; svdb_nearest_batch_sharded does it itself) */

/* One result of a nearest query; also the unit exchanged between shards. 32 bytes. */
typedef struct svdb_candidate {
    double   dist;               /* reference-order squared distance over kd_dim coords; +inf if none */
    uint64_t seq;                /* global log sequence number (tie-break key); UINT64_MAX if none */
    uint64_t index;              /* index carried by the log entry; SVDB_NONE if none */
    uint64_t flags;              /* SVDB_CAND_* */
} svdb_candidate;

#define SVDB_CAND_UNSAFE 1ull    /* candidate set could not be proven complete: rerun exact */
#define SVDB_CAND_TIE    2ull    /* sharded stores: two or more entries may sit at exactly the minimal distance;
                                    position 0 is the lowest seq until the tie is resolved across the shards */

const char *svdb_last_error(void);
const char *svdb_version(void);
int svdb_device_count(void);

int  svdb_engine_create(const svdb_config *cfg, svdb_engine **out);
void svdb_engine_destroy(svdb_engine *e);
/* Launch on this CUDA stream (a cudaStream_t; NULL is CUDA's legacy default stream).
 * SVDB_STREAM_OWN switches back to the engine's own non-blocking stream (the default). */
#define SVDB_STREAM_OWN ((void *)(intptr_t)-1)
int  svdb_set_stream(svdb_engine *e, void *stream);

/* ---- store deltas; rows are host buffers of n x ld doubles (ld >= dimension) ---- */
int svdb_insert_batch(svdb_engine *e, const double *rows, size_t n, size_t ld, size_t *first_index);
int svdb_update_batch(svdb_engine *e, const size_t *index, const double *rows, size_t n, size_t ld);
int svdb_delete_batch(svdb_engine *e, const size_t *index, size_t n);      /* applied in order */
/* Bare log append (n kd-points of ld >= kd_dim doubles, each with the index to report). */
int svdb_append_kdpoints(svdb_engine *e, const double *pts, const size_t *index, size_t n, size_t ld);
/* The same with the kd-points already in device memory (SVDB_FLAG_LOG_ONLY engines); entry i reports
 * index first_index + i. */
int svdb_append_kdpoints_device(svdb_engine *e, const double *d_pts, size_t first_index, size_t n, size_t ld);
/* Same as svdb_insert_batch with rows already in device memory (bulk ingest). */
int svdb_insert_batch_device(svdb_engine *e, const double *d_rows, size_t n, size_t ld, size_t *first_index);
/* Push pending deltas to the device (queries do this themselves). */
int svdb_flush(svdb_engine *e);

size_t svdb_size(const svdb_engine *e);        /* rows currently indexable (db->size) */
size_t svdb_log_size(const svdb_engine *e);    /* searchable log entries */
size_t svdb_dimension(const svdb_engine *e);
size_t svdb_kd_dim(const svdb_engine *e);
int svdb_read_row(svdb_engine *e, size_t index, double *out /* dimension doubles */);

/* ---- nearest ---- */
/* Host buffers.  Q: nq x ldq doubles (only the first kd_dim of each are read).
 * Outputs are nq x k, row-major, sorted by (dist, seq); missing entries get
 * SVDB_NONE / +inf / UINT64_MAX.  Any of the three outputs may be NULL. */
int svdb_nearest_batch(svdb_engine *e, const double *Q, size_t nq, size_t ldq, size_t k,
                       size_t *index_out, double *dist_out, uint64_t *seq_out);
/* Device buffers, asynchronous on the engine's stream: d_out receives nq x k candidates.
 * mode: SVDB_MODE_AUTO   tree traversal for thin kd-points (kd_dim <= 8; k = 1: the balanced median tree, with
 *                        the queries it flags as distinct-point ties re-answered by the reference's traversal),
 *                        else scan + exact re-rank;
 *       SVDB_MODE_EXACT  reference-order scan of every entry (no approximation to prove complete);
 *       SVDB_MODE_TREE   the reference's own traversal on the GPU tree (k > 1: its k-smallest generalisation,
 *                        position 0 is still the reference's answer).
 * The caller must look at flags (SVDB_CAND_UNSAFE) once the results are on the host and escalate
 * AUTO -> EXACT -> TREE for those queries (svdb_nearest_batch does exactly that). */
#define SVDB_MODE_AUTO  0
#define SVDB_MODE_EXACT 1
#define SVDB_MODE_TREE  2
#define SVDB_MODE_MTREE 3        /* k = 1, thin kd-points: balanced median tree + the reference's traversal for flagged ties */
#define SVDB_MODE_FP64  4        /* the scan of the fp64 rows + exact re-rank, whatever the batch size: no low-precision
                                    keys (bf16 planes, tensor cores).  First step of the escalation for answers that a
                                    low-precision path could not prove complete: AUTO -> FP64 -> EXACT -> TREE */
int svdb_nearest_batch_device(svdb_engine *e, const double *d_Q, size_t nq, size_t ldq, size_t k,
                              svdb_candidate *d_out, int mode);
/* Sharded store (one process per GPU): the same with the cross-shard exchange inside -- this shard's scan, the
 * peer-memory exchange and the merge; d_out receives the MERGED nq x k candidates.  A single-query call is ONE kernel
 * launch (the scan's last CTA re-ranks, stores to the peers, waits and merges).  Collective. */
struct svdb_exchange;
int svdb_nearest_batch_device_sharded(svdb_engine *e, struct svdb_exchange *x, const double *d_Q, size_t nq, size_t ldq,
                                      size_t k, svdb_candidate *d_out, int mode);
/* Cross-shard merge: d_in holds nshards blocks of nq x k candidates (an allgather result);
 * d_out receives nq x k, the k smallest of each query under (dist, seq). */
int svdb_merge_candidates_device(int device, void *stream, const svdb_candidate *d_in, size_t nshards,
                                 size_t nq, size_t k, svdb_candidate *d_out);

/* ---- cross-shard exchange over NVLink peer memory (one process per GPU) ----
 * Every rank creates an exchange, the 64-byte handles are all-gathered by the host (any transport),
 * every rank connects.  svdb_exchange_merge is then the whole "all-gather + merge" step: the rank
 * stores its candidates into every peer's buffer, publishes an epoch flag, waits for the peers'
 * flags and merges -- two small launches on `stream`, no NCCL call.  Collective: all ranks call it
 * in the same order.  max_records bounds nq * k of one call. */
typedef struct svdb_exchange svdb_exchange;
int  svdb_exchange_create(int device, int rank, int world, size_t max_records, svdb_exchange **out,
                          unsigned char handle_out[64]);
int  svdb_exchange_connect(svdb_exchange *x, const unsigned char *all_handles /* world x 64 bytes */);
void svdb_exchange_destroy(svdb_exchange *x);
int  svdb_exchange_merge(svdb_exchange *x, void *stream, const svdb_candidate *d_local, size_t nq, size_t k,
                         svdb_candidate *d_out);
/* The whole sharded query as one host call (host buffers in, MERGED answers out): this shard's scan,
 * the exchange and the merge, replayed as one CUDA graph from the second call of a shape on.
 * Collective: every rank calls it with the same queries, nq and k. */
int  svdb_nearest_batch_sharded(svdb_engine *e, svdb_exchange *x, const double *Q, size_t nq, size_t ldq, size_t k,
                                size_t *index_out, double *dist_out, uint64_t *seq_out);

/* ---- exact ties on a sharded store ----
 * The reference returns, among DISTINCT kd-points at exactly the same minimal distance, whichever its
 * insertion-order tree (kdtree.c:47-62) reaches first in the near-side-first traversal (:131-162).
 * A sharded store has no such tree in one place, so the shards walk the path of the GLOBAL tree together:
 * the node below the current one is the first log entry (lowest global seq) inside the current cell, found
 * by every shard in its slice and min-reduced; the walk goes to the query's side if a tied entry lives
 * there, else to the other side, and stops at the first tied entry it meets (or when one is left).  Two
 * small all-gathers per tree level, levels ~ depth of the tied entries' common ancestor; only queries whose
 * merged answer carries SVDB_CAND_TIE pay for it.
 *
 * svdb_tie_resolve is the walk itself, written against a backend (the shard's local primitives + an
 * all-gather); svdb_resolve_ties_sharded runs it with the CUDA backend of an engine.  Collective: every
 * rank calls it with the SAME queries and the SAME merged candidates (what the merge leaves on every rank).
 * merged (host, nq x k) is updated in place: winner first, the rest in (dist, seq) order. */
typedef struct svdb_tie_first {  /* a shard's first entry inside a cell */
    uint64_t seq;                /* global seq; UINT64_MAX if the shard has none in the cell */
    uint64_t index;
    double   v;                  /* its coordinate on the axis its level splits on */
    uint64_t tied;               /* 1 if it is itself one of the tied entries */
} svdb_tie_first;
typedef struct svdb_tie_split {  /* a shard's tied entries inside a cell, by side of a splitting value */
    uint64_t n[2];               /* side 0: coordinate < v, side 1: >= v (kdtree.c:52) */
    uint64_t min_seq[2];         /* lowest global seq on each side; UINT64_MAX if none */
    uint64_t min_index[2];
} svdb_tie_split;
typedef int (*svdb_allgather_fn)(void *ctx, const void *send, void *recv, size_t bytes_per_rank);
typedef struct svdb_tie_backend {
    void *ctx;
    int world, rank;
    size_t kd_dim;
    /* recv = world blocks of bytes_per_rank, in rank order (host memory) */
    svdb_allgather_fn allgather;
    void *allgather_ctx;
    /* remember, for each of ne events (query i: queries + i*kd_dim, distance dstar[i]), the LOCAL entries at
       exactly that reference distance.  n_local[i] = how many; same[i] = 1 iff they are all the same kd-point;
       first[i*kd_dim ..] = coordinates of the one with the lowest seq (untouched if none). */
    int (*collect)(void *ctx, size_t ne, const double *queries, const double *dstar, uint64_t *n_local,
                   uint8_t *same, double *first);
    /* For na active events (ev[i] = event id as passed to collect): the cell is the sequence of
       depth[i] (value, side) pairs path_v/path_side[i*path_ld + j], level j splitting axis j % kd_dim.
       out[i] = the first local entry with global seq > after[i] (UINT64_MAX: no lower bound) in the cell. */
    int (*first_in_cell)(void *ctx, size_t na, const uint32_t *ev, const uint32_t *depth, const double *path_v,
                         const uint8_t *path_side, size_t path_ld, const uint64_t *after, svdb_tie_first *out);
    /* out[i] = the local tied entries of event ev[i] inside the cell, split by v[i] on axis depth[i] % kd_dim */
    int (*split)(void *ctx, size_t na, const uint32_t *ev, const uint32_t *depth, const double *path_v,
                 const uint8_t *path_side, size_t path_ld, const double *v, svdb_tie_split *out);
} svdb_tie_backend;
int svdb_tie_resolve(const svdb_tie_backend *b, const double *Q, size_t nq, size_t ldq, svdb_candidate *merged,
                     size_t k, uint64_t *levels_walked /* may be NULL */);
/* The same with this engine's CUDA backend; allgather moves host buffers between the ranks (any transport).
 * x != NULL: use the peer-memory exchange instead (allgather may then be NULL). */
int svdb_resolve_ties_sharded(svdb_engine *e, svdb_exchange *x, int rank, int world, svdb_allgather_fn allgather,
                              void *allgather_ctx, const double *Q, size_t nq, size_t ldq, svdb_candidate *merged,
                              size_t k);

/* Concurrent single-query callers (the server is thread-per-connection, main.c:382, and
 * kdtree_nearest is called with no lock held, compare_handler.c:403): calls with nq == 1 that
 * arrive while a pass is running are coalesced into the next pass (no added wait). Same results. */

/* ---- bulk load / save in the reference's file format (vector_database.c:203-292) ----
 * u64 count, then per row { char uuid[37]; u64 dimension; f64 data[dimension] }, native endian.
 * load: every row must have the same dimension; rows go straight to HBM (the unaligned records
 * are unpacked on the device), uuids stay in a host table; indices = file order, as
 * vector_db_load inserts them (:277-279).  save: current rows in index order. */
int svdb_engine_load_file(const char *path, size_t kd_dim, int device, uint32_t flags, svdb_engine **out);
int svdb_save_file(svdb_engine *e, const char *path);
int svdb_get_uuid(svdb_engine *e, size_t index, char out[37]);
int svdb_set_uuid(svdb_engine *e, size_t index, const char *uuid);

/* ---- compare ---- */
/* out[i] = metric(row[index1[i]], row[index2[i]]); out-of-range pairs give -1.0f
 * (the reference's mismatch sentinel, vector_database.c:302-305). Host buffers. */
int svdb_compare_batch(svdb_engine *e, int metric, const size_t *index1, const size_t *index2,
                       size_t n, float *out);
/* All three metrics in one pass over the rows: out is n x 3 (cosine, euclidean, dot). */
int svdb_compare_batch_all(svdb_engine *e, const size_t *index1, const size_t *index2, size_t n, float *out);
/* Device index / output buffers, asynchronous on the engine's stream. metric 3 = all (n x 3). */
int svdb_compare_batch_device(svdb_engine *e, int metric, const uint64_t *d_index1, const uint64_t *d_index2,
                              size_t n, float *d_out);
/* Two host vectors that need not be stored anywhere (the by-value call of the reference). */
int svdb_compare_vectors(int device, int metric, const double *a, const double *b, size_t D, float *out);

/* ---- introspection used by bench.py / tests ---- */
typedef struct svdb_stats {
    uint64_t kernels_launched;   /* our kernels launched by this engine so far */
    uint64_t exact_reruns;       /* queries that needed the exact fallback scan */
    uint64_t tree_reruns;        /* queries that needed the tree traversal as the last resort */
    uint64_t tree_rounds;        /* level-synchronous insertion rounds run so far (K5) */
    uint64_t coalesced_calls;    /* single-query calls answered inside another caller's pass */
    uint64_t coalesced_passes;   /* passes that carried more than one caller's query */
    uint64_t hbm_bytes_mapped;   /* physical HBM currently mapped by the arenas */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t tie_events;         /* sharded queries whose exact tie went through svdb_resolve_ties_sharded */
    uint64_t tie_levels;         /* tree levels walked for them, summed */
    uint64_t mtree_builds;       /* (re)builds of the balanced median tree (K8) */
    uint64_t mtree_levels;       /* its internal levels after the last build (2^levels leaves of <= 32 points) */
    uint64_t mtree_rows;         /* log entries it covers; later ones are scanned as a tail */
    uint64_t fp64_reruns;        /* queries a low-precision path flagged and the fp64 scan (K1) re-answered */
    uint64_t scan_plane_last;    /* what the last scan pass read: 0 fp64 rows (K1 / exact), 1 hi + lo bf16 planes (K11), 2 hi plane (K12),
                                    3 one-byte plane (K13) */
    uint64_t tree_dropped;       /* 1: the reference-shaped tree was deeper than "tree.max_depth" (degenerate insertion
                                    order, e.g. sorted input) and has been dropped: distinct kd-points at exactly equal
                                    minimal distance now resolve to the lowest seq, not to the reference's traversal order */
} svdb_stats;
int svdb_get_stats(const svdb_engine *e, svdb_stats *out);
/* Tuning knobs (name/value).  K1's launch shape: "scan.variant" (0 TMA ring, 1 plain loads), "scan.warps", "scan.stages",
 * "scan.ctas_per_sm", "scan.tile_rows", "scan.assign" (tile order), "scan.nq_per_pass" (queries sharing an fp64 pass, <= 8);
 * "scan.force_exact" (1: every query through the exact-order scan K1'), "nearest.tree_max_k" (largest k the reference-shaped
 * tree K6 answers), "host.graphs" (0: the host entry points launch directly instead of replaying captured CUDA graphs);
 * "nearest.umma_min_queries" (batches of at least this many queries take the tcgen05 path K10; 0 = never; default: chosen per
 * call from the store's size), "nearest.mma_min_queries" (same for the FP64 DMMA path K2),
 * "nearest.umma_min_kd_dim",
 * "scan.plane" (which copy of the log calls of few queries may scan: 3 = a one-byte plane, K13, exact integer keys, 1 byte per
 * coordinate -- the default; 2 = the bf16 hi plane of the split-bf16 shadow, K12, 2 bytes; 1 = hi + lo planes, K11, 4 bytes;
 * 0 = the fp64 rows, K1, 8 bytes.  Same answers on every setting: the survivors are re-ranked from the fp64 rows in the
 * reference's operation order and whatever cannot be proven complete is re-answered from the fp64 rows),
 * "scan.plane8_max_k" / "scan.plane_max_k" (largest k the one-byte / bf16 plane serves: 16 / 24),
 * "scan.plane8_max_queries" (calls of up to this many queries are K13 passes), "scan.plane8_pair" (two queries share a pass),
 * "scan.overlap_steps" (1, default: on the device entry points the scan of a single-query call may start, by programmatic
 * dependent launch, while the tail of the call before it still runs; a query buffer that changed under such a scan is
 * detected and the answers come back flagged SVDB_CAND_UNSAFE),
 * "scan.shadow" (round-1 name: 1 = scan.plane 1, 0 = scan.plane 0),
 * "scan.fuse_tail" (1, default: the scan's last CTA runs the re-rank and the cross-shard exchange itself),
 * "nearest.mtree" (AUTO may use the median tree), "mtree.lanes" (32/16/8 lanes per query), "mtree.tail_max";
 * "log.index_base": added to the index every log entry written from now on reports (a shard whose local
 * row i is global row lo + i sets it to lo, so that merged answers carry global row numbers). */
int svdb_set_option(svdb_engine *e, const char *name, long value);
/* Scan the log for nq queries already on the device, timing the scan kernel alone with
 * CUDA events on the engine's stream: returns average milliseconds per launch. */
int svdb_time_scan(svdb_engine *e, const double *d_Q, size_t nq, size_t ldq, size_t k, int iters, float *ms_out);
/* With option "profile.scan_events" = 1 every scan launch is bracketed by CUDA events on the
 * engine's stream; this returns their summed duration and count since the last call. */
int svdb_take_scan_time(svdb_engine *e, float *total_ms, uint64_t *launches);
/* Diagnostics of the tcgen05 batch path (K10): after a batch call made with option "umma.debug_keys" = 1, copies
 * the approximate keys of log rows 0..127 against the first bn queries of that call ([128][bn] floats, bn = 64,
 * 128 or 256 by batch size) to keys_out -- tests use it to check the key error bound.  With count = 128 * 256 + 16 the
 * last 16 floats are kilocycles CTA 0 spent per role (producer / MMA issuer / epilogue; scripts/k10_role_cycles.py). */
int svdb_debug_filter_keys(svdb_engine *e, float *keys_out, size_t count);
/* Diagnostics of the fused scan tail: with option "scan.tail_debug" = 1 the last fused scan launch leaves %globaltimer
 * stamps (ns): [0] the last CTA took its ticket, [1] CTA lists merged, [2] re-rank done, [3] answers stored to the peers,
 * [4] every peer's answers have landed, [5] merged; [32 + b] CTA b finished its part of the scan ([9..14]: phases inside the re-rank).  count <= 32 + CTAs. */
int svdb_debug_tail_times(svdb_engine *e, unsigned long long *out, size_t count);
/* Diagnostics of K13 (the one-byte plane of the log): par_out = { lo, step, measured max |x - x^| }, bytes_out (may be NULL)
 * receives the first nbytes bytes of the plane. */
int svdb_debug_plane8(svdb_engine *e, double par_out[3], unsigned char *bytes_out, size_t nbytes);

#ifdef __cplusplus
}
#endif
#endif /* SVDB_B200_H */
