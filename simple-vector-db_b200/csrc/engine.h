// engine.h -- host side of one shard: HBM arenas, delta staging, query orchestration.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/svdb_b200.h"
#include "arena.h"
#include "kernels.h"

namespace svdb {

// bumped whenever any Scratch / PinnedScratch block is reallocated: captured CUDA graphs that baked the
// old pointers in must not be replayed
unsigned long long scratch_generation();

void set_last_error(const std::string &s);
const std::string &get_last_error();

// cudaMalloc'ed scratch that only ever grows.
struct Scratch {
    void *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes, std::string &err);
    void free_();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinnedScratch {
    void *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes, std::string &err);
    void free_();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace svdb

struct svdb_exchange;
namespace svdb {
struct TieGpu;                       // device scratch of the cross-shard tie walk (tie_protocol.cu)
void tie_state_free(TieGpu *t);
cudaError_t exchange_enqueue(svdb_exchange *x, cudaStream_t st, const svdb_candidate *d_local, size_t nq, size_t k,
                             svdb_candidate *out);
bool exchange_fits(const svdb_exchange *x, size_t nq, size_t k);
void exchange_fill_tail(const svdb_exchange *x, TailArgs &t, svdb_candidate *xout);
int exchange_rank(const svdb_exchange *x);
int exchange_world(const svdb_exchange *x);
int exchange_allgather_host(svdb_exchange *x, cudaStream_t st, const void *send, void *recv, size_t bytes);
// tie_protocol.cu: position 0 of TIE-flagged merged answers = the entry the reference's global tree reaches
// first (collective over the shards; callers hold e->mu)
int resolve_ties_engine(svdb_engine *e, svdb_exchange *x, int rank, int world, svdb_allgather_fn ag, void *ag_ctx,
                        const double *Q, size_t nq, size_t ldq, svdb_candidate *merged, size_t k);
}  // namespace svdb

// Layout in HBM (all fp64, row-major, one entry per VERSION = per insert/update/log append):
//   rows     [versions][Dpad]   Dpad = D rounded up to 16 doubles (128-byte rows), zero padded
//   kdpts    [versions][kstride] first K coordinates; aliases rows when K == D on the wide path
//   log_idx  [versions] u64     index the entry was appended with (what nearest returns)
//   norms    [versions] f32     float-order self dot product (cosine), computed at insert
//   cur      [size] u64         index -> version of its current row (shifted by deletes)
//   xnorm    [versions] f64     |kd-point|^2 (any order) for the GEMM-form keys of K2; wide engines only
//   child    [versions][2] u32  reference-shaped KD tree links (less / greater-or-equal child), K5
struct svdb_engine {
    svdb_config cfg;
    int D = 0, K = 0, Dpad = 0, kstride = 0;
    bool log_only = false, no_log = false, alias = false, wide = false, use_tree = false;
    int mma_min_q = 4;                   // AUTO: batches of at least this many queries take the DMMA path (K2)
    // K10: batches of at least umma_min_q queries over kd-points of at least umma_min_k coordinates take the tcgen05 path
    // (split-bf16 keys + the same exact re-rank); 0 switches it off.  Needs a bf16 shadow of the log (4 bytes per
    // coordinate, built on first use and extended incrementally); if that does not fit, K2 keeps serving.
    // Thresholds from profiles/r02_sweep_batch_paths.jsonl: with 64-query CTA groups K10 answers 3..64 queries in the time of
    // ~0.65 fp64 HBM passes at kd_dim 768 (K2: 1.4 - 4 passes), so it takes over from 3 queries on for kd_dim >= 256; for
    // shorter rows its per-tile epilogue dominates and K2 stays ahead up to 16 queries (set in init(); -1 = not chosen yet).
    int umma_min_q = -1, umma_min_k = 32;
    // largest k the single-plane scans serve (a coarser key means a wider re-rank window: more rows for the tail to fetch)
    int plane_max_k = 24, plane8_max_k = 16;
    // K13 streams an eighth of what K2 reads and a quarter of what K10 reads per pass: repeated passes beat both up to this
    // many queries per call (cost model in DESIGN.md section 4, from profiles/r02_sweep_batch_paths.jsonl and r02_window_counts_*)
    int plane8_max_q = 4;
    // Back-to-back single-query steps (device entry points): the K13 scan of step i+1 is launched with programmatic stream
    // serialization and starts while the tail of step i (re-rank, proof, exchange) still runs on one SM; it uses one CTA
    // less than the machine holds so that all of them start at once.  pdl_mark = stats.kernels_launched right after such a
    // launch: the attribute is only set while nothing else of this engine ran in between (plane builds, inserts, ...).
    int overlap_steps = 1;
    bool in_host_call = false;           // set for the duration of nearest_host (host buffers in and out): plain launches
    bool plane8_pair = true;             // K13: two queries of a call share a pass where the kernel supports it (Kp >= 192); 0: one each (A/B)
    bool umma_sparse_checks = true;      // K10: prune check / threshold refresh every fourth tile once thresholds are tight (0: every tile, A/B)
    bool umma_group_min = true;          // K10: thresholds from the group's published minima (umma_filter.cu); 0: from gtau alone (A/B)
    uint64_t pdl_mark = ~0ull;
    bool umma_min_user = false, mma_min_user = false, plane8_max_q_user = false;      // thresholds set through svdb_set_option: taken literally
    bool byte_plane_serves(size_t k) const;                // K13 usable for a call asking for k neighbours per query
    int byte_plane_max_queries(size_t k) const;            // ... for calls of up to this many queries (0: not at all)
    void batch_thresholds(size_t k, int &uq, int &mq) const;   // from how many queries K10 / K2 take a call of this k
    bool umma_ok = true, shadow_ready = false;
    size_t shadow_n = 0;                 // log entries present in the hi plane ...
    size_t shadow_lo_n = 0;              // ... and in the lo plane (built when K10 / K11 first ask for it: K12 reads hi only)
    size_t shadow_mapped_counted = 0;    // part of the shadow's mapped bytes already included in stats.hbm_bytes_mapped
    svdb::DeviceBuffer shadow_hi, shadow_lo;   // [versions][Kp] bf16 each: x ~ hi + lo (split_bf16_kernel)
    // K13: the one-byte plane ([versions][Kp] bytes on one store-wide grid), its grid + measured error (plane8_par: a
    // svdb::Plane8Par followed by the error word)
    svdb::DeviceBuffer plane8;
    svdb::Scratch plane8_par;
    bool plane8_ready = false, plane8_ok = true;
    size_t plane8_n = 0, plane8_mapped_counted = 0;
    // queries K13 answered / could not prove, by result size (k <= 4, k > 4: larger k means a wider re-rank window): the engine
    // stops using the plane for a class whose proofs mostly fail (plane8_ok covers k <= 4 and every k, plane8_ok_bigk only k > 4)
    uint64_t p8_calls[2] = {0, 0}, p8_unsafe[2] = {0, 0};
    bool plane8_ok_bigk = true;
                                               // for data its grid resolves badly (nearest_host)
    int ensure_plane8();                 // SVDB_OK, an error, or -1000: not available
    svdb::Scratch qsplit, ubuf, udbg, plane_err, ticket, xlocal;
    // Which copy of the log 1-2 query calls scan (option "scan.plane"): 0 the fp64 rows (K1), 1 the hi + lo shadow (K11, 4 bytes
    // per coordinate), 2 the hi plane alone (K12, 2 bytes per coordinate), 3 the one-byte plane (K13; kd_dim 193..1024,
    // else 2 serves).  Same answers on every setting:
    // whatever the re-rank cannot prove complete is re-answered from the fp64 rows.
    int scan_plane = 3;
    bool fuse_tail = true;               // the scan's last CTA runs finalize (and the cross-shard exchange) itself
    int dyn_tiles = 0;                   // option "scan.dynamic_tiles": eighths of the log K12 hands out dynamically (A/B; slower)
    bool tail_debug = false;             // option "scan.tail_debug": fused tails leave %globaltimer stamps in tail_dbg
    svdb::Scratch tail_dbg;
    int last_scan_plane = 0;             // what the last scan pass of nearest_device read (escalation: skip a redundant K1 rerun)
    int ensure_shadow(bool need_lo);     // SVDB_OK, an error, or -1000: not available
    bool umma_debug = false;             // next K10 launch dumps the keys of its first tile into udbg
    bool umma_resident = true;           // K10 keeps the query planes in shared memory when they fit (option umma.resident_queries)
    int nearest_umma(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out);   // SVDB_OK, an error, or -1000: not available
    int tree_max_depth = 8192;           // deeper than this (degenerate insertion order): the tree is dropped
    int tree_max_k = 8;                  // K <= this and k == 1: answer by tree traversal (K6)
    // K8/K9: balanced median tree over thin kd-points (median_tree.cu); serves k = 1, flags distinct-point ties for K6
    bool use_mtree = false;              // engine may keep one (thin log, kd_dim <= 8)
    int mtree_auto = 1;                  // AUTO prefers it over K6 for k = 1
    int mtree_lanes = 0;                 // lanes per query in K9 (32, 16, 8); 0: chosen from the number of queries in the call
    int mtree_block = 1;                 // 1: split values in heap order; 3: 64-byte blocks of three levels (measured slower, kept as an option)
    size_t mtree_tail_min = 256, mtree_tail_max = 4096;   // rebuild once the unindexed tail exceeds clamp(n_built/8, min, max)
    uint64_t index_base = 0;             // added to the index a log entry reports (a shard's rows are global rows lo..)
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::mutex mu;

    svdb::DeviceBuffer rows, kdpts, log_idx, norms, cur, child, xnorm;
    std::vector<uint64_t> cur_host;      // authoritative index map
    size_t cur_uploaded = 0;             // entries of cur_host valid on the device ...
    size_t cur_dirty_lo = 0;             // ... below this index
    size_t n_versions = 0;               // entries resident on the device

    // pinned staging of not-yet-uploaded versions
    svdb::PinnedScratch stage_rows, stage_idx;
    size_t stage_ld = 0, stage_n = 0, stage_cap = 0;

    // query scratch
    svdb::Scratch qpad, qraw, lists, outc, idx1, idx2, fout, tree_pn, tree_pds, tree_flag, qnorm, xnmax;
    svdb::Scratch mt_split, mt_pts, mt_seq, mt_marks;
    svdb::MtreeView mt;
    svdb::PinnedScratch hq, hout, hidx, hf, tree_hflag;

    // host-side uuid of each index (file round trip); shifts with deletes like the reference's structs
    std::vector<std::array<char, 37>> uuids;

    // coalescing of concurrent single-query callers
    struct PendingQuery {
        const double *q;
        size_t k;
        svdb_candidate *res;     // k results
        int rc = 0;
        bool done = false;
    };
    std::mutex bq_mu;
    std::condition_variable bq_cv;
    std::vector<PendingQuery *> bq;
    bool bq_leader = false;
    int nearest_one_coalesced(const double *q, size_t k, svdb_candidate *res);

    // CUDA graphs of the host single/small-batch query path (H2D, kernels, D2H): one launch instead of
    // five driver calls.  Valid for one (nq, k) shape while the log, the scratch blocks, the options
    // and the stream stay the same.
    struct HostGraph {
        size_t nq, k, n_versions;
        svdb_exchange *x;
        unsigned long long gen;
        cudaStream_t stream;
        cudaGraphExec_t exec;
        uint64_t launches;
    };
    std::vector<HostGraph> graphs;
    bool graphs_enabled = true;
    unsigned long long opt_gen = 0;          // bumped by svdb_set_option / svdb_set_stream
    size_t last_nq = 0, last_k = 0;          // a shape is captured the second time it shows up in a row
    void drop_graphs();

    svdb::TieGpu *tie = nullptr;             // created by the first tie walk, kept for the next ones

    svdb::ScanTuning tune;
    bool force_exact = false;
    bool profile_scan = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> scan_events;
    size_t scan_events_used = 0;
    svdb_stats stats{};

    const double *kd_ptr() const { return alias ? rows.as<double>() : kdpts.as<double>(); }
    size_t max_versions = 0;

    // all return svdb_status; callers hold mu
    int init(const svdb_config &c);
    void destroy();
    int stage_one(const double *row, size_t ncopy, uint64_t index);
    int flush();
    int upload_cur();
    int tree_append(size_t n0, size_t m);
    bool mtree_wanted(size_t k, int mode) const;
    int mtree_update();                  // (re)build the median tree when the unindexed tail has outgrown its limit
    // x != NULL (sharded store): d_out receives the MERGED answers of all shards (collective, one exchange epoch per call)
    int nearest_device(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out, int mode,
                       svdb_exchange *x = nullptr);
    int nearest_local(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out, int mode, svdb_exchange *x,
                      svdb_candidate *d_merged, bool *exchanged);
    // x != NULL: this engine is one shard; the local candidates go through the peer-memory exchange and the
    // outputs are the merged answers (collective: every rank calls with the same nq and k)
    int nearest_host(const double *Q, size_t nq, size_t ldq, size_t k, size_t *index_out, double *dist_out,
                     uint64_t *seq_out, svdb_candidate *cand_out = nullptr, svdb_exchange *x = nullptr);
    int ingest_device_rows(const double *d_rows, size_t n, size_t ld, size_t *first_index);
    int compare_device(int mode, const uint64_t *d_i1, const uint64_t *d_i2, size_t n, float *d_out);
    int compare_host(int mode, const size_t *i1, const size_t *i2, size_t n, float *out);
    int fail_cuda(const char *what, cudaError_t e);
    int fail(int code, const std::string &msg);
};
