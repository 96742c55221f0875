// kernels.h -- host-callable launchers of the sm_100a kernels (definitions in *.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/svdb_b200.h"

namespace svdb {

typedef unsigned long long u64;

struct Cand;  // 16-byte (d, seq) record, see common.cuh

// Tunables of the scan launch (engine options "scan.*").
struct ScanTuning {
    int variant = 0;        // 0: TMA bulk-copy ring per warp; 1: direct LDG.128 streaming
    int warps = 4;          // warps per CTA (bench A/B: 2 CTAs/SM x 4 warps x 2 stages beats 1 x 8 x 3 by 1.3 %)
    int stages = 2;         // ring depth per warp (variant 0)
    int tile_rows = 0;      // rows per tile, 0 = choose from the row size
    int ctas_per_sm = 2;
    int assign = 0;         // 0: tiles dealt round-robin over all warps; 1: every CTA streams one contiguous slab
    int nq_per_pass = 4;    // queries sharing one pass over the log (1, 2, 4, 8)
    int thin_max_k = 12;    // K <= this: thread-per-row exact kernel is the primary path (measured 0.84-0.92 x HBM peak
                            // at K = 9..12 but 0.33 x at K = 16, where 128-byte rows make every lane hit its own line)
    int num_sms = 148;
};

// ---- what follows a scan: finalize (merge of the per-CTA lists, reference-order re-rank, completeness proof) ----
struct FinalArgs {
    const Cand *lists;
    int nlists, cap, nq, k;
    const double *pts;
    int K, stride;
    const double *q;
    int ldq;
    const u64 *log_index;   // index carried by each log entry
    u64 seq_base;
    double eps;             // relative error bound of the approximate keys; < 0: keys are exact
    double eabs_coef;       // GEMM-form keys (K2): absolute error bound = eabs_coef * (xn_max + |q|^2); 0 otherwise
    const double *qnorm;    // |q|^2 per query of this launch (K2), else NULL
    const unsigned long long *xn_max_bits;   // largest |x|^2 in the log, as double bits (K2), else NULL
    double scale_lo, scale_hi;   // K10 (fp32 keys): queries whose max|x|^2 + |q|^2 lies outside [lo, hi] are flagged UNSAFE; 0, 0: no check
    // sqrt-form keys (K12, the scan of ONE low-precision plane of the log): key = |x^ - q^|^2 in fp32, hence
    //   |sqrt(key) - sqrt(d)| <= E + gamma (sqrt(d) + E),   E = max_r |x_r - x^_r| + |q - fl32(q)| + underflow slack
    // sq_mode = 1: eps / eabs_coef / qnorm are ignored; finalize forms |q - fl32(q)| and |q|^2 itself
    int sq_mode;            // 1: K12 (q^ = fl32(q)); 2: K13 (q^ on the byte plane's query grid, plane8 != NULL)
    double sq_gamma;
    const unsigned long long *plane_err_bits;   // max_r |x_r - x^_r| over the log, as double bits
    const struct Plane8Par *plane8;             // K13: the grid the queries are quantised on
    const uint32_t *child;  // reference-shaped tree links (tree.cuh) for exact tie order; NULL: ties -> lowest seq
    int mark_ties;          // shards (child == NULL): flag SVDB_CAND_TIE when distinct kd-points may tie at the minimum
    svdb_candidate *out;    // [nq][k]
};
cudaError_t launch_finalize(const FinalArgs &a, cudaStream_t st);

// ---- the fused tail: the LAST CTA of a scan launch to finish (atomic ticket) runs finalize for the launch's queries
// itself and, on a sharded store, stores the answers into every peer's gather buffer over NVLink, waits for the
// peers' and merges (K7) -- the whole single-query step is ONE launch (the fixed ~44 us of
// finalize + push + merge launches behind a 1 ms scan).
constexpr int XCH_MAX_WORLD = 16;
struct PeerPtrs {
    unsigned char *p[XCH_MAX_WORLD];
};
struct TailArgs {
    unsigned *ticket;       // device word, zero between launches; NULL: no fused tail (finalize_kernel is launched after the scan)
    FinalArgs fin;          // fin.lists = the launch's lists
    int world, rank;        // world <= 1: no exchange, fin.out is the answer
    PeerPtrs peers;         // exchange.cu: every rank's gather buffer
    size_t max_rec;
    svdb_candidate *xout;   // merged answers [nq][k]
    unsigned long long *dbg; // NULL, or 32 + gridDim.x words of %globaltimer stamps (option "scan.tail_debug"):
                            // [0] last CTA took its ticket, [1] lists merged, [2] re-rank done (finalize end), [3] pushed to
                            // the peers, [4] every peer's data has landed, [5] merged; [32 + b] CTA b took its ticket; [9..14] phases inside finalize
};

struct ScanArgs {
    const double *pts;      // log: entry s at pts + s * stride
    u64 n;                  // entries
    int K;                  // coordinates measured
    int stride;             // doubles between entries
    const double *q;        // device queries, query i at q + i * ldq (zero padded to stride for the wide path)
    int ldq;
    int nq;                 // queries in this pass (<= 8)
    int cap;                // candidates each CTA emits per query
    Cand *lists;            // [nq][nlists][cap]
    int assign;             // ScanTuning::assign
    TailArgs tail;          // scan_wide_kernel only
};

// Number of per-CTA lists a scan with this tuning writes per query.
int scan_num_lists(const ScanTuning &t, bool wide);
// Approximate (FMA, lane-parallel) scan for wide rows; needs stride even and 16-B aligned pts.
cudaError_t launch_scan_wide(const ScanTuning &t, const ScanArgs &a, cudaStream_t st);
// Reference-order exact scan, one thread per log entry (primary for thin rows, fallback for any).
cudaError_t launch_scan_exact(const ScanTuning &t, const ScanArgs &a, cudaStream_t st);

// K11: the scan over the split-bf16 shadow of the log (4 bytes per coordinate instead of 8)
struct ShadowScanArgs {
    const uint16_t *xhi, *xlo; // [n][Kp] bf16 each: hi plane, lo plane (launch_split_bf16)
    u64 n;
    int K, Kp;
    const double *q;        // device queries (fp64), query i at q + i * ldq
    int ldq;
    int nq;                 // 1, 2, 4 or 8 queries share the pass
    int cap;
    Cand *lists;            // [nq][nlists][cap]
    TailArgs tail;
};
cudaError_t launch_scan_shadow(const ScanTuning &t, const ShadowScanArgs &a, cudaStream_t st);
double shadow_eps(int K);
double shadow_eabs_coef();

// K12: the single-query scan over the bf16 HI plane alone (2 bytes per coordinate); sqrt-form bound (FinalArgs::sq_mode)
struct PlaneScanArgs {
    const uint16_t *xhi;    // [n][Kp] bf16
    u64 n;
    int K, Kp;
    const double *q;        // device queries (fp64), query i at q + i * ldq; K coordinates each are read
    int ldq;
    int nq;                 // 1 or 2 queries share the pass (plane_scan_supports)
    int dyn_eighths;        // 0: tiles dealt round-robin (default); 1..8: that many eighths of the log handed out dynamically
    int cap;
    Cand *lists;            // [nq][nlists][cap]
    TailArgs tail;
    int grid;               // CTAs (= lists) to launch; 0: scan_num_lists()
    int pdl;                // see Plane8ScanArgs
};
bool plane_scan_supports(int Kp, int nq);
double plane_gamma(int Kp);
cudaError_t launch_scan_plane(const ScanTuning &t, const PlaneScanArgs &a, cudaStream_t st);

// K13: the single-query scan over a ONE-BYTE plane of the log (x^ = lo + step * u on one store-wide grid; 1 byte per
// coordinate).  Keys |x^ - q^|^2 are formed EXACTLY from integer dot products (dp4a); sqrt-form bound, FinalArgs::sq_mode = 2.
struct Plane8Par {          // device-resident parameters of the grid
    double lo, step;        // fixed when the plane is first built (rows appended later are clamped; their error is measured)
    unsigned long long min_ord, max_ord;   // running range of the log while the grid is being chosen (order-preserving bit patterns)
};
struct Plane8ScanArgs {
    const unsigned char *x8; // [n][Kp] bytes
    const Plane8Par *par;
    u64 n;
    int K, Kp;
    const double *q;        // device queries (fp64), query i at q + i * ldq; K coordinates each are read
    int ldq;
    int nq;                 // 1, or 2 queries sharing the pass (plane8_scan_supports_two)
    int cap;
    Cand *lists;            // [nq][nlists][cap]
    TailArgs tail;
    int grid;               // CTAs (= lists) to launch; 0: scan_num_lists()
    int pdl;                // 1: launched with programmatic stream serialization -- the scan may start while the tail of the launch
                            // in front of it still runs; everything mutable is touched only after pdl_wait(), and the query,
                            // read before it, is re-checked after it (tail.ticket[3] = "a query changed under a scan")
};
bool plane8_scan_supports(int Kp);
bool plane8_scan_supports_two(int Kp);
cudaError_t launch_scan_plane8(const ScanTuning &t, const Plane8ScanArgs &a, cudaStream_t st);
// rows [first, first+n) of the log -> the byte plane; choose_grid: first take lo / step from the range of those rows.
// err_bits: running max over the rows of |x - x^|_2 (double bits).
cudaError_t launch_plane8_build(const double *src, int ld, int K, int Kp, u64 first, u64 n, Plane8Par *par, bool choose_grid,
                                unsigned char *dst, unsigned long long *err_bits, int num_sms, cudaStream_t st);



cudaError_t launch_merge_candidates(const svdb_candidate *in, int nshards, int nq, int k, svdb_candidate *out,
                                    cudaStream_t st);

// K2: batched queries on the FP64 tensor cores (mma_kernels.cu)
struct MmaArgs {
    const double *pts;      // log rows, stride doubles apart (even, 16-byte aligned rows, zero padded)
    u64 n;
    int K, stride;
    const double *xnorm;    // |x_r|^2 per log entry
    const double *q;        // padded queries [ngroups*group][ldq], zeros beyond K and beyond nq
    const double *qnorm;    // |q|^2 per padded query
    int ldq, nq;            // nq real queries
    int group;              // queries per CTA group: 64, 32 or 16 (mma_group_size)
    int ngroups, nstreams;  // grid = ngroups * nstreams CTAs
    int cap;
    Cand *lists;            // [ngroups*group][nstreams][cap]
};
cudaError_t launch_scan_mma(const MmaArgs &a, cudaStream_t st);
int mma_group_size(size_t nq);
cudaError_t launch_rownorm(const double *pts, int stride, int K, u64 first, u64 n, double *out, unsigned long long *max_bits,
                           int num_sms, cudaStream_t st);
cudaError_t launch_prep_queries(const double *src, int ldq, int K, int nq, int nq_pad, double *dst, int ldp, double *qnorm,
                                cudaStream_t st);

// K10: batched queries on the 5th-generation tensor cores (tcgen05 + TMEM) with split-bf16 keys (umma_filter.cu)
struct UmmaArgs {
    const uint16_t *xhi, *xlo; // [n][Kp] bf16 each: hi plane and lo plane of the log rows (launch_split_bf16)
    u64 n;                  // log entries, < 2^31
    int K, Kp;              // Kp = umma_kpad(K)
    const double *xnorm;    // |x_r|^2 per log entry (fp64, as for K2)
    const uint16_t *qhi, *qlo; // [ngroups*bn][Kp] bf16 each: the planes of the padded queries
    const double *qnorm;    // |q|^2 per padded query
    int nq;                 // real queries
    int bn;                 // queries per CTA group: 64, 128 or 256 (umma_group_size)
    int ngroups, nstreams;  // grid = ngroups * nstreams CTAs
    int cap;
    Cand *lists;            // [ngroups*bn][nstreams][cap]
    void *bufs;             // umma_buf_bytes(ngroups, nstreams, bn) of scratch
    uint32_t *gtau;         // [ngroups*bn] words, all 0xffffffff at launch: per query the smallest cap-th key published so far
    int sparse_checks;      // 1: the per-tile bookkeeping between the CTA barriers runs every fourth tile once thresholds are tight
    uint32_t *gmin;         // [ngroups*bn][nstreams] words behind gtau, all 0xffffffff at launch: per query and CTA of its group the
                            // smallest key that CTA has kept so far (NULL: thresholds from gtau alone)
    float *dbg_keys;        // NULL, or [128][bn]: the keys of rows 0..127 against the first query group (diagnostics)
    int qres;               // 0: query planes streamed next to the rows; else umma_resident_stages(bn, Kp): they stay in shared
                            // memory and the ring carries rows only, that many stages deep
};
int umma_resident_stages(int bn, int Kp);
int umma_kpad(int K);
int umma_group_size(size_t nq);
size_t umma_buf_bytes(int ngroups, int nstreams, int bn);
double umma_eabs_coef(int K);   // absolute key error <= coef * (max|x|^2 + |q|^2)
// fp64 rows [first, first+n) -> the hi / lo bf16 planes; err_bits (may be NULL): running max over the rows of |x - hi|_2 (double bits)
cudaError_t launch_split_bf16(const double *src, int ld, int K, int Kp, u64 first, u64 n, uint16_t *dst_hi, uint16_t *dst_lo,
                              unsigned long long *err_bits, int num_sms, cudaStream_t st);
cudaError_t launch_umma_filter(const UmmaArgs &a, cudaStream_t st, std::string *why);

// K5: insert log entries [n0, n0+m) into the reference-shaped tree (level-synchronous; see tree_kernels.cu).
// pn/pds: m-entry u32 scratch; d_flag: device word; h_flag_pinned: pinned host word. Synchronizes the stream.
// Gives up after max_rounds rounds (tree deeper than that: degenerate insertion order) and reports a
// NEGATIVE round count; the tree is then incomplete and must not be used.
cudaError_t launch_tree_insert(const double *pts, int stride, int K, uint32_t *child, u64 n0, u64 m, uint32_t *pn,
                               uint32_t *pds, unsigned *d_flag, unsigned *h_flag_pinned, int num_sms, cudaStream_t st,
                               int max_rounds, int *rounds_out);
// K6: the reference's traversal, one thread per query (k = 1), or its k-smallest generalisation (k > 1).
// only_marked (device array of nq words, or NULL): answer only the queries whose word is non-zero (K9 left a
// distinct-point tie there), leave the other entries of out[] as they are.
cudaError_t launch_tree_nearest(const double *pts, int stride, int K, const uint32_t *child, u64 n, const double *Q,
                                int ldq, int nq, int k, const u64 *log_index, u64 seq_base, svdb_candidate *out,
                                cudaStream_t st, const unsigned *only_marked = nullptr);

// K8/K9: balanced median KD tree over thin kd-points (median_tree.cu).  Implicit layout: node (level l, j) has heap id
// 2^l + j and covers positions [(j*n)>>l, ((j+1)*n)>>l) of the leaf-ordered point array; only split values are stored.
struct MtreeView {
    const double *split = nullptr;    // split values; block_levels = 1: [2^levels] in heap order (entry 0 unused);
                                      // 3: 64-byte blocks of three levels each (median_tree.cu: mt_slot)
    int block_levels = 1;
    const double *mpts = nullptr;     // [n_built][K] kd-points in leaf order
    const uint32_t *mseq = nullptr;   // [n_built] log sequence number of each
    u64 n_built = 0;                  // log entries [0, n_built) are in the tree
    int levels = 0;                   // internal levels; 2^levels leaves of <= 32 points
};
int mtree_levels(u64 n);
size_t mtree_split_count(u64 n, int block_levels);
// Build over log entries [0, n): median splits by radix-select partitioning (large segments) and in-CTA sorts (segments
// of <= 2048).  split/mpts/mseq: mtree_split_count(n, block_levels) doubles, n*K doubles, n u32.  Allocates and frees its own scratch
// (~17 bytes per entry) and synchronizes the stream.
cudaError_t launch_mtree_build(const double *pts, int stride, int K, u64 n, int block_levels, double *split, double *mpts, uint32_t *mseq,
                               int num_sms, cudaStream_t st, int *levels_out, int *launches_out);
// K <= 8.  k = 1: `lanes` (32, 16 or 8) lanes per query; k > 1: a warp per query keeps the k smallest (distance, seq),
// nq x k candidates.  Entries [t.n_built, n) of the raw log are scanned after the tree.
// Answers: smallest (reference-order distance, seq); SVDB_CAND_TIE (if mark_ties) when entries with different
// coordinates tie at the minimum -- only then can the reference's answer differ (rerun those through K6);
// marks (device, nq words, may be NULL) receives 1 for every flagged query, else 0.
cudaError_t launch_mtree_nearest(const MtreeView &t, const double *pts, int stride, int K, u64 n, const double *Q, int ldq,
                                 int nq, int k, const u64 *log_index, u64 seq_base, int mark_ties, int lanes, unsigned *marks,
                                 svdb_candidate *out, cudaStream_t st);

struct CompareArgs {
    const double *rows;     // version rows, row s at rows + s * ldr; ldr % 16 == 0, zero padded
    int ldr, D;
    const u64 *cur;         // index -> version row; NULL: pair members are version rows already
    u64 nrows;              // valid indices are < nrows
    const u64 *i1, *i2;     // NULL (mode 4 only): member = first + pair id
    u64 first;
    u64 n;                  // pairs
    const float *norm;      // float-order self dot per version row (cosine)
    float *out;             // n floats (modes 0,1,2,4) or n x 3 (mode 3)
    int mode;               // 0 cosine, 1 euclidean, 2 dot, 3 all three, 4 self-dot (norm precompute)
};
cudaError_t launch_compare(const CompareArgs &a, int num_sms, cudaStream_t st);

// Copy the first K coordinates of n rows (stride ld_src) into a compact array (stride ld_dst), zero padding.
cudaError_t launch_extract_prefix(const double *src, int ld_src, double *dst, int ld_dst, int K, u64 n,
                                  cudaStream_t st);
// dst[i] = base + i  (index carried by freshly inserted log entries)
cudaError_t launch_iota(u64 *dst, u64 base, u64 n, cudaStream_t st);
// Pad user queries (nq x ldq, K used) into nq x ldp with zeros.
cudaError_t launch_pad_queries(const double *src, int ldq, double *dst, int ldp, int K, int nq, cudaStream_t st);

}  // namespace svdb
