// engine.cu -- one shard of the store: arenas, delta staging, query orchestration, C-ABI.
#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace svdb {

static thread_local std::string g_last_error;
void set_last_error(const std::string &s) { g_last_error = s; }
const std::string &get_last_error() { return g_last_error; }

static std::atomic<unsigned long long> g_scratch_gen{0};
unsigned long long scratch_generation() { return g_scratch_gen.load(); }

bool Scratch::ensure(size_t bytes, std::string &err) {
    if (bytes <= cap) return true;
    g_scratch_gen++;
    size_t want = std::max(bytes, cap + cap / 2);
    void *np = nullptr;
    if (cudaMalloc(&np, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&np, want) != cudaSuccess) {
            err = std::string("cudaMalloc(scratch) failed: ") + cudaGetErrorString(cudaGetLastError());
            return false;
        }
    }
    if (p) cudaFree(p);   // implicit device sync: nothing in flight may still read the old block
    p = np;
    cap = want;
    return true;
}
void Scratch::free_() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
bool PinnedScratch::ensure(size_t bytes, std::string &err) {
    if (bytes <= cap) return true;
    g_scratch_gen++;
    const size_t want = std::max(bytes, cap + cap / 2);
    void *np = nullptr;
    if (cudaMallocHost(&np, want) != cudaSuccess) {
        err = std::string("cudaMallocHost failed: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    if (p) cudaFreeHost(p);
    p = np;
    cap = want;
    return true;
}
void PinnedScratch::free_() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

// The reference's save file packs rows as { char uuid[37]; u64 dim; f64 data[D] }: 45 + 8D bytes per
// record, so the doubles are never 8-byte aligned.  Unpack / pack them on the device, byte-wise.
constexpr size_t REC_HDR = 37 + 8;
__global__ void unpack_records_kernel(const unsigned char *__restrict__ raw, size_t rec_bytes, int D, size_t n,
                                      double *__restrict__ rows) {
    const size_t total = n * (size_t)D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / D, c = i % D;
        const unsigned char *src = raw + r * rec_bytes + REC_HDR + c * 8;
        unsigned long long v = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) v |= (unsigned long long)src[b] << (8 * b);
        rows[i] = __longlong_as_double((long long)v);
    }
}
__global__ void pack_records_kernel(const double *__restrict__ rows, int ldr, const unsigned long long *__restrict__ cur,
                                    size_t first, int D, size_t n, size_t rec_bytes, unsigned char *__restrict__ raw) {
    const size_t total = n * (size_t)D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / D, c = i % D;
        const unsigned long long v = (unsigned long long)__double_as_longlong(rows[cur[first + r] * (size_t)ldr + c]);
        unsigned char *dst = raw + r * rec_bytes + REC_HDR + c * 8;
#pragma unroll
        for (int b = 0; b < 8; b++) dst[b] = (unsigned char)(v >> (8 * b));
    }
}

__global__ void fill_f32_kernel(float *dst, float v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

}  // namespace svdb

using namespace svdb;

#define CK(call)                                            \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return fail_cuda(#call, e__); \
    } while (0)

int svdb_engine::fail_cuda(const char *what, cudaError_t e) {
    set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaGetLastError();
    return SVDB_ERR_CUDA;
}
int svdb_engine::fail(int code, const std::string &msg) {
    set_last_error(msg);
    return code;
}

int svdb_engine::init(const svdb_config &c) {
    cfg = c;
    if (c.dimension < 1 || c.kd_dim < 1) return fail(SVDB_ERR_ARG, "dimension and kd_dim must be >= 1");
    if (c.kd_dim > c.dimension)
        return fail(SVDB_ERR_ARG, "kd_dim > dimension: the reference would read past the row (kdtree.c:26-28)");
    if (c.dimension > (1u << 20)) return fail(SVDB_ERR_ARG, "dimension too large");
    log_only = (c.flags & SVDB_FLAG_LOG_ONLY) != 0;
    no_log = (c.flags & SVDB_FLAG_NO_LOG) != 0;
    if (log_only && no_log) return fail(SVDB_ERR_ARG, "LOG_ONLY and NO_LOG exclude each other");
    if (log_only && c.kd_dim != c.dimension) return fail(SVDB_ERR_ARG, "log-only engine needs dimension == kd_dim");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SVDB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                       " (this library has no CPU path)");
    if (c.device < 0 || c.device >= count) return fail(SVDB_ERR_ARG, "device ordinal out of range");
    device = c.device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SVDB_ERR_CUDA, "kernels are built for sm_100a only; this device is older");
    tune.num_sms = prop.multiProcessorCount;
    if (const char *v = getenv("SVDB_SCAN_VARIANT")) tune.variant = atoi(v);
    if (const char *v = getenv("SVDB_THIN_MAX_K")) tune.thin_max_k = atoi(v);     // A/B of the exact vs the wide scan at small kd_dim

    D = (int)c.dimension;
    K = (int)c.kd_dim;
    Dpad = (int)round_up(D, 16);
    wide = K > tune.thin_max_k;
    alias = !log_only && !no_log && wide && K == D;
    kstride = alias ? Dpad : (wide ? (int)round_up(K, 2) : K);

    const size_t va = (size_t)256 << 30;
    const size_t main_row_bytes = (size_t)(log_only ? kstride : Dpad) * 8;
    max_versions = std::min<size_t>(va / main_row_bytes, 0xfffffffeull);   // tree links are u32
    use_tree = !no_log && !(c.flags & SVDB_FLAG_SHARD) && !getenv("SVDB_NO_TREE");
    graphs_enabled = !getenv("SVDB_NO_GRAPH");
    use_mtree = !no_log && !wide && K <= 8 && !getenv("SVDB_NO_MTREE");
    if (const char *v = getenv("SVDB_MTREE")) mtree_auto = atoi(v);
    if (const char *v = getenv("SVDB_MTREE_LANES")) mtree_lanes = atoi(v);
    if (const char *v = getenv("SVDB_MTREE_BLOCK")) mtree_block = atoi(v) == 1 ? 1 : 3;
    std::string err;
    if (!log_only) {
        if (!rows.init(device, max_versions * (size_t)Dpad * 8, err)) return fail(SVDB_ERR_CUDA, err);
        if (!norms.init(device, max_versions * 4, err)) return fail(SVDB_ERR_CUDA, err);
        if (!cur.init(device, max_versions * 8, err)) return fail(SVDB_ERR_CUDA, err);
    }
    if (!alias && !no_log && !kdpts.init(device, max_versions * (size_t)kstride * 8, err)) return fail(SVDB_ERR_CUDA, err);
    if (!no_log && !log_idx.init(device, max_versions * 8, err)) return fail(SVDB_ERR_CUDA, err);
    if (wide && !no_log) {
        if (!xnorm.init(device, max_versions * 8, err)) return fail(SVDB_ERR_CUDA, err);
        if (!xnmax.ensure(8, err)) return fail(SVDB_ERR_OOM, err);
        CK(cudaMemset(xnmax.p, 0, 8));
    }
    if (use_tree) {
        if (!child.init(device, max_versions * 8, err)) return fail(SVDB_ERR_CUDA, err);
        if (!tree_flag.ensure(16, err) || !tree_hflag.ensure(16, err)) return fail(SVDB_ERR_OOM, err);
    }

    CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    // the fused tail's arrival counter: zero between launches.  Allocated and cleared HERE, never lazily: a first use
    // inside a stream capture (nearest_host) would record the memset into a graph that is then thrown away
    if (!no_log) {
        if (!ticket.ensure(16, err)) return fail(SVDB_ERR_OOM, err);
        CK(cudaMemset(ticket.p, 0, 16));
    }

    stage_ld = log_only ? (size_t)kstride : (size_t)Dpad;
    stage_cap = std::min<size_t>(65536, std::max<size_t>(1, ((size_t)8 << 20) / (stage_ld * 8)));
    if (!stage_rows.ensure(stage_cap * stage_ld * 8, err) || !stage_idx.ensure(stage_cap * 8, err))
        return fail(SVDB_ERR_OOM, err);

    if (c.reserve_rows) {
        const size_t n = std::min(c.reserve_rows, max_versions);
        bool ok = true;
        if (!log_only) ok = ok && rows.ensure(n * (size_t)Dpad * 8, stream, err) && norms.ensure(n * 4, stream, err) &&
                            cur.ensure(n * 8, stream, err);
        if (!alias && !no_log) ok = ok && kdpts.ensure(n * (size_t)kstride * 8, stream, err);
        if (!no_log) ok = ok && log_idx.ensure(n * 8, stream, err);
        if (use_tree) ok = ok && child.ensure(n * 8, stream, err);
        if (!ok) return fail(SVDB_ERR_OOM, err);
        cur_host.reserve(n);
    }
    return SVDB_OK;
}

void svdb_engine::drop_graphs() {
    for (auto &g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
}

void svdb_engine::destroy() {
    cudaSetDevice(device);
    if (own_stream) cudaStreamSynchronize(own_stream);
    drop_graphs();
    for (auto &ev : scan_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    scan_events.clear();
    rows.release();
    kdpts.release();
    log_idx.release();
    norms.release();
    cur.release();
    child.release();
    xnorm.release();
    shadow_hi.release();
    shadow_lo.release();
    plane8.release();
    plane8_ready = false;
    plane8_n = 0;
    shadow_ready = false;
    shadow_n = shadow_lo_n = 0;
    shadow_mapped_counted = 0;
    for (Scratch *s : {&qsplit, &ubuf, &udbg, &plane_err, &ticket, &xlocal, &tail_dbg, &plane8_par}) s->free_();
    for (Scratch *s : {&qpad, &qraw, &lists, &outc, &idx1, &idx2, &fout, &tree_pn, &tree_pds, &tree_flag, &qnorm, &xnmax,
                       &mt_split, &mt_pts, &mt_seq, &mt_marks}) s->free_();
    for (PinnedScratch *s : {&stage_rows, &stage_idx, &hq, &hout, &hidx, &hf, &tree_hflag}) s->free_();
    tie_state_free(tie);
    tie = nullptr;
    if (own_stream) cudaStreamDestroy(own_stream);
    own_stream = stream = nullptr;
}

int svdb_engine::stage_one(const double *row, size_t ncopy, uint64_t index) {
    if (stage_n == stage_cap) {
        int rc = flush();
        if (rc) return rc;
    }
    if (n_versions + stage_n >= max_versions) return fail(SVDB_ERR_OOM, "store exceeds the reserved address range");
    double *dst = stage_rows.as<double>() + stage_n * stage_ld;
    memcpy(dst, row, ncopy * sizeof(double));
    if (ncopy < stage_ld) memset(dst + ncopy, 0, (stage_ld - ncopy) * sizeof(double));
    stage_idx.as<uint64_t>()[stage_n] = index;
    stage_n++;
    return SVDB_OK;
}

// K5: link log entries [n0, n0+m) into the reference-shaped tree. Their kd-points must be resident.
int svdb_engine::tree_append(size_t n0, size_t m) {
    if (m == 0) return SVDB_OK;
    std::string err;
    if (wide && !no_log) {   // |x|^2 for the GEMM-form keys of the batched path (K2)
        if (!xnorm.ensure((n0 + m) * 8, stream, err)) return fail(SVDB_ERR_OOM, err);
        CK(launch_rownorm(kd_ptr(), kstride, K, n0, m, xnorm.as<double>(), xnmax.as<unsigned long long>(), tune.num_sms, stream));
        stats.kernels_launched++;
    }
    if (!use_tree) return SVDB_OK;
    if (!child.ensure((n0 + m) * 8, stream, err) || !tree_pn.ensure(m * 4, err) || !tree_pds.ensure(m * 4, err))
        return fail(SVDB_ERR_OOM, err);
    int rounds = 0;
    CK(launch_tree_insert(kd_ptr(), kstride, K, child.as<uint32_t>(), n0, m, tree_pn.as<uint32_t>(), tree_pds.as<uint32_t>(),
                          tree_flag.as<unsigned>(), tree_hflag.as<unsigned>(), tune.num_sms, stream, tree_max_depth, &rounds));
    if (rounds < 0) {
        // The reference's tree for this insertion order is deeper than tree_max_depth (sorted input makes it
        // a linked list: its own recursive insert/search would be O(N) deep).  Keep serving from the scan:
        // exact nearest neighbours, exact-distance ties between distinct points fall back to the lowest seq.
        fprintf(stderr, "svdb_b200: KD tree deeper than %d levels (degenerate insertion order); tree dropped, "
                        "nearest is answered by the scan from now on\n", tree_max_depth);
        use_tree = false;
        stats.tree_dropped = 1;              // svdb_get_stats: callers can see that exact-tie order is no longer the reference's
        rounds = -rounds;
    }
    stats.tree_rounds += rounds;
    stats.kernels_launched += rounds + 1;
    return SVDB_OK;
}

bool svdb_engine::mtree_wanted(size_t k, int mode) const {
    if (k < 1 || k > SVDB_MAX_K || !use_mtree) return false;
    return mode == SVDB_MODE_MTREE || (mode == SVDB_MODE_AUTO && mtree_auto && !force_exact && K <= tree_max_k);
}

// K8: the median tree covers log entries [0, mt.n_built); what was appended since is scanned as a tail by K9.
// Rebuild (from scratch: the tree is static) once the tail exceeds clamp(n_built / 8, tail_min, tail_max).
int svdb_engine::mtree_update() {
    if (!use_mtree) return SVDB_OK;
    if (mt.n_built && mt.block_levels != mtree_block) mt = MtreeView{};      // layout option changed: rebuild
    const size_t n = n_versions, tail = n - (size_t)mt.n_built;
    const size_t limit = std::min(std::max((size_t)mt.n_built / 8, mtree_tail_min), std::max(mtree_tail_max, mtree_tail_min));
    if (tail <= limit) return SVDB_OK;
    std::string err;
    if (!mt_split.ensure(mtree_split_count(n, mtree_block) * 8, err) || !mt_pts.ensure(n * (size_t)K * 8, err) || !mt_seq.ensure(n * 4, err))
        return fail(SVDB_ERR_OOM, err);
    int levels = 0, launches = 0;
    mt = MtreeView{};                    // nothing usable until the build has finished
    CK(launch_mtree_build(kd_ptr(), kstride, K, n, mtree_block, mt_split.as<double>(), mt_pts.as<double>(), mt_seq.as<uint32_t>(),
                          tune.num_sms, stream, &levels, &launches));
    mt.split = mt_split.as<double>();
    mt.block_levels = mtree_block;
    mt.mpts = mt_pts.as<double>();
    mt.mseq = mt_seq.as<uint32_t>();
    mt.n_built = n;
    mt.levels = levels;
    opt_gen++;                           // captured graphs baked the old view in
    stats.mtree_builds++;
    stats.mtree_levels = (uint64_t)levels;
    stats.mtree_rows = n;
    stats.kernels_launched += launches;
    return SVDB_OK;
}

int svdb_engine::flush() {
    if (stage_n == 0) return SVDB_OK;
    pdl_mark = ~0ull;                          // whatever follows new rows starts fully ordered (engine.h: overlap_steps)
    CK(cudaSetDevice(device));
    const size_t n0 = n_versions, m = stage_n, n1 = n0 + m;
    std::string err;
    bool ok = no_log || log_idx.ensure(n1 * 8, stream, err);
    if (!log_only) ok = ok && rows.ensure(n1 * (size_t)Dpad * 8, stream, err) && norms.ensure(n1 * 4, stream, err);
    if (!alias && !no_log) ok = ok && kdpts.ensure(n1 * (size_t)kstride * 8, stream, err);
    if (!ok) return fail(SVDB_ERR_OOM, err);

    double *dst_rows = log_only ? kdpts.as<double>() + n0 * (size_t)kstride : rows.as<double>() + n0 * (size_t)Dpad;
    CK(cudaMemcpyAsync(dst_rows, stage_rows.p, m * stage_ld * 8, cudaMemcpyHostToDevice, stream));
    if (!no_log) CK(cudaMemcpyAsync(log_idx.as<uint64_t>() + n0, stage_idx.p, m * 8, cudaMemcpyHostToDevice, stream));
    stats.h2d_bytes += m * stage_ld * 8 + m * 8;
    if (!log_only) {
        if (!alias && !no_log) {
            CK(launch_extract_prefix(rows.as<double>() + n0 * (size_t)Dpad, Dpad, kdpts.as<double>() + n0 * (size_t)kstride,
                                     kstride, K, m, stream));
            stats.kernels_launched++;
        }
        CompareArgs ca{};
        ca.rows = rows.as<double>();
        ca.ldr = Dpad;
        ca.D = D;
        ca.first = n0;
        ca.n = m;
        ca.out = norms.as<float>() + n0;
        ca.mode = 4;
        CK(launch_compare(ca, tune.num_sms, stream));
        stats.kernels_launched++;
    }
    {
        int rc = tree_append(n0, m);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(stream));   // staging buffers are reused by the caller
    n_versions = n1;
    stage_n = 0;
    stats.hbm_bytes_mapped = rows.mapped() + kdpts.mapped() + log_idx.mapped() + norms.mapped() + cur.mapped() + child.mapped() + xnorm.mapped() + shadow_hi.mapped() + shadow_lo.mapped() + plane8.mapped();
    shadow_mapped_counted = shadow_hi.mapped() + shadow_lo.mapped();
    plane8_mapped_counted = plane8.mapped();
    return SVDB_OK;
}

int svdb_engine::upload_cur() {
    const size_t n = cur_host.size();
    if (cur_dirty_lo >= n) {
        cur_dirty_lo = n;
        return SVDB_OK;
    }
    std::string err;
    if (!cur.ensure(n * 8, stream, err)) return fail(SVDB_ERR_OOM, err);
    CK(cudaMemcpyAsync(cur.as<uint64_t>() + cur_dirty_lo, cur_host.data() + cur_dirty_lo, (n - cur_dirty_lo) * 8,
                       cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));   // pageable source: do not let the host vector change under the copy
    stats.h2d_bytes += (n - cur_dirty_lo) * 8;
    cur_dirty_lo = n;
    return SVDB_OK;
}

static int largest_pass(size_t remaining, int limit) {
    int p = 1;
    while (p * 2 <= limit && (size_t)(p * 2) <= remaining && p * 2 <= 8) p *= 2;
    return p;
}

// Sharded stores: the local answers go through the peer-memory exchange (exchange.cu) and d_out receives the merged
// ones.  A call that is ONE scan pass (the single-query step) does all of it inside the scan launch (scan_tail);
// everything else exchanges once, after its local answers are complete -- one epoch per call on every path, so the
// ranks stay in step even if they took different paths.
int svdb_engine::nearest_device(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out, int mode,
                                svdb_exchange *x) {
    if (!x) return nearest_local(d_Q, nq, ldq, k, d_out, mode, nullptr, nullptr, nullptr);
    if (!exchange_fits(x, nq, k)) return fail(SVDB_ERR_ARG, "nq * k exceeds the exchange capacity");
    if (nq == 0) return SVDB_OK;
    std::string err;
    if (!xlocal.ensure(nq * k * sizeof(svdb_candidate), err)) return fail(SVDB_ERR_OOM, err);
    bool exchanged = false;
    if (!d_out) return fail(SVDB_ERR_ARG, "NULL output");
    int rc = nearest_local(d_Q, nq, ldq, k, xlocal.as<svdb_candidate>(), mode, x, d_out, &exchanged);
    if (rc) return rc;
    if (!exchanged) {
        CK(exchange_enqueue(x, stream, xlocal.as<svdb_candidate>(), nq, k, d_out));
        stats.kernels_launched += 2;
    }
    return SVDB_OK;
}

// d_out: this engine's own answers.  x != NULL: the caller will exchange them and merge into d_merged afterwards; a
// call that is a single scan pass does that itself in the scan's tail and reports *exchanged = true.
int svdb_engine::nearest_local(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out, int mode,
                               svdb_exchange *x, svdb_candidate *d_merged, bool *exchanged) {
    const bool fp64_only = mode == SVDB_MODE_FP64;       // escalation step: wide fp64 scan + re-rank, no low-precision keys
    if (fp64_only) mode = SVDB_MODE_AUTO;
    if (k < 1 || k > SVDB_MAX_K) return fail(SVDB_ERR_ARG, "k must be in 1..SVDB_MAX_K");
    if (no_log) return fail(SVDB_ERR_ARG, "this engine was created without a log (SVDB_FLAG_NO_LOG)");
    if (nq == 0) return SVDB_OK;
    if (!d_Q || !d_out || ldq < (size_t)K) return fail(SVDB_ERR_ARG, "bad query buffer");
    int rc = flush();
    if (rc) return rc;
    CK(cudaSetDevice(device));
    if (mode == SVDB_MODE_TREE && !use_tree)
        return fail(SVDB_ERR_ARG, "tree traversal needs an engine that keeps the tree");
    if (mode == SVDB_MODE_MTREE && !mtree_wanted(k, mode))
        return fail(SVDB_ERR_ARG, "the median tree serves engines with thin kd-points (kd_dim <= 8)");
    // K9: balanced median tree; the queries it flags (distinct points tied at the minimum) go through K6
    if (!fp64_only && mtree_wanted(k, mode)) {
        rc = mtree_update();
        if (rc) return rc;
        std::string err;
        if (use_tree && !mt_marks.ensure(nq * 4, err)) return fail(SVDB_ERR_OOM, err);
        // a shard keeps no reference-shaped tree: its flag travels with the candidates and the tie is settled
        // across the shards (tie_protocol.cu); plain (distance, seq) order is what the merge wants anyway
        const bool shard = (cfg.flags & SVDB_FLAG_SHARD) != 0;
        // lanes per query (k = 1): a full warp answers one query fastest, narrower groups keep more queries in flight --
        // measured best (profiles/r01_mtree_K6_vs_K9_ab.jsonl): 32 up to a few thousand queries per call, 16 up to a few
        // hundred thousand, 8 beyond
        const int lanes = mtree_lanes ? mtree_lanes : (nq < 8192 ? 32 : (nq < 524288 ? 16 : 8));
        CK(launch_mtree_nearest(mt, kd_ptr(), kstride, K, n_versions, d_Q, (int)ldq, (int)nq, (int)k, log_idx.as<u64>(),
                                cfg.seq_base, (use_tree || shard) ? 1 : 0, lanes,
                                use_tree ? mt_marks.as<unsigned>() : nullptr, d_out, stream));
        stats.kernels_launched++;
        if (use_tree) {
            CK(launch_tree_nearest(kd_ptr(), kstride, K, child.as<uint32_t>(), n_versions, d_Q, (int)ldq, (int)nq, (int)k,
                                   log_idx.as<u64>(), cfg.seq_base, d_out, stream, mt_marks.as<unsigned>()));
            stats.kernels_launched++;
        }
        return SVDB_OK;
    }
    // K6: thin kd-points prune well, and the traversal IS the reference's algorithm
    if (mode == SVDB_MODE_TREE || (mode == SVDB_MODE_AUTO && !fp64_only && use_tree && !force_exact && K <= tree_max_k)) {
        CK(launch_tree_nearest(kd_ptr(), kstride, K, child.as<uint32_t>(), n_versions, d_Q, (int)ldq, (int)nq, (int)k,
                               log_idx.as<u64>(), cfg.seq_base, d_out, stream));
        stats.kernels_launched++;
        return SVDB_OK;
    }
    const bool use_exact = mode == SVDB_MODE_EXACT || force_exact || !wide;
    int cap = (int)std::min<size_t>(32, k + 8);
    // K10: larger batches go to the tcgen05 tensor cores (split-bf16 keys, same exact re-rank)
    int umma_from, mma_from;
    batch_thresholds(k, umma_from, mma_from);
    if (!use_exact && !fp64_only && mode == SVDB_MODE_AUTO && n_versions && umma_ok && umma_from > 0 && nq >= (size_t)umma_from &&
        K >= umma_min_k && n_versions < (1ull << 31)) {
        const int r = nearest_umma(d_Q, nq, ldq, k, d_out);
        if (r != -1000) return r;
    }
    // K2: a batch large enough to be compute-bound goes to the FP64 tensor cores
    if (!use_exact && !fp64_only && mode == SVDB_MODE_AUTO && n_versions && mma_from > 0 && nq >= (size_t)mma_from) {
        const int G = mma_group_size(nq);
        std::string err;
        const int ldp = kstride;
        size_t done = 0;
        while (done < nq) {
            const size_t nqp = std::min<size_t>(nq - done, (size_t)G * tune.num_sms);
            const int ngroups = (int)((nqp + G - 1) / G);
            const int nstreams = std::max(1, tune.num_sms / ngroups);
            const size_t nq_pad = (size_t)ngroups * G;
            if (!qpad.ensure(nq_pad * (size_t)ldp * 8, err) || !qnorm.ensure(nq_pad * 8, err) ||
                !lists.ensure(nq_pad * (size_t)nstreams * cap * sizeof(Cand), err))
                return fail(SVDB_ERR_OOM, err);
            CK(launch_prep_queries(d_Q + done * ldq, (int)ldq, K, (int)nqp, (int)nq_pad, qpad.as<double>(), ldp,
                                   qnorm.as<double>(), stream));
            MmaArgs ma{};
            ma.pts = kd_ptr();
            ma.n = n_versions;
            ma.K = K;
            ma.stride = kstride;
            ma.xnorm = xnorm.as<double>();
            ma.q = qpad.as<double>();
            ma.qnorm = qnorm.as<double>();
            ma.ldq = ldp;
            ma.nq = (int)nqp;
            ma.group = G;
            ma.ngroups = ngroups;
            ma.nstreams = nstreams;
            ma.cap = cap;
            ma.lists = lists.as<Cand>();
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (profile_scan) {
                if (scan_events_used == scan_events.size()) {
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a));
                    CK(cudaEventCreate(&b));
                    scan_events.emplace_back(a, b);
                }
                ev0 = scan_events[scan_events_used].first;
                ev1 = scan_events[scan_events_used].second;
                scan_events_used++;
                CK(cudaEventRecord(ev0, stream));
            }
            CK(launch_scan_mma(ma, stream));
            if (ev1) CK(cudaEventRecord(ev1, stream));
            FinalArgs fa{};
            fa.lists = lists.as<Cand>();
            fa.nlists = nstreams;
            fa.cap = cap;
            fa.nq = (int)nqp;
            fa.k = (int)k;
            fa.pts = kd_ptr();
            fa.K = K;
            fa.stride = kstride;
            fa.q = qpad.as<double>();
            fa.ldq = ldp;
            fa.log_index = log_idx.as<u64>();
            fa.seq_base = cfg.seq_base;
            fa.eps = 0.0;
            fa.eabs_coef = 8.0 * (double)(K + 8) * ldexp(1.0, -53);
            fa.qnorm = qnorm.as<double>();
            fa.xn_max_bits = xnmax.as<unsigned long long>();
            fa.child = use_tree ? child.as<uint32_t>() : nullptr;
            fa.mark_ties = (cfg.flags & SVDB_FLAG_SHARD) ? 1 : 0;
            fa.out = d_out + done * k;
            CK(launch_finalize(fa, stream));
            stats.kernels_launched += 3;
            done += nqp;
        }
        return SVDB_OK;
    }
    const int nlists = n_versions ? scan_num_lists(tune, !use_exact) : 0;
    std::string err;

    // Which copy of the log the scan reads (option scan.plane): 2 = the bf16 hi plane (K12, 2 bytes per coordinate),
    // 1 = hi + lo planes (K11, 4 bytes), 0 = the fp64 rows (K1, 8 bytes).  Same answers; see engine.h.
    const int Kp = umma_kpad(K);
    int plane = 0;
    if (!use_exact && !fp64_only && mode == SVDB_MODE_AUTO && scan_plane > 0 && n_versions && umma_ok && K >= umma_min_k &&
        n_versions < (1ull << 31)) {
        // never build or extend the shadow under stream capture (see nearest_host): fall back to the fp64 rows there
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cs);
        // The coarser the plane, the wider the re-rank window: K12's holds ~2.3 candidates per requested neighbour at
        // 10M x 768, K13's ~4; a window that overflows the tail's candidate slots costs an fp64 scan on top.  Beyond
        // plane_max_k the hi + lo planes (K11: a window of exactly k rows) are the cheaper choice.
        // (K11's kernel is not tuned for short rows -- 0.27 ms vs K1's 0.19 at 1M x 128 -- so those stay on the fp64 rows.)
        int want = scan_plane >= 2 && plane_scan_supports(Kp, 1) && k <= (size_t)plane_max_k ? 2 : (Kp >= 384 || scan_plane == 1 ? 1 : 0);
        // the byte plane's window is ~1.5 x K12's: at 10M x 768 it holds ~4 candidates per requested neighbour; the tail
        // re-ranks up to FIN_NC = 128 of them (one thread each)
        if (nq <= (size_t)byte_plane_max_queries(k)) want = 3;
        if (want == 3) {
            const bool have8 = plane8_ready && plane8_n == n_versions;
            if (cs == cudaStreamCaptureStatusNone || have8) {
                const int sr = ensure_plane8();
                if (sr == SVDB_OK) plane = 3;
                else if (sr != -1000) return sr;
            }
            if (plane != 3) want = 2;
        }
        const bool have = shadow_ready && shadow_n == n_versions && (want == 2 || shadow_lo_n == n_versions);
        if (plane != 3 && want != 0 && (cs == cudaStreamCaptureStatusNone || have)) {
            const int sr = ensure_shadow(want == 1);
            if (sr == SVDB_OK) plane = want;
            else if (sr != -1000) return sr;
        }
    }
    // the single-plane scans' tail selects its candidates from ALL the CTA lists at once (tail.cuh, selection path: lists
    // shorter than 16 entries that fit shared memory together); a CTA that holds more than `cap` of a query's window
    // fails the completeness proof and the query is re-answered from the fp64 rows
    if (plane >= 2) cap = std::min(cap, 15);
    if (!fp64_only) {                    // an escalation step must not hide which copy of the log the call itself scanned
        last_scan_plane = plane;
        stats.scan_plane_last = (uint64_t)plane;
    }

    const double *qbase = d_Q;
    int qld = (int)ldq;
    if (plane == 1) {
        // padded fp64 copies (the re-rank reads them) and |q|^2 per query (the key error bound scales with it)
        if (!qpad.ensure(nq * (size_t)kstride * 8, err) || !qnorm.ensure(nq * 8, err)) return fail(SVDB_ERR_OOM, err);
        CK(launch_prep_queries(d_Q, (int)ldq, K, (int)nq, (int)nq, qpad.as<double>(), kstride, qnorm.as<double>(), stream));
        stats.kernels_launched++;
        qbase = qpad.as<double>();
        qld = kstride;
    }
    // the wide scan stages whole query rows of kstride doubles with 16-byte bulk copies:
    // use the caller's buffer in place when it already has that shape, else pad a copy
    const bool q_in_place = K == kstride && (ldq % 2) == 0 && (reinterpret_cast<uintptr_t>(d_Q) % 16) == 0;
    if (!use_exact && plane == 0 && !q_in_place) {
        if (!qpad.ensure(nq * (size_t)kstride * 8, err)) return fail(SVDB_ERR_OOM, err);
        CK(launch_pad_queries(d_Q, (int)ldq, qpad.as<double>(), kstride, K, (int)nq, stream));
        stats.kernels_launched++;
        qbase = qpad.as<double>();
        qld = kstride;
    }
    int limit = std::max(1, tune.nq_per_pass);
    if (plane == 2) limit = plane_scan_supports(Kp, 2) ? std::min(limit, 2) : 1;
    if (plane == 3) limit = plane8_pair && plane8_scan_supports_two(Kp) ? 2 : 1;
    if (nlists && !lists.ensure((size_t)8 * nlists * cap * sizeof(Cand), err)) return fail(SVDB_ERR_OOM, err);
    // the scan's last CTA finalizes (and exchanges) itself -- every wide scan but the LDG variant and the exact kernel
    const bool fuse = fuse_tail && ticket.p && nlists && !use_exact && (plane > 0 || tune.variant == 0);

    // K13 with a fused tail, outside stream capture: consecutive launches overlap (engine.h: overlap_steps)
    bool overlap = false;
    if (plane >= 2 && fuse && overlap_steps && !in_host_call && nlists > 1 && dyn_tiles == 0) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cs);
        overlap = cs == cudaStreamCaptureStatusNone;
    }
    const int nl = overlap ? nlists - 1 : nlists;

    size_t done = 0;
    while (done < nq) {
        const int nqp = largest_pass(nq - done, limit);
        FinalArgs fa{};
        fa.lists = lists.as<Cand>();
        fa.nlists = nl;
        fa.cap = cap;
        fa.nq = nqp;
        fa.k = (int)k;
        fa.pts = kd_ptr();
        fa.K = K;
        fa.stride = kstride;
        fa.q = qbase + done * (size_t)qld;
        fa.ldq = qld;
        fa.log_index = log_idx.as<u64>();
        fa.seq_base = cfg.seq_base;
        fa.eps = use_exact ? -1.0 : 4.0 * (double)(K + 2) * ldexp(1.0, -53);
        if (plane == 1) {
            fa.eps = shadow_eps(K);
            fa.eabs_coef = shadow_eabs_coef();
            fa.qnorm = qnorm.as<double>() + done;
            fa.xn_max_bits = xnmax.as<unsigned long long>();
            fa.scale_lo = 1e-24;               // fp32 keys: see nearest_umma
            fa.scale_hi = 1e30;
        } else if (plane == 3) {
            fa.sq_mode = 2;
            fa.sq_gamma = ldexp(1.0, -50);     // the keys are exact integers times one rounded constant
            fa.plane8 = plane8_par.as<Plane8Par>();
            fa.plane_err_bits = reinterpret_cast<const unsigned long long *>(plane8_par.as<unsigned char>() + sizeof(Plane8Par));
            fa.xn_max_bits = xnmax.as<unsigned long long>();
        } else if (plane == 2) {
            fa.sq_mode = 1;
            fa.sq_gamma = plane_gamma(Kp);
            fa.plane_err_bits = plane_err.as<unsigned long long>();
            fa.xn_max_bits = xnmax.as<unsigned long long>();
            fa.scale_lo = 0.0;                 // underflow is inside the bound (tail.cuh); squares must not overflow fp32
            fa.scale_hi = 1e30;
        }
        fa.child = use_tree ? child.as<uint32_t>() : nullptr;
        fa.mark_ties = (cfg.flags & SVDB_FLAG_SHARD) ? 1 : 0;
        fa.out = d_out + done * k;
        TailArgs ta{};
        bool fused_exchange = false;
        if (fuse) {
            ta.ticket = ticket.as<unsigned>();
            ta.fin = fa;
            if (tail_debug) {
                if (!tail_dbg.ensure((size_t)(32 + 3 * nlists) * 8, err)) return fail(SVDB_ERR_OOM, err);
                ta.dbg = tail_dbg.as<unsigned long long>();
            }
            if (x && exchanged && done == 0 && (size_t)nqp == nq) {      // the whole call is this one pass
                exchange_fill_tail(x, ta, d_merged);
                fused_exchange = true;
            }
        }
        if (nlists) {
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (profile_scan) {
                if (scan_events_used == scan_events.size()) {
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a));
                    CK(cudaEventCreate(&b));
                    scan_events.emplace_back(a, b);
                }
                ev0 = scan_events[scan_events_used].first;
                ev1 = scan_events[scan_events_used].second;
                scan_events_used++;
                CK(cudaEventRecord(ev0, stream));
            }
            if (plane == 3) {
                Plane8ScanArgs pa{};
                pa.x8 = plane8.as<unsigned char>();
                pa.par = plane8_par.as<Plane8Par>();
                pa.n = n_versions;
                pa.K = K;
                pa.Kp = Kp;
                pa.q = fa.q;
                pa.ldq = qld;
                pa.nq = nqp;
                pa.cap = cap;
                pa.lists = lists.as<Cand>();
                pa.tail = ta;
                pa.grid = overlap ? nl : 0;
                // the scan in front must be one that waits for ITS predecessor before it writes (a chain of K13 launches)
                pa.pdl = overlap && !profile_scan && stats.kernels_launched == pdl_mark ? 1 : 0;
                CK(launch_scan_plane8(tune, pa, stream));
                if (overlap) pdl_mark = stats.kernels_launched + 1;
            } else if (plane == 2) {
                PlaneScanArgs pa{};
                pa.xhi = shadow_hi.as<uint16_t>();
                pa.n = n_versions;
                pa.K = K;
                pa.Kp = Kp;
                pa.q = fa.q;
                pa.ldq = qld;
                pa.nq = nqp;
                pa.cap = cap;
                pa.lists = lists.as<Cand>();
                pa.tail = ta;
                pa.dyn_eighths = dyn_tiles;
                pa.grid = overlap ? nl : 0;
                pa.pdl = overlap && !profile_scan && stats.kernels_launched == pdl_mark ? 1 : 0;
                CK(launch_scan_plane(tune, pa, stream));
                if (overlap) pdl_mark = stats.kernels_launched + 1;
            } else if (plane == 1) {
                ShadowScanArgs ha{};
                ha.xhi = shadow_hi.as<uint16_t>();
                ha.xlo = shadow_lo.as<uint16_t>();
                ha.n = n_versions;
                ha.K = K;
                ha.Kp = Kp;
                ha.q = fa.q;
                ha.ldq = qld;
                ha.nq = nqp;
                ha.cap = cap;
                ha.lists = lists.as<Cand>();
                ha.tail = ta;
                CK(launch_scan_shadow(tune, ha, stream));
            } else {
                ScanArgs sa{};
                sa.pts = kd_ptr();
                sa.n = n_versions;
                sa.K = K;
                sa.stride = kstride;
                sa.q = fa.q;
                sa.ldq = qld;
                sa.nq = nqp;
                sa.cap = cap;
                sa.lists = lists.as<Cand>();
                sa.assign = tune.assign;
                sa.tail = ta;
                CK(use_exact ? launch_scan_exact(tune, sa, stream) : launch_scan_wide(tune, sa, stream));
            }
            if (ev1) CK(cudaEventRecord(ev1, stream));
            stats.kernels_launched++;
        }
        if (!fuse) {
            CK(launch_finalize(fa, stream));
            stats.kernels_launched++;
        }
        if (fused_exchange) *exchanged = true;
        done += nqp;
    }
    return SVDB_OK;
}

// The split-bf16 shadow of the kd log (two planes of [versions][Kp] bf16; K10, K11 and K12 read it): created on first use, extended by
// the entries appended since.  -1000: no HBM for it (or no address space) -- the fp64 paths keep serving.
int svdb_engine::ensure_shadow(bool need_lo) {
    std::string err;
    const int Kp = umma_kpad(K);
    const size_t plane_row = (size_t)Kp * 2;
    if (!shadow_ready) {
        if (!plane_err.ensure(8, err)) return fail(SVDB_ERR_OOM, err);
        CK(cudaMemsetAsync(plane_err.p, 0, 8, stream));
        if (!shadow_hi.init(device, max_versions * plane_row, err) || !shadow_lo.init(device, max_versions * plane_row, err)) {
            umma_ok = false;
            return -1000;
        }
        shadow_ready = true;
    }
    bool grown = false;
    if (shadow_n < n_versions) {
        if (!shadow_hi.ensure(n_versions * plane_row, stream, err)) {
            umma_ok = false;                 // not enough HBM next to the store: nothing is lost
            cudaGetLastError();
            return -1000;
        }
        CK(launch_split_bf16(kd_ptr(), kstride, K, Kp, shadow_n, n_versions - shadow_n, shadow_hi.as<uint16_t>(), nullptr,
                             plane_err.as<unsigned long long>(), tune.num_sms, stream));
        stats.kernels_launched++;
        shadow_n = n_versions;
        grown = true;
    }
    if (need_lo && shadow_lo_n < n_versions) {
        if (!shadow_lo.ensure(n_versions * plane_row, stream, err)) {
            cudaGetLastError();
            return -1000;                    // K12 keeps its hi plane; K10 / K11 are not available (K2 / K12 serve)
        }
        CK(launch_split_bf16(kd_ptr(), kstride, K, Kp, shadow_lo_n, n_versions - shadow_lo_n, nullptr, shadow_lo.as<uint16_t>(),
                             nullptr, tune.num_sms, stream));
        stats.kernels_launched++;
        shadow_lo_n = n_versions;
        grown = true;
    }
    if (grown) {
        const size_t mapped = shadow_hi.mapped() + shadow_lo.mapped();
        stats.hbm_bytes_mapped += mapped - shadow_mapped_counted;
        shadow_mapped_counted = mapped;
    }
    return SVDB_OK;
}

// K13's one-byte plane: created (grid chosen from the rows present) on first use, extended by the entries appended since.
// -1000: no HBM for it -- K12 keeps serving.
bool svdb_engine::byte_plane_serves(size_t k) const {
    return scan_plane >= 3 && plane8_ok && (k <= 4 || plane8_ok_bigk) && k <= (size_t)plane8_max_k && wide && umma_ok && K >= umma_min_k &&
           plane8_scan_supports(umma_kpad(K));
}
// Up to how many queries per call K13 passes answer (0: K13 does not serve this call).  Measured (profiles/r02_sweep_*.jsonl,
// ms per top-10 call of <= 64 queries, e = rows x padded kd_dim elements):
//   K13 ~ nq (0.035 + 0.147e-9 e)       one launch per query, an eighth of the fp64 bytes each (a pair of queries: 1.5 x one)
//   K10 ~ 0.15 + 0.72e-9 e              (64-query groups; reads 4 bytes per element once)
//   K2  ~ 0.06 + 2.0e-9 e               (<= 16 queries; reads the 8-byte rows once), DMMA-bound beyond
// -> K13 passes win up to 4 queries per call from ~4e7 elements on (1M x 128: 0.18 vs 0.25 / 0.32 ms; 2M x 768: 1.08 vs 1.26 /
//    2.75; 20M x 128: 1.80 vs 2.00 / 5.4) and up to 2 below (250k x 128: 4 passes 0.15 ms, K2 0.12);
//    K10 beats K2 from ~7e7 elements on (1M x 128: 0.25 vs 0.32; 500k x 128: 0.20 vs 0.19; 250k x 128: 0.18 vs 0.12).
int svdb_engine::byte_plane_max_queries(size_t k) const {
    if (!byte_plane_serves(k)) return 0;
    const u64 e = n_versions * (u64)umma_kpad(K);
    if (plane8_max_q_user) return plane8_max_q;           // set through svdb_set_option: taken literally
    // two queries share a pass at 1.5 x the time of one (Kp >= 192): on large stores three pairs still beat one K10 call
    // (10M x 768: 6 queries 4.9 vs 5.6 ms; 2M x 768: 1.0 vs 1.2; 1M x 256: 0.33 vs 0.28 -- there it stays at 4)
    if (plane8_pair && plane8_scan_supports_two(umma_kpad(K)) && e >= (1ull << 29)) return plane8_max_q + 2;
    return e >= (40ull << 20) ? plane8_max_q : std::min(2, plane8_max_q);
}
void svdb_engine::batch_thresholds(size_t k, int &uq, int &mq) const {
    uq = umma_min_q >= 0 ? umma_min_q : ((K >= 256 || n_versions * (u64)umma_kpad(K) > (1ull << 26)) ? 3 : 32);
    mq = mma_min_q;
    const int p8 = byte_plane_max_queries(k);
    if (p8 > 0) {
        if (!umma_min_user && uq > 0) uq = std::max(uq, p8 + 1);
        if (!mma_min_user && mq > 0) mq = std::max(mq, p8 + 1);
    }
}

int svdb_engine::ensure_plane8() {
    std::string err;
    const int Kp = umma_kpad(K);
    if (!plane8_ready) {
        if (!plane8_par.ensure(sizeof(Plane8Par) + 16, err)) return fail(SVDB_ERR_OOM, err);
        CK(cudaMemsetAsync(plane8_par.p, 0, sizeof(Plane8Par) + 16, stream));
        if (!plane8.init(device, max_versions * (size_t)Kp, err)) {
            plane8_ok = false;
            return -1000;
        }
        plane8_ready = true;
    }
    if (plane8_n < n_versions) {
        if (!plane8.ensure(n_versions * (size_t)Kp, stream, err)) {
            plane8_ok = false;
            cudaGetLastError();
            return -1000;
        }
        unsigned long long *errw = reinterpret_cast<unsigned long long *>(plane8_par.as<unsigned char>() + sizeof(Plane8Par));
        pdl_mark = ~0ull;                      // the next scan reads what this writes: a fully ordered launch
        CK(launch_plane8_build(kd_ptr(), kstride, K, Kp, plane8_n, n_versions - plane8_n, plane8_par.as<Plane8Par>(), plane8_n == 0,
                               plane8.as<unsigned char>(), errw, tune.num_sms, stream));
        stats.kernels_launched += plane8_n == 0 ? 4 : 1;
        stats.hbm_bytes_mapped += plane8.mapped() - plane8_mapped_counted;
        plane8_mapped_counted = plane8.mapped();
        plane8_n = n_versions;
    }
    return SVDB_OK;
}

// K10 (umma_filter.cu).  Returns -1000 when the path cannot serve this engine (no HBM for the shadow, no tensor-map
// entry point): the caller falls through to K2.
int svdb_engine::nearest_umma(const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out) {
    std::string err;
    const int Kp = umma_kpad(K);
    const size_t row_bytes = (size_t)2 * Kp * 2;         // both planes of one query
    // never build or extend the shadow under stream capture (see nearest_host): K2 serves that call
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs != cudaStreamCaptureStatusNone && !(shadow_ready && shadow_n == n_versions && shadow_lo_n == n_versions)) return -1000;
    const int sr = ensure_shadow(true);
    if (sr) return sr;
    const int bn = umma_group_size(nq);
    const int cap = (int)std::min<size_t>(32, k + 14);
    const int ldp = kstride;
    size_t done = 0;
    while (done < nq) {
        const size_t nqp = std::min<size_t>(nq - done, (size_t)bn * tune.num_sms);
        const int ngroups = (int)((nqp + bn - 1) / bn);
        const int nstreams = std::max(1, tune.num_sms / ngroups);
        const size_t nq_pad = (size_t)ngroups * bn;
        if (!qpad.ensure(nq_pad * (size_t)ldp * 8, err) || !qnorm.ensure(nq_pad * 8, err) || !qsplit.ensure(nq_pad * row_bytes, err) ||
            !lists.ensure(nq_pad * (size_t)nstreams * cap * sizeof(Cand), err) || !ubuf.ensure(umma_buf_bytes(ngroups, nstreams, bn) + nq_pad * 4 * (size_t)(1 + nstreams), err))
            return fail(SVDB_ERR_OOM, err);
        if (umma_debug && !udbg.ensure((size_t)(128 * 256 + 16) * 4, err)) return fail(SVDB_ERR_OOM, err);
        CK(launch_prep_queries(d_Q + done * ldq, (int)ldq, K, (int)nqp, (int)nq_pad, qpad.as<double>(), ldp, qnorm.as<double>(), stream));
        uint16_t *qhi = qsplit.as<uint16_t>(), *qlo = qhi + nq_pad * (size_t)Kp;
        CK(launch_split_bf16(qpad.as<double>(), ldp, K, Kp, 0, nq_pad, qhi, qlo, nullptr, tune.num_sms, stream));
        UmmaArgs ua{};
        ua.xhi = shadow_hi.as<uint16_t>();
        ua.xlo = shadow_lo.as<uint16_t>();
        ua.n = n_versions;
        ua.K = K;
        ua.Kp = Kp;
        ua.xnorm = xnorm.as<double>();
        ua.qhi = qhi;
        ua.qlo = qlo;
        ua.qnorm = qnorm.as<double>();
        ua.nq = (int)nqp;
        ua.bn = bn;
        ua.ngroups = ngroups;
        ua.nstreams = nstreams;
        ua.cap = cap;
        ua.lists = lists.as<Cand>();
        ua.bufs = ubuf.p;
        ua.gtau = reinterpret_cast<uint32_t *>(static_cast<char *>(ubuf.p) + umma_buf_bytes(ngroups, nstreams, bn));
        ua.gmin = umma_group_min ? ua.gtau + nq_pad : nullptr;
        ua.sparse_checks = umma_sparse_checks && umma_group_min ? 1 : 0;
        CK(cudaMemsetAsync(ua.gtau, 0xff, nq_pad * 4 * (size_t)(1 + (umma_group_min ? nstreams : 0)), stream));
        ua.dbg_keys = (umma_debug && done == 0) ? udbg.as<float>() : nullptr;
        ua.qres = umma_resident ? umma_resident_stages(bn, Kp) : 0;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (profile_scan) {
            if (scan_events_used == scan_events.size()) {
                cudaEvent_t a, b;
                CK(cudaEventCreate(&a));
                CK(cudaEventCreate(&b));
                scan_events.emplace_back(a, b);
            }
            ev0 = scan_events[scan_events_used].first;
            ev1 = scan_events[scan_events_used].second;
            scan_events_used++;
            CK(cudaEventRecord(ev0, stream));
        }
        cudaError_t ce = launch_umma_filter(ua, stream, &err);
        if (ce == cudaErrorNotSupported && done == 0) {     // no cuTensorMapEncodeTiled in this driver
            umma_ok = false;
            if (profile_scan) scan_events_used--;
            return -1000;
        }
        CK(ce);
        if (ev1) CK(cudaEventRecord(ev1, stream));
        FinalArgs fa{};
        fa.lists = lists.as<Cand>();
        fa.nlists = nstreams;
        fa.cap = cap;
        fa.nq = (int)nqp;
        fa.k = (int)k;
        fa.pts = kd_ptr();
        fa.K = K;
        fa.stride = kstride;
        fa.q = qpad.as<double>();
        fa.ldq = ldp;
        fa.log_index = log_idx.as<u64>();
        fa.seq_base = cfg.seq_base;
        fa.eps = 0.0;
        fa.eabs_coef = umma_eabs_coef(K);
        fa.qnorm = qnorm.as<double>();
        fa.xn_max_bits = xnmax.as<unsigned long long>();
        // fp32 keys: products of coordinates below ~1e-19 underflow, squares above ~1e38 overflow; outside this range
        // of max|x|^2 + |q|^2 the error bound does not hold and the query is answered by the exact scan
        fa.scale_lo = 1e-24;
        fa.scale_hi = 1e30;
        fa.child = use_tree ? child.as<uint32_t>() : nullptr;
        fa.mark_ties = (cfg.flags & SVDB_FLAG_SHARD) ? 1 : 0;
        fa.out = d_out + done * k;
        CK(launch_finalize(fa, stream));
        stats.kernels_launched += 4;
        done += nqp;
    }
    umma_debug = false;
    return SVDB_OK;
}

int svdb_engine::nearest_host(const double *Q, size_t nq, size_t ldq, size_t k, size_t *index_out, double *dist_out,
                              uint64_t *seq_out, svdb_candidate *cand_out, svdb_exchange *x) {
    if (x && !exchange_fits(x, nq, k)) return fail(SVDB_ERR_ARG, "nq * k exceeds the exchange capacity");
    if (k < 1 || k > SVDB_MAX_K) return fail(SVDB_ERR_ARG, "k must be in 1..SVDB_MAX_K");
    if (nq == 0) return SVDB_OK;
    if (!Q || ldq < (size_t)K) return fail(SVDB_ERR_ARG, "bad query buffer");
    CK(cudaSetDevice(device));
    // A host call copies its queries in, waits for the answers and copies them out: nothing of it can overlap with the call
    // before or after it, so its launches are plain ones (the overlap of engine.h: overlap_steps is for the device entry
    // points, where consecutive calls queue up on the stream with the queries already resident).
    struct HostCall {
        bool &f;
        explicit HostCall(bool &flag) : f(flag) { f = true; }
        ~HostCall() { f = false; }
    } host_call_guard(in_host_call);
    std::string err;
    if (!hq.ensure(nq * (size_t)K * 8, err) || !qraw.ensure(nq * (size_t)K * 8, err) ||
        !outc.ensure((nq * k + 1) * sizeof(svdb_candidate), err) || !hout.ensure(nq * k * sizeof(svdb_candidate), err))
        return fail(SVDB_ERR_OOM, err);
    for (size_t i = 0; i < nq; i++) memcpy(hq.as<double>() + i * K, Q + i * ldq, (size_t)K * 8);
    int rc = flush();
    if (rc) return rc;
    if (mtree_wanted(k, SVDB_MODE_AUTO)) {       // before any capture: a rebuild allocates and synchronizes
        rc = mtree_update();
        if (rc) return rc;
    }
    int umma_from, mma_from;
    batch_thresholds(k, umma_from, mma_from);
    const bool to_umma = umma_from > 0 && nq >= (size_t)umma_from;     // K10 serves the call
    const bool few = !to_umma && (mma_from <= 0 || nq < (size_t)mma_from);   // the single-query scans (K13 / K12 / K11) do
    if (((scan_plane > 0 && few) || to_umma) && wide && !force_exact && umma_ok && K >= umma_min_k &&
        n_versions && n_versions < (1ull << 31)) {
        // the shadow K10 / K11 read likewise: building it inside a capture that is later discarded would leave shadow_n
        // ahead of what was actually converted
        const bool byte_plane = few && nq <= (size_t)byte_plane_max_queries(k);
        if (byte_plane) {
            rc = ensure_plane8();
            if (rc && rc != -1000) return rc;
        }
        if (!(byte_plane && plane8_ready && plane8_n == n_versions) || to_umma) {
            rc = ensure_shadow(to_umma || scan_plane == 1 || !plane_scan_supports(umma_kpad(K), 1) || k > (size_t)plane_max_k);
            if (rc && rc != -1000) return rc;
        }
    }
    stats.h2d_bytes += nq * (size_t)K * 8;
    stats.d2h_bytes += nq * k * sizeof(svdb_candidate);

    // the enqueue sequence of one call: H2D of the queries, the kernels, D2H of the candidates
    // Small transfers are latency, not bandwidth: the kernels write the candidates straight into the
    // pinned host block (posted PCIe writes, no separate D2H op), and a thin query batch (a few hundred
    // bytes, read once per CTA) is read straight from pinned memory instead of being copied first.
    const bool q_zero_copy = !wide && nq * (size_t)K * 8 <= 2048;
    const double *d_q = q_zero_copy ? hq.as<double>() : qraw.as<double>();
    auto enqueue_mode = [&](int mode) -> int {
        // sharded (x): this shard's scan, the peer-memory exchange and the merge; the merged answers land in hout
        return nearest_device(d_q, nq, K, k, hout.as<svdb_candidate>(), mode, x);
    };
    auto enqueue = [&]() -> int {
        if (!q_zero_copy) CK(cudaMemcpyAsync(qraw.p, hq.p, nq * (size_t)K * 8, cudaMemcpyHostToDevice, stream));
        return enqueue_mode(SVDB_MODE_AUTO);
    };
    const unsigned long long gen = scratch_generation() + opt_gen;
    bool done = false;
    if (graphs_enabled && !profile_scan && nq <= 64 && (n_versions > 0 || x)) {
        for (auto &g : graphs) {
            if (g.nq == nq && g.k == k && g.x == x && g.n_versions == n_versions && g.gen == gen && g.stream == stream) {
                CK(cudaGraphLaunch(g.exec, stream));
                stats.kernels_launched += g.launches;
                done = true;
                break;
            }
        }
        if (!done && last_nq == nq && last_k == k) {
            // second call of this shape in a row (buffers are warm, nothing will allocate): capture it
            if (graphs.size() >= 8 || (!graphs.empty() && (graphs[0].n_versions != n_versions || graphs[0].gen != gen))) drop_graphs();
            const uint64_t l0 = stats.kernels_launched;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            cudaError_t ce = cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed);
            int r = ce == cudaSuccess ? enqueue() : SVDB_ERR_CUDA;
            cudaError_t ce2 = ce == cudaSuccess ? cudaStreamEndCapture(stream, &graph) : ce;
            if (r == SVDB_OK && ce2 == cudaSuccess && graph && scratch_generation() + opt_gen == gen &&
                cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                graphs.push_back(HostGraph{nq, k, n_versions, x, gen, stream, exec, stats.kernels_launched - l0});
                CK(cudaGraphLaunch(exec, stream));
                done = true;
            } else {
                cudaGetLastError();
                // a scratch block grew during the capture (its pointers are baked in): just try again next time; anything
                // else means something in the sequence is not capturable here -- plain launches from now on
                if (r != SVDB_OK || ce2 != cudaSuccess || !graph || scratch_generation() + opt_gen == gen) graphs_enabled = false;
                stats.kernels_launched = l0;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        last_nq = nq;
        last_k = k;
    }
    if (!done) {
        rc = enqueue();
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(stream));
    svdb_candidate *res = hout.as<svdb_candidate>();
    // Escalation for queries whose answer could not be proven complete:
    //   low-precision keys (K10 / K11 / K12 / K2) -> FP64 (K1: fp64 rows, relative error ~K 2^-53) -> EXACT (mass
    //   near-ties defeated every approximate candidate set, or the traversal stack overflowed) -> TREE (more
    //   exactly-tied entries than a candidate list holds; k = 1).
    if (x) {
        // merged flags are identical on every rank, so every rank takes (or skips) these branches together
        auto any_flag = [&](uint64_t f) {
            for (size_t i = 0; i < nq; i++)
                if (res[i * k].flags & f) return true;
            return false;
        };
        if (any_flag(SVDB_CAND_UNSAFE) && wide) {
            stats.fp64_reruns += nq;
            rc = enqueue_mode(SVDB_MODE_FP64);
            if (rc) return rc;
            CK(cudaStreamSynchronize(stream));
        }
        if (any_flag(SVDB_CAND_UNSAFE)) {
            stats.exact_reruns += nq;
            rc = enqueue_mode(SVDB_MODE_EXACT);
            if (rc) return rc;
            CK(cudaStreamSynchronize(stream));
        }
        // distinct kd-points at exactly the minimal distance: which one does the reference's (global) tree reach
        // first?  Decided by all shards together; again every rank sees the same flags.
        if (any_flag(SVDB_CAND_TIE)) {
            rc = resolve_ties_engine(this, x, exchange_rank(x), exchange_world(x), nullptr, nullptr, hq.as<double>(), nq,
                                     (size_t)K, res, k);
            if (rc) return rc;
        }
    }
    if (last_scan_plane == 3) {
        // a store-wide uniform grid resolves some data badly (heavy tails, a few huge coordinates): the measured plane error
        // then makes most proofs fail and every query pays for two scans.  Stop using the plane when that shows.
        const int cls = k > 4 ? 1 : 0;
        p8_calls[cls] += nq;
        for (size_t i = 0; i < nq; i++) p8_unsafe[cls] += (res[i * k].flags & SVDB_CAND_UNSAFE) ? 1 : 0;
        if (p8_calls[cls] >= 8 && p8_unsafe[cls] * 4 > p8_calls[cls]) {
            (cls ? plane8_ok_bigk : plane8_ok) = false;
            opt_gen++;                       // captured graphs baked the plane's scan in
        }
    }
    const bool low_precision_first = wide && (last_scan_plane > 0 || nq >= (size_t)std::max(1, mma_from));
    for (size_t i = 0; i < nq && !x; i++) {
        svdb_candidate *r = res + i * k;
        if (!(r[0].flags & SVDB_CAND_UNSAFE)) continue;
        if (low_precision_first) {
            stats.fp64_reruns++;
            rc = nearest_device(d_q + i * K, 1, K, k, r, SVDB_MODE_FP64);
            if (rc) return rc;
            CK(cudaStreamSynchronize(stream));
            if (!(r[0].flags & SVDB_CAND_UNSAFE)) continue;
        }
        stats.exact_reruns++;
        rc = nearest_device(d_q + i * K, 1, K, k, r, SVDB_MODE_EXACT);
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        if (!(r[0].flags & SVDB_CAND_UNSAFE) || !use_tree) continue;
        // more exactly-tied entries than a list holds: ask the tree which one the reference reaches first
        stats.tree_reruns++;
        if (!outc.ensure((nq * k + 1) * sizeof(svdb_candidate), err)) return fail(SVDB_ERR_OOM, err);
        svdb_candidate *d_one = outc.as<svdb_candidate>() + nq * k;
        svdb_candidate w;
        rc = nearest_device(d_q + i * K, 1, K, 1, d_one, SVDB_MODE_TREE);
        if (rc) return rc;
        CK(cudaMemcpyAsync(&w, d_one, sizeof w, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (w.flags & SVDB_CAND_UNSAFE) continue;      // traversal stack overflow (degenerate tree): keep (d, seq) order
        // winner first, the rest stay in (dist, seq) order
        size_t pos = k - 1;
        for (size_t j = 0; j < k; j++)
            if (r[j].seq == w.seq) { pos = j; break; }
        for (size_t j = pos; j > 0; j--) r[j] = r[j - 1];
        r[0] = w;
        for (size_t j = 0; j < k; j++) r[j].flags &= ~SVDB_CAND_UNSAFE;
    }
    for (size_t i = 0; i < nq * k; i++) {
        if (index_out) index_out[i] = (size_t)res[i].index;
        if (dist_out) dist_out[i] = res[i].dist;
        if (seq_out) seq_out[i] = res[i].seq;
    }
    if (cand_out) memcpy(cand_out, res, nq * k * sizeof(svdb_candidate));
    return SVDB_OK;
}

// Concurrent single-query callers: whoever finds no pass in flight becomes the leader and
// answers everything queued so far in ONE pass (queries of the same k share it); callers that
// arrive meanwhile queue up for the next leader.  Nobody waits for a timer.
int svdb_engine::nearest_one_coalesced(const double *q, size_t k, svdb_candidate *res) {
    PendingQuery me;
    me.q = q;
    me.k = k;
    me.res = res;
    std::unique_lock<std::mutex> lk(bq_mu);
    bq.push_back(&me);
    while (!me.done) {
        if (bq_leader) {
            bq_cv.wait(lk);
            continue;
        }
        bq_leader = true;
        std::vector<PendingQuery *> batch;
        for (auto it = bq.begin(); it != bq.end();) {
            if ((*it)->k == k) {
                batch.push_back(*it);
                it = bq.erase(it);
            } else {
                ++it;
            }
        }
        lk.unlock();
        const size_t nb = batch.size();
        std::vector<double> Q(nb * (size_t)K);
        for (size_t i = 0; i < nb; i++) memcpy(Q.data() + i * K, batch[i]->q, (size_t)K * 8);
        std::vector<svdb_candidate> out(nb * k);
        int rc;
        {
            std::lock_guard<std::mutex> g(mu);
            rc = nearest_host(Q.data(), nb, K, k, nullptr, nullptr, nullptr, out.data());
            if (nb > 1) {
                stats.coalesced_passes++;
                stats.coalesced_calls += nb - 1;
            }
        }
        const std::string err = rc ? get_last_error() : std::string();
        lk.lock();
        for (size_t i = 0; i < nb; i++) {
            if (!rc) memcpy(batch[i]->res, out.data() + i * k, k * sizeof(svdb_candidate));
            batch[i]->rc = rc;
            batch[i]->done = true;
        }
        bq_leader = false;
        bq_cv.notify_all();
        if (rc) set_last_error(err);
    }
    return me.rc;
}

int svdb_engine::compare_device(int mode, const uint64_t *d_i1, const uint64_t *d_i2, size_t n, float *d_out) {
    if (mode < 0 || mode > 3) return fail(SVDB_ERR_ARG, "unknown metric");
    if (log_only) return fail(SVDB_ERR_ARG, "a log-only engine stores no rows to compare");
    if (n == 0) return SVDB_OK;
    int rc = flush();
    if (rc) return rc;
    rc = upload_cur();
    if (rc) return rc;
    CK(cudaSetDevice(device));
    if (cur_host.empty()) {   // every pair is out of range: the reference's -1.0f sentinel
        const size_t total = n * (mode == 3 ? 3 : 1);
        fill_f32_kernel<<<(unsigned)std::min<size_t>(1024, (total + 255) / 256), 256, 0, stream>>>(d_out, -1.0f, total);
        CK(cudaGetLastError());
        stats.kernels_launched++;
        return SVDB_OK;
    }
    CompareArgs ca{};
    ca.rows = rows.as<double>();
    ca.ldr = Dpad;
    ca.D = D;
    ca.cur = cur.as<u64>();
    ca.nrows = cur_host.size();
    ca.i1 = reinterpret_cast<const u64 *>(d_i1);
    ca.i2 = reinterpret_cast<const u64 *>(d_i2);
    ca.n = n;
    ca.norm = norms.as<float>();
    ca.out = d_out;
    ca.mode = mode;
    CK(launch_compare(ca, tune.num_sms, stream));
    stats.kernels_launched++;
    return SVDB_OK;
}

int svdb_engine::compare_host(int mode, const size_t *i1, const size_t *i2, size_t n, float *out) {
    if (n == 0) return SVDB_OK;
    if (!i1 || !i2 || !out) return fail(SVDB_ERR_ARG, "NULL buffer");
    CK(cudaSetDevice(device));
    const size_t per = mode == 3 ? 3 : 1;
    std::string err;
    if (!hidx.ensure(2 * n * 8, err) || !idx1.ensure(n * 8, err) || !idx2.ensure(n * 8, err) ||
        !fout.ensure(n * per * 4, err) || !hf.ensure(n * per * 4, err))
        return fail(SVDB_ERR_OOM, err);
    memcpy(hidx.as<uint64_t>(), i1, n * 8);
    memcpy(hidx.as<uint64_t>() + n, i2, n * 8);
    CK(cudaMemcpyAsync(idx1.p, hidx.as<uint64_t>(), n * 8, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(idx2.p, hidx.as<uint64_t>() + n, n * 8, cudaMemcpyHostToDevice, stream));
    stats.h2d_bytes += 2 * n * 8;
    int rc = compare_device(mode, idx1.as<uint64_t>(), idx2.as<uint64_t>(), n, fout.as<float>());
    if (rc) return rc;
    CK(cudaMemcpyAsync(hf.p, fout.p, n * per * 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    stats.d2h_bytes += n * per * 4;
    memcpy(out, hf.p, n * per * 4);
    return SVDB_OK;
}

int svdb_engine::ingest_device_rows(const double *d_rows, size_t n, size_t ld, size_t *first_index) {
    svdb_engine *e = this;
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "log-only engine: use svdb_append_kdpoints");
    if (ld < (size_t)e->D) return e->fail(SVDB_ERR_ARG, "ld < dimension");
    int rc = e->flush();
    if (rc) return rc;
    if (first_index) *first_index = e->cur_host.size();
    if (n == 0) return SVDB_OK;
    if (cudaSetDevice(e->device) != cudaSuccess) return e->fail_cuda("cudaSetDevice", cudaGetLastError());
    const size_t n0 = e->n_versions, n1 = n0 + n;
    if (n1 > e->max_versions) return e->fail(SVDB_ERR_OOM, "store exceeds the reserved address range");
    std::string err;
    bool ok = (e->no_log || e->log_idx.ensure(n1 * 8, e->stream, err)) &&
              e->rows.ensure(n1 * (size_t)e->Dpad * 8, e->stream, err) && e->norms.ensure(n1 * 4, e->stream, err);
    if (!e->alias && !e->no_log) ok = ok && e->kdpts.ensure(n1 * (size_t)e->kstride * 8, e->stream, err);
    if (!ok) return e->fail(SVDB_ERR_OOM, err);
    double *dst = e->rows.as<double>() + n0 * (size_t)e->Dpad;
    cudaError_t ce;
    if (e->Dpad != e->D) {
        ce = cudaMemsetAsync(dst, 0, n * (size_t)e->Dpad * 8, e->stream);
        if (ce != cudaSuccess) return e->fail_cuda("cudaMemsetAsync", ce);
    }
    ce = cudaMemcpy2DAsync(dst, (size_t)e->Dpad * 8, d_rows, ld * 8, (size_t)e->D * 8, n, cudaMemcpyDeviceToDevice, e->stream);
    if (ce != cudaSuccess) return e->fail_cuda("cudaMemcpy2DAsync", ce);
    const size_t base_index = e->cur_host.size();
    if (!e->no_log) {
        ce = launch_iota(e->log_idx.as<u64>() + n0, base_index + e->index_base, n, e->stream);
        if (ce != cudaSuccess) return e->fail_cuda("iota", ce);
        e->stats.kernels_launched++;
    }
    if (!e->alias && !e->no_log) {
        ce = launch_extract_prefix(dst, e->Dpad, e->kdpts.as<double>() + n0 * (size_t)e->kstride, e->kstride, e->K, n, e->stream);
        if (ce != cudaSuccess) return e->fail_cuda("extract_prefix", ce);
        e->stats.kernels_launched++;
    }
    CompareArgs ca{};
    ca.rows = e->rows.as<double>();
    ca.ldr = e->Dpad;
    ca.D = e->D;
    ca.first = n0;
    ca.n = n;
    ca.out = e->norms.as<float>() + n0;
    ca.mode = 4;
    ce = launch_compare(ca, e->tune.num_sms, e->stream);
    if (ce != cudaSuccess) return e->fail_cuda("norm precompute", ce);
    e->stats.kernels_launched++;
    rc = e->tree_append(n0, n);
    if (rc) return rc;
    e->cur_host.reserve(base_index + n);
    for (size_t i = 0; i < n; i++) e->cur_host.push_back(n0 + i);
    e->uuids.resize(e->cur_host.size(), std::array<char, 37>{});
    e->n_versions = n1;
    e->stats.hbm_bytes_mapped = e->rows.mapped() + e->kdpts.mapped() + e->log_idx.mapped() + e->norms.mapped() +
                                e->cur.mapped() + e->child.mapped() + e->xnorm.mapped() + e->shadow_hi.mapped() + e->shadow_lo.mapped() + e->plane8.mapped();
    e->shadow_mapped_counted = e->shadow_hi.mapped() + e->shadow_lo.mapped();
    e->plane8_mapped_counted = e->plane8.mapped();
    return SVDB_OK;
}


// =====================================================================================
// C-ABI
// =====================================================================================
extern "C" {

const char *svdb_last_error(void) { return get_last_error().c_str(); }
const char *svdb_version(void) { return "svdb_b200 0.1 (sm_100a)"; }

int svdb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int svdb_engine_create(const svdb_config *cfg, svdb_engine **out) {
    if (!cfg || !out) {
        set_last_error("NULL argument");
        return SVDB_ERR_ARG;
    }
    *out = nullptr;
    svdb_engine *e = new (std::nothrow) svdb_engine();
    if (!e) return SVDB_ERR_OOM;
    int rc = e->init(*cfg);
    if (rc) {
        const std::string keep = get_last_error();
        e->destroy();
        delete e;
        set_last_error(keep);
        return rc;
    }
    *out = e;
    return SVDB_OK;
}

void svdb_engine_destroy(svdb_engine *e) {
    if (!e) return;
    e->destroy();
    delete e;
}

int svdb_set_stream(svdb_engine *e, void *stream) {
    if (!e) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    e->stream = stream == SVDB_STREAM_OWN ? e->own_stream : (cudaStream_t)stream;
    e->opt_gen++;
    return SVDB_OK;
}

int svdb_insert_batch(svdb_engine *e, const double *rows, size_t n, size_t ld, size_t *first_index) {
    if (!e || (!rows && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "log-only engine: use svdb_append_kdpoints");
    if (ld < (size_t)e->D) return e->fail(SVDB_ERR_ARG, "ld < dimension");
    if (first_index) *first_index = e->cur_host.size();
    for (size_t i = 0; i < n; i++) {
        // vector_database.c:113-115: the entry carries index = size before the increment
        const uint64_t ver = e->n_versions + e->stage_n;
        int rc = e->stage_one(rows + i * ld, e->D, e->cur_host.size() + e->index_base);
        if (rc) return rc;
        e->cur_host.push_back(ver);
        e->uuids.emplace_back();
    }
    return SVDB_OK;
}

int svdb_update_batch(svdb_engine *e, const size_t *index, const double *rows, size_t n, size_t ld) {
    if (!e || ((!rows || !index) && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "log-only engine has no rows to update");
    if (ld < (size_t)e->D) return e->fail(SVDB_ERR_ARG, "ld < dimension");
    for (size_t i = 0; i < n; i++) {
        if (index[i] >= e->cur_host.size()) continue;   // vector_database.c:171: silent no-op
        const uint64_t ver = e->n_versions + e->stage_n;
        int rc = e->stage_one(rows + i * ld, e->D, index[i] + e->index_base);   // :174 re-append with the same index
        if (rc) return rc;
        e->cur_host[index[i]] = ver;
        e->cur_dirty_lo = std::min(e->cur_dirty_lo, index[i]);
    }
    return SVDB_OK;
}

int svdb_delete_batch(svdb_engine *e, const size_t *index, size_t n) {
    if (!e || (!index && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "log-only engine has no rows to delete");
    for (size_t i = 0; i < n; i++) {
        if (index[i] >= e->cur_host.size()) continue;   // vector_database.c:187: silent no-op
        e->cur_host.erase(e->cur_host.begin() + index[i]);   // :189-191 shift; the log keeps its entries
        if (index[i] < e->uuids.size()) e->uuids.erase(e->uuids.begin() + index[i]);
        e->cur_dirty_lo = std::min(e->cur_dirty_lo, index[i]);
    }
    return SVDB_OK;
}

int svdb_append_kdpoints(svdb_engine *e, const double *pts, const size_t *index, size_t n, size_t ld) {
    if (!e || ((!pts || !index) && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (ld < (size_t)e->K) return e->fail(SVDB_ERR_ARG, "ld < kd_dim");
    if (e->no_log) return e->fail(SVDB_ERR_ARG, "this engine was created without a log");
    for (size_t i = 0; i < n; i++) {
        int rc = e->stage_one(pts + i * ld, e->K, index[i]);   // kdtree.c:20-28 copies K coordinates
        if (rc) return rc;
    }
    return SVDB_OK;
}

// kd-points already in device memory -> a bare log (SVDB_FLAG_LOG_ONLY): entry i reports index first_index + i
int svdb_append_kdpoints_device(svdb_engine *e, const double *d_pts, size_t first_index, size_t n, size_t ld) {
    if (!e || (!d_pts && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->log_only) return e->fail(SVDB_ERR_ARG, "svdb_append_kdpoints_device needs a SVDB_FLAG_LOG_ONLY engine");
    if (ld < (size_t)e->K) return e->fail(SVDB_ERR_ARG, "ld < kd_dim");
    int rc = e->flush();
    if (rc) return rc;
    if (n == 0) return SVDB_OK;
    if (cudaSetDevice(e->device) != cudaSuccess) return e->fail_cuda("cudaSetDevice", cudaGetLastError());
    const size_t n0 = e->n_versions, n1 = n0 + n;
    if (n1 > e->max_versions) return e->fail(SVDB_ERR_OOM, "store exceeds the reserved address range");
    std::string err;
    if (!e->log_idx.ensure(n1 * 8, e->stream, err) || !e->kdpts.ensure(n1 * (size_t)e->kstride * 8, e->stream, err))
        return e->fail(SVDB_ERR_OOM, err);
    double *dst = e->kdpts.as<double>() + n0 * (size_t)e->kstride;
    cudaError_t ce = cudaSuccess;
    if (e->kstride != e->K) ce = cudaMemsetAsync(dst, 0, n * (size_t)e->kstride * 8, e->stream);
    if (ce == cudaSuccess)
        ce = cudaMemcpy2DAsync(dst, (size_t)e->kstride * 8, d_pts, ld * 8, (size_t)e->K * 8, n, cudaMemcpyDeviceToDevice, e->stream);
    if (ce == cudaSuccess) ce = launch_iota(e->log_idx.as<u64>() + n0, first_index, n, e->stream);
    if (ce != cudaSuccess) return e->fail_cuda("svdb_append_kdpoints_device", ce);
    e->stats.kernels_launched++;
    rc = e->tree_append(n0, n);
    if (rc) return rc;
    e->n_versions = n1;
    e->stats.hbm_bytes_mapped = e->rows.mapped() + e->kdpts.mapped() + e->log_idx.mapped() + e->norms.mapped() +
                                e->cur.mapped() + e->child.mapped() + e->xnorm.mapped() + e->shadow_hi.mapped() + e->shadow_lo.mapped() + e->plane8.mapped();
    e->shadow_mapped_counted = e->shadow_hi.mapped() + e->shadow_lo.mapped();
    e->plane8_mapped_counted = e->plane8.mapped();
    return SVDB_OK;
}

int svdb_insert_batch_device(svdb_engine *e, const double *d_rows, size_t n, size_t ld, size_t *first_index) {
    if (!e || (!d_rows && n)) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    return e->ingest_device_rows(d_rows, n, ld, first_index);
}

int svdb_flush(svdb_engine *e) {
    if (!e) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    int rc = e->flush();
    if (rc) return rc;
    if (!e->log_only) rc = e->upload_cur();
    return rc;
}

size_t svdb_size(const svdb_engine *e) { return e ? e->cur_host.size() : 0; }
size_t svdb_log_size(const svdb_engine *e) { return e ? e->n_versions + e->stage_n : 0; }
size_t svdb_dimension(const svdb_engine *e) { return e ? (size_t)e->D : 0; }
size_t svdb_kd_dim(const svdb_engine *e) { return e ? (size_t)e->K : 0; }

int svdb_read_row(svdb_engine *e, size_t index, double *out) {
    if (!e || !out) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "log-only engine stores no rows");
    if (index >= e->cur_host.size()) return e->fail(SVDB_ERR_RANGE, "index out of range");
    int rc = e->flush();
    if (rc) return rc;
    cudaSetDevice(e->device);
    cudaError_t ce = cudaMemcpyAsync(out, e->rows.as<double>() + e->cur_host[index] * (size_t)e->Dpad, (size_t)e->D * 8,
                                     cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) return e->fail_cuda("read_row", ce);
    e->stats.d2h_bytes += (size_t)e->D * 8;
    return SVDB_OK;
}

int svdb_nearest_batch(svdb_engine *e, const double *Q, size_t nq, size_t ldq, size_t k, size_t *index_out,
                       double *dist_out, uint64_t *seq_out) {
    if (!e) return SVDB_ERR_ARG;
    if (nq == 1 && Q && k >= 1 && k <= SVDB_MAX_K && ldq >= (size_t)e->K) {
        svdb_candidate res[SVDB_MAX_K];
        int rc = e->nearest_one_coalesced(Q, k, res);
        if (rc) return rc;
        for (size_t i = 0; i < k; i++) {
            if (index_out) index_out[i] = (size_t)res[i].index;
            if (dist_out) dist_out[i] = res[i].dist;
            if (seq_out) seq_out[i] = res[i].seq;
        }
        return SVDB_OK;
    }
    std::lock_guard<std::mutex> g(e->mu);
    return e->nearest_host(Q, nq, ldq, k, index_out, dist_out, seq_out);
}

/* One shard's part of a sharded query, host buffers in and out: local scan, peer-memory exchange, merge.
 * Collective -- every rank calls it with the same nq and k.  Outputs are the MERGED answers. */
int svdb_nearest_batch_sharded(svdb_engine *e, svdb_exchange *x, const double *Q, size_t nq, size_t ldq, size_t k,
                               size_t *index_out, double *dist_out, uint64_t *seq_out) {
    if (!e || !x) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    return e->nearest_host(Q, nq, ldq, k, index_out, dist_out, seq_out, nullptr, x);
}

int svdb_nearest_batch_device(svdb_engine *e, const double *d_Q, size_t nq, size_t ldq, size_t k, svdb_candidate *d_out,
                              int mode) {
    if (!e) return SVDB_ERR_ARG;
    if (mode < SVDB_MODE_AUTO || mode > SVDB_MODE_FP64) return e->fail(SVDB_ERR_ARG, "unknown mode");
    std::lock_guard<std::mutex> g(e->mu);
    return e->nearest_device(d_Q, nq, ldq, k, d_out, mode);
}

/* The sharded query with device buffers, asynchronous on the engine's stream: this shard's scan, the peer-memory
 * exchange and the merge; d_out receives the MERGED nq x k candidates.  Collective (same nq, k on every rank). */
int svdb_nearest_batch_device_sharded(svdb_engine *e, svdb_exchange *x, const double *d_Q, size_t nq, size_t ldq, size_t k,
                                      svdb_candidate *d_out, int mode) {
    if (!e || !x) return SVDB_ERR_ARG;
    if (mode < SVDB_MODE_AUTO || mode > SVDB_MODE_FP64) return e->fail(SVDB_ERR_ARG, "unknown mode");
    std::lock_guard<std::mutex> g(e->mu);
    return e->nearest_device(d_Q, nq, ldq, k, d_out, mode, x);
}

int svdb_merge_candidates_device(int device, void *stream, const svdb_candidate *d_in, size_t nshards, size_t nq, size_t k,
                                 svdb_candidate *d_out) {
    if (!d_in || !d_out || k < 1 || k > SVDB_MAX_K || nshards < 1) {
        set_last_error("bad argument to svdb_merge_candidates_device");
        return SVDB_ERR_ARG;
    }
    if (nq == 0) return SVDB_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = launch_merge_candidates(d_in, (int)nshards, (int)nq, (int)k, d_out, (cudaStream_t)stream);
    if (ce != cudaSuccess) {
        set_last_error(std::string("merge_candidates: ") + cudaGetErrorString(ce));
        return SVDB_ERR_CUDA;
    }
    return SVDB_OK;
}

int svdb_compare_batch(svdb_engine *e, int metric, const size_t *index1, const size_t *index2, size_t n, float *out) {
    if (!e) return SVDB_ERR_ARG;
    if (metric < 0 || metric > 2) return e->fail(SVDB_ERR_ARG, "unknown metric");
    std::lock_guard<std::mutex> g(e->mu);
    return e->compare_host(metric, index1, index2, n, out);
}

int svdb_compare_batch_all(svdb_engine *e, const size_t *index1, const size_t *index2, size_t n, float *out) {
    if (!e) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    return e->compare_host(3, index1, index2, n, out);
}

int svdb_compare_batch_device(svdb_engine *e, int metric, const uint64_t *d_index1, const uint64_t *d_index2, size_t n,
                              float *d_out) {
    if (!e) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    return e->compare_device(metric, d_index1, d_index2, n, d_out);
}

// Two loose host vectors: upload both (zero padded to 128-byte rows), K4 for the norms, K3.
int svdb_compare_vectors(int device, int metric, const double *a, const double *b, size_t D, float *out) {
    static std::mutex mu;
    static Scratch dev[64];
    static PinnedScratch host[64];
    if (!a || !b || !out || D < 1 || metric < 0 || metric > 2 || device < 0 || device >= 64) {
        set_last_error("bad argument to svdb_compare_vectors");
        return SVDB_ERR_ARG;
    }
    std::lock_guard<std::mutex> g(mu);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count) {
        cudaGetLastError();
        set_last_error("no CUDA device (this library has no CPU path)");
        return SVDB_ERR_CUDA;
    }
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) {
        set_last_error(std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
        return SVDB_ERR_CUDA;
    }
    const size_t Dpad = round_up(D, 16);
    const size_t row_bytes = Dpad * 8;
    // device block: [row a][row b][idx 0,1 as u64 x2][norms f32 x2 + pad][result f32]
    const size_t bytes = 2 * row_bytes + 16 + 16 + 16;
    std::string err;
    if (!dev[device].ensure(bytes, err) || !host[device].ensure(bytes, err)) {
        set_last_error(err);
        return SVDB_ERR_OOM;
    }
    unsigned char *h = host[device].as<unsigned char>();
    memset(h, 0, bytes);
    memcpy(h, a, D * 8);
    memcpy(h + row_bytes, b, D * 8);
    uint64_t idx[2] = {0, 1};
    memcpy(h + 2 * row_bytes, idx, 16);
    unsigned char *d = dev[device].as<unsigned char>();
    cudaStream_t st = cudaStreamPerThread;
    ce = cudaMemcpyAsync(d, h, 2 * row_bytes + 16, cudaMemcpyHostToDevice, st);
    CompareArgs ca{};
    ca.rows = reinterpret_cast<const double *>(d);
    ca.ldr = (int)Dpad;
    ca.D = (int)D;
    ca.nrows = 2;
    float *norms = reinterpret_cast<float *>(d + 2 * row_bytes + 16);
    float *res = reinterpret_cast<float *>(d + 2 * row_bytes + 32);
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) sms = 148;
    if (ce == cudaSuccess && metric == SVDB_COSINE) {
        ca.first = 0;
        ca.n = 2;
        ca.out = norms;
        ca.mode = 4;
        ce = launch_compare(ca, sms, st);
    }
    if (ce == cudaSuccess) {
        ca.i1 = reinterpret_cast<const u64 *>(d + 2 * row_bytes);
        ca.i2 = ca.i1 + 1;
        ca.n = 1;
        ca.norm = norms;
        ca.out = res;
        ca.mode = metric;
        ce = launch_compare(ca, sms, st);
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h + 2 * row_bytes + 32, res, 4, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) {
        set_last_error(std::string("svdb_compare_vectors: ") + cudaGetErrorString(ce));
        cudaGetLastError();
        return SVDB_ERR_CUDA;
    }
    memcpy(out, h + 2 * row_bytes + 32, 4);
    return SVDB_OK;
}

// Stream `count` records of `rec` bytes from f into e (caller owns f and e).
static int load_records(svdb_engine *e, FILE *f, uint64_t count, uint64_t dim) {
    std::lock_guard<std::mutex> g(e->mu);
    const size_t rec = REC_HDR + (size_t)dim * 8;
    const size_t chunk_rows = std::max<size_t>(1, ((size_t)64 << 20) / rec);
    PinnedScratch hraw;
    Scratch draw, drows;
    std::string err;
    int rc = SVDB_OK;
    if (!hraw.ensure(chunk_rows * rec, err) || !draw.ensure(chunk_rows * rec, err) ||
        !drows.ensure(chunk_rows * (size_t)dim * 8, err))
        rc = e->fail(SVDB_ERR_OOM, err);
    cudaSetDevice(e->device);
    size_t done = 0;
    while (done < count && rc == SVDB_OK) {
        const size_t m = std::min<size_t>(chunk_rows, count - done);
        const size_t got = fread(hraw.p, rec, m, f);
        if (got == 0) break;   // short file: keep what was read (the reference does not check fread at all)
        const unsigned char *h = hraw.as<unsigned char>();
        for (size_t i = 0; i < got && rc == SVDB_OK; i++) {
            uint64_t d;
            memcpy(&d, h + i * rec + 37, 8);
            if (d != dim) rc = e->fail(SVDB_ERR_ARG, "rows of different dimensions in one file: not supported by the bulk loader");
        }
        if (rc) break;
        cudaError_t ce = cudaMemcpyAsync(draw.p, hraw.p, got * rec, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess) {
            const size_t total = got * (size_t)dim;
            unpack_records_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)e->tune.num_sms * 16), 256, 0, e->stream>>>(
                draw.as<unsigned char>(), rec, (int)dim, got, drows.as<double>());
            ce = cudaGetLastError();
        }
        if (ce != cudaSuccess) {
            rc = e->fail_cuda("bulk load", ce);
            break;
        }
        e->stats.kernels_launched++;
        e->stats.h2d_bytes += got * rec;
        const size_t base = e->cur_host.size();
        rc = e->ingest_device_rows(drows.as<double>(), got, (size_t)dim, nullptr);
        if (rc) break;
        ce = cudaStreamSynchronize(e->stream);   // hraw / draw / drows are reused by the next chunk
        if (ce != cudaSuccess) {
            rc = e->fail_cuda("bulk load", ce);
            break;
        }
        for (size_t i = 0; i < got; i++) {
            memcpy(e->uuids[base + i].data(), h + i * rec, 37);
            e->uuids[base + i][36] = 0;
        }
        done += got;
        if (got < m) break;
    }
    hraw.free_();
    draw.free_();
    drows.free_();
    return rc;
}

int svdb_engine_load_file(const char *path, size_t kd_dim, int device, uint32_t flags, svdb_engine **out) {
    if (!path || !out) {
        set_last_error("NULL argument");
        return SVDB_ERR_ARG;
    }
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_last_error(std::string("cannot open ") + path);
        return SVDB_ERR_ARG;
    }
    uint64_t count = 0, dim = 0;
    char uuid0[37];
    if (fread(&count, 8, 1, f) != 1) count = 0;
    const long data0 = ftell(f);
    if (count && (fread(uuid0, 1, 37, f) != 37 || fread(&dim, 8, 1, f) != 1 || dim == 0)) {
        fclose(f);
        set_last_error("truncated or empty first record");
        return SVDB_ERR_ARG;
    }
    if (count == 0) dim = kd_dim;   // nothing to learn the row dimension from
    svdb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.dimension = (size_t)dim;
    cfg.kd_dim = kd_dim;
    cfg.device = device;
    cfg.flags = flags;
    cfg.reserve_rows = (size_t)count;
    svdb_engine *e = nullptr;
    int rc = svdb_engine_create(&cfg, &e);
    if (rc == SVDB_OK && count) {
        fseek(f, data0, SEEK_SET);
        rc = load_records(e, f, count, dim);
        if (rc) {
            const std::string keep = get_last_error();
            svdb_engine_destroy(e);
            e = nullptr;
            set_last_error(keep);
        }
    }
    fclose(f);
    if (rc) return rc;
    *out = e;
    return SVDB_OK;
}

int svdb_save_file(svdb_engine *e, const char *path) {
    if (!e || !path) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->log_only) return e->fail(SVDB_ERR_ARG, "a log-only engine stores no rows");
    int rc = e->flush();
    if (rc) return rc;
    rc = e->upload_cur();
    if (rc) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return e->fail(SVDB_ERR_ARG, std::string("cannot open ") + path + " for writing");
    const uint64_t count = e->cur_host.size(), dim = (uint64_t)e->D;
    fwrite(&count, 8, 1, f);
    const size_t rec = REC_HDR + (size_t)dim * 8;
    const size_t chunk_rows = std::max<size_t>(1, ((size_t)64 << 20) / rec);
    PinnedScratch hraw;
    Scratch draw;
    std::string err;
    if (count && (!hraw.ensure(chunk_rows * rec, err) || !draw.ensure(chunk_rows * rec, err))) {
        fclose(f);
        return e->fail(SVDB_ERR_OOM, err);
    }
    cudaSetDevice(e->device);
    rc = SVDB_OK;
    for (size_t done = 0; done < count; done += chunk_rows) {
        const size_t m = std::min<size_t>(chunk_rows, count - done);
        const size_t total = m * (size_t)dim;
        pack_records_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)e->tune.num_sms * 16), 256, 0, e->stream>>>(
            e->rows.as<double>(), e->Dpad, e->cur.as<unsigned long long>(), done, (int)dim, m, rec, draw.as<unsigned char>());
        cudaError_t ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(hraw.p, draw.p, m * rec, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) {
            rc = e->fail_cuda("save", ce);
            break;
        }
        e->stats.kernels_launched++;
        e->stats.d2h_bytes += m * rec;
        unsigned char *h = hraw.as<unsigned char>();
        for (size_t i = 0; i < m; i++) {
            memcpy(h + i * rec, e->uuids[done + i].data(), 37);
            memcpy(h + i * rec + 37, &dim, 8);
        }
        if (fwrite(h, rec, m, f) != m) {
            rc = e->fail(SVDB_ERR_ARG, "short write");
            break;
        }
    }
    fclose(f);
    hraw.free_();
    draw.free_();
    return rc;
}

int svdb_get_uuid(svdb_engine *e, size_t index, char out[37]) {
    if (!e || !out) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (index >= e->cur_host.size()) return e->fail(SVDB_ERR_RANGE, "index out of range");
    memcpy(out, e->uuids[index].data(), 37);
    out[36] = 0;
    return SVDB_OK;
}

int svdb_set_uuid(svdb_engine *e, size_t index, const char *uuid) {
    if (!e || !uuid) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (index >= e->cur_host.size()) return e->fail(SVDB_ERR_RANGE, "index out of range");
    memset(e->uuids[index].data(), 0, 37);
    strncpy(e->uuids[index].data(), uuid, 36);
    return SVDB_OK;
}

int svdb_get_stats(const svdb_engine *e, svdb_stats *out) {
    if (!e || !out) return SVDB_ERR_ARG;
    *out = e->stats;
    return SVDB_OK;
}

int svdb_set_option(svdb_engine *e, const char *name, long value) {
    if (!e || !name) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    const std::string n(name);
    e->opt_gen++;
    if (n == "host.graphs") { e->graphs_enabled = value != 0; return SVDB_OK; }
    if (n == "scan.variant") e->tune.variant = (int)value;
    else if (n == "scan.warps") e->tune.warps = (int)value;
    else if (n == "scan.stages") e->tune.stages = (int)value;
    else if (n == "scan.tile_rows") e->tune.tile_rows = (int)value;
    else if (n == "scan.ctas_per_sm") e->tune.ctas_per_sm = (int)value;
    else if (n == "scan.assign") e->tune.assign = (int)value;
    else if (n == "scan.nq_per_pass") e->tune.nq_per_pass = (int)value;
    else if (n == "scan.force_exact") e->force_exact = value != 0;
    else if (n == "nearest.tree_max_k") e->tree_max_k = (int)value;
    else if (n == "log.index_base") e->index_base = (uint64_t)value;
    else if (n == "nearest.mtree") e->mtree_auto = value != 0;
    else if (n == "mtree.lanes") {
        if (value != 0 && value != 32 && value != 16 && value != 8) return e->fail(SVDB_ERR_ARG, "mtree.lanes must be 0 (auto), 32, 16 or 8");
        e->mtree_lanes = (int)value;
    }
    else if (n == "mtree.block_levels") {
        if (value != 1 && value != 3) return e->fail(SVDB_ERR_ARG, "mtree.block_levels must be 1 or 3");
        e->mtree_block = (int)value;
    }
    else if (n == "mtree.tail_max") e->mtree_tail_max = (size_t)std::max(0l, value);
    else if (n == "mtree.tail_min") e->mtree_tail_min = (size_t)std::max(0l, value);
    else if (n == "tree.max_depth") e->tree_max_depth = (int)value;
    else if (n == "nearest.mma_min_queries") e->mma_min_q = (int)value, e->mma_min_user = true;
    else if (n == "nearest.umma_min_queries") e->umma_min_q = (int)value, e->umma_min_user = true;
    else if (n == "scan.plane8_max_queries") e->plane8_max_q = (int)std::max(0l, value), e->plane8_max_q_user = true;
    else if (n == "scan.overlap_steps") e->overlap_steps = value != 0;
    else if (n == "umma.group_min") e->umma_group_min = value != 0;
    else if (n == "umma.sparse_checks") e->umma_sparse_checks = value != 0;
    else if (n == "scan.plane8_pair") e->plane8_pair = value != 0;
    else if (n == "nearest.umma_min_kd_dim") e->umma_min_k = (int)std::max(1l, value);
    else if (n == "scan.plane_max_k") e->plane_max_k = (int)std::max(0l, value);
    else if (n == "scan.plane8_max_k") e->plane8_max_k = (int)std::max(0l, value);
    else if (n == "umma.debug_keys") e->umma_debug = value != 0;
    else if (n == "umma.resident_queries") e->umma_resident = value != 0;
    else if (n == "scan.shadow") e->scan_plane = value != 0 ? 1 : 0;      // round-1 name: 1 = K11 (hi + lo planes), 0 = fp64 rows
    else if (n == "scan.plane") {
        if (value < 0 || value > 3) return e->fail(SVDB_ERR_ARG, "scan.plane must be 0 (fp64 rows), 1 (hi + lo planes), 2 (hi plane) or 3 (byte plane)");
        e->scan_plane = (int)value;
    }
    else if (n == "scan.fuse_tail") e->fuse_tail = value != 0;
    else if (n == "scan.tail_debug") e->tail_debug = value != 0;
    else if (n == "scan.dynamic_tiles") e->dyn_tiles = (int)std::min(8l, std::max(0l, value));
    else if (n == "profile.scan_events") e->profile_scan = value != 0;
    else return e->fail(SVDB_ERR_ARG, "unknown option " + n);
    return SVDB_OK;
}

int svdb_time_scan(svdb_engine *e, const double *d_Q, size_t nq, size_t ldq, size_t k, int iters, float *ms_out) {
    if (!e || !d_Q || !ms_out || iters < 1 || nq < 1 || k < 1 || k > SVDB_MAX_K) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    int rc = e->flush();
    if (rc) return rc;
    if (e->n_versions == 0 || e->no_log) return e->fail(SVDB_ERR_ARG, "empty log");
    cudaSetDevice(e->device);
    const bool use_exact = e->force_exact || !e->wide;
    const int nqp = largest_pass(nq, std::max(1, e->tune.nq_per_pass));
    const int cap = (int)std::min<size_t>(32, k + 8);
    const int nlists = scan_num_lists(e->tune, !use_exact);
    std::string err;
    if (!e->lists.ensure((size_t)8 * nlists * cap * sizeof(Cand), err)) return e->fail(SVDB_ERR_OOM, err);
    const double *qbase = d_Q;
    int qld = (int)ldq;
    if (!use_exact) {
        if (!e->qpad.ensure(nq * (size_t)e->kstride * 8, err)) return e->fail(SVDB_ERR_OOM, err);
        cudaError_t ce = launch_pad_queries(d_Q, (int)ldq, e->qpad.as<double>(), e->kstride, e->K, (int)nq, e->stream);
        if (ce != cudaSuccess) return e->fail_cuda("pad_queries", ce);
        qbase = e->qpad.as<double>();
        qld = e->kstride;
    }
    ScanArgs sa{};
    sa.pts = e->kd_ptr();
    sa.n = e->n_versions;
    sa.K = e->K;
    sa.stride = e->kstride;
    sa.q = qbase;
    sa.ldq = qld;
    sa.nq = nqp;
    sa.cap = cap;
    sa.lists = e->lists.as<Cand>();
    sa.assign = e->tune.assign;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaError_t ce = use_exact ? launch_scan_exact(e->tune, sa, e->stream) : launch_scan_wide(e->tune, sa, e->stream);
    cudaEventRecord(a, e->stream);
    for (int i = 0; i < iters && ce == cudaSuccess; i++)
        ce = use_exact ? launch_scan_exact(e->tune, sa, e->stream) : launch_scan_wide(e->tune, sa, e->stream);
    cudaEventRecord(b, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    float ms = 0;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (ce != cudaSuccess) return e->fail_cuda("time_scan", ce);
    e->stats.kernels_launched += iters + 1;
    *ms_out = ms / iters;
    return SVDB_OK;
}

/* Sum of the scan-kernel durations recorded since the last call (option profile.scan_events). */
int svdb_take_scan_time(svdb_engine *e, float *total_ms, uint64_t *launches) {
    if (!e || !total_ms || !launches) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    cudaSetDevice(e->device);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    float sum = 0;
    for (size_t i = 0; i < e->scan_events_used && ce == cudaSuccess; i++) {
        float ms = 0;
        ce = cudaEventElapsedTime(&ms, e->scan_events[i].first, e->scan_events[i].second);
        sum += ms;
    }
    if (ce != cudaSuccess) return e->fail_cuda("take_scan_time", ce);
    *total_ms = sum;
    *launches = e->scan_events_used;
    e->scan_events_used = 0;
    return SVDB_OK;
}

/* Diagnostics of the fused tail (option "scan.tail_debug" = 1): the %globaltimer stamps (ns) the last fused scan launch
 * left -- see kernels.h: TailArgs::dbg; count <= 8 + number of CTAs. */
int svdb_debug_tail_times(svdb_engine *e, unsigned long long *out, size_t count) {
    if (!e || !out) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->tail_dbg.p || count * 8 > e->tail_dbg.cap) return e->fail(SVDB_ERR_ARG, "no tail debug capture (set scan.tail_debug before the call)");
    cudaSetDevice(e->device);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, e->tail_dbg.p, count * 8, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return e->fail_cuda("debug_tail_times", ce);
    return SVDB_OK;
}

/* Diagnostics of K13: the byte plane's grid (lo, step), its measured error and the first `nbytes` bytes of the plane. */
int svdb_debug_plane8(svdb_engine *e, double par_out[3], unsigned char *bytes_out, size_t nbytes) {
    if (!e || !par_out) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->plane8_ready) return e->fail(SVDB_ERR_ARG, "no byte plane yet");
    cudaSetDevice(e->device);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    Plane8Par hp;
    unsigned long long errw = 0;
    if (ce == cudaSuccess) ce = cudaMemcpy(&hp, e->plane8_par.p, sizeof hp, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(&errw, e->plane8_par.as<unsigned char>() + sizeof(Plane8Par), 8, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && bytes_out && nbytes) ce = cudaMemcpy(bytes_out, e->plane8.as<unsigned char>(), nbytes, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return e->fail_cuda("debug_plane8", ce);
    par_out[0] = hp.lo;
    par_out[1] = hp.step;
    memcpy(&par_out[2], &errw, 8);
    return SVDB_OK;
}

/* Diagnostics of K10: after a batch call made with option "umma.debug_keys" = 1, the approximate keys of log rows
 * 0..127 against the first `bn` queries of that call, [128][bn] floats (bn = 64, 128 or 256 by batch size). */
int svdb_debug_filter_keys(svdb_engine *e, float *keys_out, size_t count) {
    if (!e || !keys_out) return SVDB_ERR_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->udbg.p || count > (size_t)128 * 256 + 16) return e->fail(SVDB_ERR_ARG, "no K10 debug capture (set umma.debug_keys before the call)");
    cudaSetDevice(e->device);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpy(keys_out, e->udbg.p, count * 4, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return e->fail_cuda("debug_filter_keys", ce);
    return SVDB_OK;
}

}  // extern "C"
