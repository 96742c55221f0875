// tie_protocol.cu -- which of several DISTINCT kd-points at exactly the minimal distance does the
// reference return, when the log is sharded over several GPUs and no rank holds the whole tree?
//
// The reference's tree (src/kdtree.c:47-62) is an insertion-order BST: the node below node P on side s
// is the EARLIEST log entry after P that falls into P's cell on side s.  So the path of the global tree
// can be walked without ever materialising the tree: "first entry inside a cell" is a scan of every
// shard's slice followed by a min over the shards.  kdtree_nearest_rec (:131-162) keeps the first
// minimum it reaches in near-side-first preorder (strict '<', :139), and that node is never pruned (see
// tree.cuh).  Hence the walk (same rule as resolve_tie() in tree.cuh, one level per round):
//     node P = first entry in the current cell;  P is tied -> P wins;
//     else go to the query's side of P if any tied entry lives there, otherwise to the other side;
//     one tied entry left in the cell -> it wins.
//
// svdb_tie_resolve() is that walk for a batch of queries in lock-step, on the host, written against a
// backend: three shard-local primitives (collect / first_in_cell / split) and an all-gather.  The CUDA
// backend (below) implements the primitives with kernels over this engine's slice of the log; tests
// drive the same walk with a numpy backend over gloo (tests/test_sharding_gloo.py).
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.h"
#include "kernels.h"

using namespace svdb;

// =====================================================================================
// the walk (host, transport- and device-agnostic)
// =====================================================================================
namespace {

struct TieEvent {
    size_t q;                       // query number
    uint32_t depth = 0;
    uint64_t after = ~0ull;         // global seq of the last node walked through
    std::vector<double> pv;         // the cell: (value, side) per level
    std::vector<uint8_t> ps;
    bool done = false;
    uint64_t win_seq = ~0ull, win_index = ~0ull;
};

void apply_winner(svdb_candidate *r, size_t k, double dstar, uint64_t seq, uint64_t index) {
    // winner first, the rest stay in (dist, seq) order -- what finalize does on one GPU
    size_t pos = k - 1;
    for (size_t j = 0; j < k; j++)
        if (r[j].seq == seq) { pos = j; break; }
    for (size_t j = pos; j > 0; j--) r[j] = r[j - 1];
    r[0].dist = dstar;
    r[0].seq = seq;
    r[0].index = index;
    for (size_t j = 0; j < k; j++) r[j].flags &= ~SVDB_CAND_TIE;
}

}  // namespace

extern "C" int svdb_tie_resolve(const svdb_tie_backend *b, const double *Q, size_t nq, size_t ldq, svdb_candidate *merged,
                                size_t k, uint64_t *levels_walked) {
    if (levels_walked) *levels_walked = 0;
    if (!b || !b->allgather || !b->collect || !b->first_in_cell || !b->split || b->world < 1 || b->kd_dim < 1 ||
        (nq && (!Q || !merged)) || k < 1 || ldq < b->kd_dim) {
        set_last_error("bad argument to svdb_tie_resolve");
        return SVDB_ERR_ARG;
    }
    const size_t K = b->kd_dim;
    const int W = b->world;
    std::vector<TieEvent> ev;
    for (size_t i = 0; i < nq; i++)
        if ((merged[i * k].flags & SVDB_CAND_TIE) && merged[i * k].seq != ~0ull) {
            TieEvent e;
            e.q = i;
            ev.push_back(std::move(e));
        }
    const size_t ne = ev.size();
    if (ne == 0) return SVDB_OK;             // the merged flags are the same on every rank: nobody communicates

    // ---- the tied sets ----
    std::vector<double> qs(ne * K), dstar(ne), first(ne * K, 0.0);
    std::vector<uint64_t> nloc(ne, 0);
    std::vector<uint8_t> same(ne, 1);
    for (size_t i = 0; i < ne; i++) {
        memcpy(&qs[i * K], Q + ev[i].q * ldq, K * 8);
        dstar[i] = merged[ev[i].q * k].dist;
    }
    int rc = b->collect(b->ctx, ne, qs.data(), dstar.data(), nloc.data(), same.data(), first.data());
    if (rc) return rc;
    // one message per event: { n_local, all-the-same-point, coordinates of the local first }
    const size_t mlen = 2 + K;
    std::vector<uint64_t> snd(ne * mlen), rcv((size_t)W * ne * mlen);
    for (size_t i = 0; i < ne; i++) {
        snd[i * mlen] = nloc[i];
        snd[i * mlen + 1] = same[i];
        memcpy(&snd[i * mlen + 2], &first[i * K], K * 8);
    }
    rc = b->allgather(b->allgather_ctx, snd.data(), rcv.data(), ne * mlen * 8);
    if (rc) return rc;
    for (size_t i = 0; i < ne; i++) {
        uint64_t total = 0;
        bool dup_only = true;
        const uint64_t *ref = nullptr;
        for (int r = 0; r < W; r++) {
            const uint64_t *m = &rcv[((size_t)r * ne + i) * mlen];
            if (m[0] == 0) continue;
            total += m[0];
            if (!m[1]) dup_only = false;
            if (!ref) ref = m + 2;
            else if (memcmp(ref, m + 2, K * 8) != 0) dup_only = false;    // bit patterns: -0.0 vs 0.0 walks the tree
        }
        if (total == 0) {
            set_last_error("svdb_tie_resolve: no shard holds an entry at the merged minimal distance "
                           "(were the merged candidates identical on every rank?)");
            return SVDB_ERR_STATE;
        }
        if (total == 1 || dup_only) {
            // identical kd-points: the earliest insert is an ancestor of the others (kdtree.c:52 sends
            // equal coordinates right) and strict '<' keeps it -- that is merged[0] already
            ev[i].done = true;
            ev[i].win_seq = merged[ev[i].q * k].seq;
            ev[i].win_index = merged[ev[i].q * k].index;
        }
    }

    // ---- the walk, all unresolved events one level per round ----
    uint64_t levels = 0;
    std::vector<uint32_t> aev, adepth;
    std::vector<uint64_t> aafter;
    std::vector<double> apv, av;
    std::vector<uint8_t> aps;
    std::vector<svdb_tie_first> f_loc, f_all;
    std::vector<svdb_tie_split> s_loc, s_all;
    for (;;) {
        aev.clear();
        for (size_t i = 0; i < ne; i++)
            if (!ev[i].done) aev.push_back((uint32_t)i);
        size_t na = aev.size();
        if (na == 0) break;
        size_t ld = 1;
        for (uint32_t i : aev) ld = std::max<size_t>(ld, ev[i].depth);
        adepth.resize(na);
        aafter.resize(na);
        apv.assign(na * ld, 0.0);
        aps.assign(na * ld, 0);
        for (size_t a = 0; a < na; a++) {
            const TieEvent &e = ev[aev[a]];
            adepth[a] = e.depth;
            aafter[a] = e.after;
            if (e.depth) {
                memcpy(&apv[a * ld], e.pv.data(), e.depth * 8);
                memcpy(&aps[a * ld], e.ps.data(), e.depth);
            }
        }
        // (A) the node at this level: first entry of the cell over all shards
        f_loc.resize(na);
        f_all.resize((size_t)W * na);
        rc = b->first_in_cell(b->ctx, na, aev.data(), adepth.data(), apv.data(), aps.data(), ld, aafter.data(), f_loc.data());
        if (rc) return rc;
        rc = b->allgather(b->allgather_ctx, f_loc.data(), f_all.data(), na * sizeof(svdb_tie_first));
        if (rc) return rc;
        av.resize(na);
        std::vector<svdb_tie_first> node(na);
        for (size_t a = 0; a < na; a++) {
            svdb_tie_first best{~0ull, ~0ull, 0.0, 0};
            for (int r = 0; r < W; r++) {
                const svdb_tie_first &c = f_all[(size_t)r * na + a];
                if (c.seq < best.seq) best = c;
            }
            if (best.seq == ~0ull) {
                set_last_error("svdb_tie_resolve: empty cell on the way to a tied entry (inconsistent shards)");
                return SVDB_ERR_STATE;
            }
            node[a] = best;
            av[a] = best.v;
            if (best.tied) {          // visited before everything below it
                TieEvent &e = ev[aev[a]];
                e.done = true;
                e.win_seq = best.seq;
                e.win_index = best.index;
            }
        }
        levels += na;
        // (B) where do the tied entries of the cell live, relative to the node?
        s_loc.resize(na);
        s_all.resize((size_t)W * na);
        rc = b->split(b->ctx, na, aev.data(), adepth.data(), apv.data(), aps.data(), ld, av.data(), s_loc.data());
        if (rc) return rc;
        rc = b->allgather(b->allgather_ctx, s_loc.data(), s_all.data(), na * sizeof(svdb_tie_split));
        if (rc) return rc;
        for (size_t a = 0; a < na; a++) {
            TieEvent &e = ev[aev[a]];
            if (e.done) continue;
            uint64_t n[2] = {0, 0}, ms[2] = {~0ull, ~0ull}, mi[2] = {~0ull, ~0ull};
            for (int r = 0; r < W; r++) {
                const svdb_tie_split &c = s_all[(size_t)r * na + a];
                for (int s = 0; s < 2; s++) {
                    n[s] += c.n[s];
                    if (c.min_seq[s] < ms[s]) {
                        ms[s] = c.min_seq[s];
                        mi[s] = c.min_index[s];
                    }
                }
            }
            const int axis = (int)(e.depth % K);
            const int near_side = Q[e.q * ldq + axis] < node[a].v ? 0 : 1;          // kdtree.c:147
            const int side = n[near_side] > 0 ? near_side : 1 - near_side;
            if (n[side] == 0) {
                set_last_error("svdb_tie_resolve: the tied entries left the cell (inconsistent shards)");
                return SVDB_ERR_STATE;
            }
            if (n[side] == 1) {
                e.done = true;
                e.win_seq = ms[side];
                e.win_index = mi[side];
                continue;
            }
            e.pv.push_back(node[a].v);
            e.ps.push_back((uint8_t)side);
            e.depth++;
            e.after = node[a].seq;
        }
    }
    for (size_t i = 0; i < ne; i++) apply_winner(merged + ev[i].q * k, k, dstar[i], ev[i].win_seq, ev[i].win_index);
    if (levels_walked) *levels_walked = levels;
    return SVDB_OK;
}

// =====================================================================================
// CUDA backend: the three primitives over this engine's slice of the log
// =====================================================================================
namespace svdb {

constexpr int TIE_EV_PER_PASS = 8;

// entries at exactly the reference distance dstar[e] from query e (kdtree.c:134-137 order; the running sum
// never decreases, so an entry is dropped as soon as it exceeds every event's distance)
__global__ void __launch_bounds__(256) tie_collect_kernel(const double *__restrict__ pts, u64 n, int K, int stride,
                                                          const double *__restrict__ q, const double *__restrict__ dstar,
                                                          int ne, int ev0, u64 *__restrict__ out_seq,
                                                          uint32_t *__restrict__ out_ev, unsigned long long *counter, u64 cap) {
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (u64)gridDim.x * blockDim.x) {
        const double *row = pts + s * (u64)stride;
        double acc[TIE_EV_PER_PASS];
#pragma unroll
        for (int e = 0; e < TIE_EV_PER_PASS; e++) acc[e] = 0.0;
        unsigned live = (1u << ne) - 1u;
        for (int i0 = 0; i0 < K && live; i0 += 16) {
            const int i1 = min(K, i0 + 16);
            for (int i = i0; i < i1; i++) {
                const double x = __ldg(row + i);
#pragma unroll
                for (int e = 0; e < TIE_EV_PER_PASS; e++) {
                    if (e < ne) {
                        const double t = __dsub_rn(x, __ldg(q + (size_t)e * K + i));
                        acc[e] = __dadd_rn(acc[e], __dmul_rn(t, t));
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < TIE_EV_PER_PASS; e++)
                if (e < ne && acc[e] > __ldg(dstar + e)) live &= ~(1u << e);
        }
#pragma unroll
        for (int e = 0; e < TIE_EV_PER_PASS; e++) {
            if (e < ne && ((live >> e) & 1u) && acc[e] == __ldg(dstar + e)) {
                const u64 pos = atomicAdd(counter, 1ull);
                if (pos < cap) {
                    out_seq[pos] = s;
                    out_ev[pos] = (uint32_t)(ev0 + e);
                }
            }
        }
    }
}

// per event: are all tied entries the same kd-point?  + the coordinates of the first (lowest seq) one
__global__ void __launch_bounds__(128) tie_summary_kernel(const double *__restrict__ pts, int K, int stride,
                                                          const u64 *__restrict__ t_seq, const u64 *__restrict__ t_off,
                                                          unsigned *__restrict__ differs, double *__restrict__ first) {
    const int e = blockIdx.x;
    const u64 lo = t_off[e], hi = t_off[e + 1];
    if (lo == hi) return;
    const double *r0 = pts + t_seq[lo] * (u64)stride;
    for (int i = threadIdx.x; i < K; i += blockDim.x) first[(size_t)e * K + i] = r0[i];
    bool d = false;
    const u64 total = (hi - lo - 1) * (u64)K;
    for (u64 w = threadIdx.x; w < total; w += blockDim.x) {
        const u64 m = lo + 1 + w / K;
        const int i = (int)(w % K);
        // bit patterns, not values: +0.0 and -0.0 compare equal but this check promises "same point"
        d |= __double_as_longlong(pts[t_seq[m] * (u64)stride + i]) != __double_as_longlong(r0[i]);
    }
    if (d) atomicOr(differs + e, 1u);
}

__device__ __forceinline__ bool in_cell(const double *__restrict__ row, int K, int depth, const double *__restrict__ pv,
                                        const uint8_t *__restrict__ ps) {
    for (int j = 0; j < depth; j++) {
        const int side = row[j % K] < pv[j] ? 0 : 1;                    // kdtree.c:52
        if (side != ps[j]) return false;
    }
    return true;
}

// found[a] = lowest local seq inside cell a (with global seq > after[a]); entries are visited in
// increasing seq per thread, so a thread stops once nothing it could still reach can improve any event
__global__ void __launch_bounds__(256) tie_first_kernel(const double *__restrict__ pts, u64 n, int K, int stride, int na,
                                                        const uint32_t *__restrict__ depth, const double *__restrict__ pv,
                                                        const uint8_t *__restrict__ ps, int path_ld,
                                                        const u64 *__restrict__ after, u64 seq_base, u64 *found) {
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (u64)gridDim.x * blockDim.x) {
        bool useful = false;
        const double *row = pts + s * (u64)stride;
        for (int a = 0; a < na; a++) {
            if (s >= *((volatile u64 *)(found + a))) continue;
            useful = true;
            const u64 af = after[a];
            if (af != SEQ_NONE && s + seq_base <= af) continue;
            if (in_cell(row, K, (int)depth[a], pv + (size_t)a * path_ld, ps + (size_t)a * path_ld)) atomicMin(found + a, s);
        }
        if (!useful) break;
    }
}

__global__ void tie_first_info_kernel(const double *__restrict__ pts, int K, int stride, int na,
                                      const uint32_t *__restrict__ ev, const uint32_t *__restrict__ depth,
                                      const u64 *__restrict__ found, const u64 *__restrict__ t_seq,
                                      const u64 *__restrict__ t_off, const u64 *__restrict__ log_index, u64 seq_base,
                                      svdb_tie_first *__restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= na) return;
    svdb_tie_first r;
    const u64 s = found[a];
    if (s == SEQ_NONE) {
        r.seq = SEQ_NONE;
        r.index = (u64)SVDB_NONE;
        r.v = 0.0;
        r.tied = 0;
    } else {
        r.seq = s + seq_base;
        r.index = log_index[s];
        r.v = pts[s * (u64)stride + depth[a] % K];
        // t_seq is sorted within an event
        u64 lo = t_off[ev[a]], hi = t_off[ev[a] + 1];
        while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            if (t_seq[mid] < s) lo = mid + 1;
            else hi = mid;
        }
        r.tied = (lo < t_off[ev[a] + 1] && t_seq[lo] == s) ? 1 : 0;
    }
    out[a] = r;
}

__global__ void __launch_bounds__(128) tie_split_kernel(const double *__restrict__ pts, int K, int stride,
                                                        const uint32_t *__restrict__ ev, const uint32_t *__restrict__ depth,
                                                        const double *__restrict__ pv, const uint8_t *__restrict__ ps,
                                                        int path_ld, const double *__restrict__ v,
                                                        const u64 *__restrict__ t_seq, const u64 *__restrict__ t_off,
                                                        const u64 *__restrict__ log_index, u64 seq_base,
                                                        svdb_tie_split *__restrict__ out) {
    const int a = blockIdx.x;
    __shared__ unsigned long long s_n[2], s_min[2];
    if (threadIdx.x < 2) {
        s_n[threadIdx.x] = 0;
        s_min[threadIdx.x] = SEQ_NONE;
    }
    __syncthreads();
    const int d = (int)depth[a];
    const int axis = d % K;
    const double split = v[a];
    for (u64 m = t_off[ev[a]] + threadIdx.x; m < t_off[ev[a] + 1]; m += blockDim.x) {
        const u64 s = t_seq[m];
        const double *row = pts + s * (u64)stride;
        if (!in_cell(row, K, d, pv + (size_t)a * path_ld, ps + (size_t)a * path_ld)) continue;
        const int side = row[axis] < split ? 0 : 1;
        atomicAdd(&s_n[side], 1ull);
        atomicMin(&s_min[side], s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        svdb_tie_split r;
        for (int s = 0; s < 2; s++) {
            r.n[s] = s_n[s];
            r.min_seq[s] = s_min[s] == SEQ_NONE ? SEQ_NONE : s_min[s] + seq_base;
            r.min_index[s] = s_min[s] == SEQ_NONE ? (u64)SVDB_NONE : log_index[s_min[s]];
        }
        out[a] = r;
    }
}

// ---- host side of the backend ----
struct TieGpu {
    svdb_engine *e = nullptr;
    Scratch d_q, d_dstar, d_pairs_seq, d_pairs_ev, d_counter, d_tseq, d_toff, d_differs, d_first, d_ev, d_depth, d_pv, d_ps,
        d_after, d_found, d_v, d_out;
    std::vector<u64> t_off;      // host copy of the segment offsets
    size_t ne = 0;
    ~TieGpu() {
        for (Scratch *s : {&d_q, &d_dstar, &d_pairs_seq, &d_pairs_ev, &d_counter, &d_tseq, &d_toff, &d_differs, &d_first, &d_ev,
                           &d_depth, &d_pv, &d_ps, &d_after, &d_found, &d_v, &d_out})
            s->free_();
    }
};

void tie_state_free(TieGpu *t) { delete t; }

#define TCK(call)                                             \
    do {                                                      \
        cudaError_t ce_ = (call);                             \
        if (ce_ != cudaSuccess) return t->e->fail_cuda(#call, ce_); \
    } while (0)

static int tie_collect(void *ctx, size_t ne, const double *queries, const double *dstar, uint64_t *n_local, uint8_t *same,
                       double *first) {
    TieGpu *t = static_cast<TieGpu *>(ctx);
    svdb_engine *e = t->e;
    const int K = e->K;
    cudaStream_t st = e->stream;
    std::string err;
    t->ne = ne;
    u64 cap = std::max<u64>(4096, ne * 64);
    if (!t->d_q.ensure(ne * (size_t)K * 8, err) || !t->d_dstar.ensure(ne * 8, err) || !t->d_counter.ensure(8, err) ||
        !t->d_toff.ensure((ne + 1) * 8, err) || !t->d_differs.ensure(ne * 4, err) || !t->d_first.ensure(ne * (size_t)K * 8, err))
        return e->fail(SVDB_ERR_OOM, err);
    TCK(cudaMemcpyAsync(t->d_q.p, queries, ne * (size_t)K * 8, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_dstar.p, dstar, ne * 8, cudaMemcpyHostToDevice, st));
    std::vector<u64> pseq;
    std::vector<uint32_t> pev;
    for (;;) {
        if (!t->d_pairs_seq.ensure(cap * 8, err) || !t->d_pairs_ev.ensure(cap * 4, err)) return e->fail(SVDB_ERR_OOM, err);
        TCK(cudaMemsetAsync(t->d_counter.p, 0, 8, st));
        const u64 n = e->n_versions;
        if (n) {
            const int grid = (int)std::min<u64>((n + 255) / 256, (u64)e->tune.num_sms * 8);
            for (size_t e0 = 0; e0 < ne; e0 += TIE_EV_PER_PASS) {
                const int nep = (int)std::min<size_t>(TIE_EV_PER_PASS, ne - e0);
                tie_collect_kernel<<<grid, 256, 0, st>>>(e->kd_ptr(), n, K, e->kstride, t->d_q.as<double>() + e0 * K,
                                                         t->d_dstar.as<double>() + e0, nep, (int)e0, t->d_pairs_seq.as<u64>(),
                                                         t->d_pairs_ev.as<uint32_t>(), t->d_counter.as<unsigned long long>(), cap);
                e->stats.kernels_launched++;
            }
            TCK(cudaGetLastError());
        }
        unsigned long long cnt = 0;
        TCK(cudaMemcpyAsync(&cnt, t->d_counter.p, 8, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
        if (cnt > cap) {        // more tied entries than room: the count is exact, run again with enough
            cap = cnt;
            continue;
        }
        pseq.resize(cnt);
        pev.resize(cnt);
        if (cnt) {
            TCK(cudaMemcpyAsync(pseq.data(), t->d_pairs_seq.p, cnt * 8, cudaMemcpyDeviceToHost, st));
            TCK(cudaMemcpyAsync(pev.data(), t->d_pairs_ev.p, cnt * 4, cudaMemcpyDeviceToHost, st));
            TCK(cudaStreamSynchronize(st));
        }
        break;
    }
    // group by event, ascending seq inside an event
    t->t_off.assign(ne + 1, 0);
    for (uint32_t x : pev) t->t_off[x + 1]++;
    for (size_t i = 0; i < ne; i++) t->t_off[i + 1] += t->t_off[i];
    std::vector<u64> tseq(pseq.size()), fill(t->t_off.begin(), t->t_off.end() - 1);
    for (size_t i = 0; i < pseq.size(); i++) tseq[fill[pev[i]]++] = pseq[i];
    for (size_t i = 0; i < ne; i++) std::sort(tseq.begin() + t->t_off[i], tseq.begin() + t->t_off[i + 1]);
    if (!t->d_tseq.ensure(std::max<size_t>(8, tseq.size() * 8), err)) return e->fail(SVDB_ERR_OOM, err);
    if (!tseq.empty()) TCK(cudaMemcpyAsync(t->d_tseq.p, tseq.data(), tseq.size() * 8, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_toff.p, t->t_off.data(), (ne + 1) * 8, cudaMemcpyHostToDevice, st));
    TCK(cudaMemsetAsync(t->d_differs.p, 0, ne * 4, st));
    TCK(cudaMemsetAsync(t->d_first.p, 0, ne * (size_t)K * 8, st));
    tie_summary_kernel<<<(unsigned)ne, 128, 0, st>>>(e->kd_ptr(), K, e->kstride, t->d_tseq.as<u64>(), t->d_toff.as<u64>(),
                                                     t->d_differs.as<unsigned>(), t->d_first.as<double>());
    TCK(cudaGetLastError());
    e->stats.kernels_launched++;
    std::vector<unsigned> differs(ne);
    TCK(cudaMemcpyAsync(differs.data(), t->d_differs.p, ne * 4, cudaMemcpyDeviceToHost, st));
    TCK(cudaMemcpyAsync(first, t->d_first.p, ne * (size_t)K * 8, cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));       // also keeps tseq / t_off alive until the uploads are done
    for (size_t i = 0; i < ne; i++) {
        n_local[i] = t->t_off[i + 1] - t->t_off[i];
        same[i] = differs[i] ? 0 : 1;
    }
    return SVDB_OK;
}

static int tie_upload_cells(TieGpu *t, size_t na, const uint32_t *ev, const uint32_t *depth, const double *pv, const uint8_t *ps,
                            size_t ld) {
    svdb_engine *e = t->e;
    cudaStream_t st = e->stream;
    std::string err;
    if (!t->d_ev.ensure(na * 4, err) || !t->d_depth.ensure(na * 4, err) || !t->d_pv.ensure(na * ld * 8, err) ||
        !t->d_ps.ensure(na * ld, err) || !t->d_after.ensure(na * 8, err) || !t->d_found.ensure(na * 8, err) ||
        !t->d_v.ensure(na * 8, err) || !t->d_out.ensure(na * sizeof(svdb_tie_split), err))
        return e->fail(SVDB_ERR_OOM, err);
    TCK(cudaMemcpyAsync(t->d_ev.p, ev, na * 4, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_depth.p, depth, na * 4, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_pv.p, pv, na * ld * 8, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_ps.p, ps, na * ld, cudaMemcpyHostToDevice, st));
    return SVDB_OK;
}

static int tie_first_in_cell(void *ctx, size_t na, const uint32_t *ev, const uint32_t *depth, const double *pv,
                             const uint8_t *ps, size_t ld, const uint64_t *after, svdb_tie_first *out) {
    TieGpu *t = static_cast<TieGpu *>(ctx);
    svdb_engine *e = t->e;
    cudaStream_t st = e->stream;
    int rc = tie_upload_cells(t, na, ev, depth, pv, ps, ld);
    if (rc) return rc;
    TCK(cudaMemcpyAsync(t->d_after.p, after, na * 8, cudaMemcpyHostToDevice, st));
    TCK(cudaMemsetAsync(t->d_found.p, 0xff, na * 8, st));
    const u64 n = e->n_versions;
    if (n) {
        // blocks in seq order, a few waves: the first wave usually finds the (early) node and the rest stop at once
        const int grid = (int)std::min<u64>((n + 255) / 256, (u64)e->tune.num_sms * 8);
        tie_first_kernel<<<grid, 256, 0, st>>>(e->kd_ptr(), n, e->K, e->kstride, (int)na, t->d_depth.as<uint32_t>(),
                                               t->d_pv.as<double>(), t->d_ps.as<uint8_t>(), (int)ld, t->d_after.as<u64>(),
                                               e->cfg.seq_base, t->d_found.as<u64>());
        e->stats.kernels_launched++;
    }
    tie_first_info_kernel<<<(unsigned)((na + 127) / 128), 128, 0, st>>>(
        e->kd_ptr(), e->K, e->kstride, (int)na, t->d_ev.as<uint32_t>(), t->d_depth.as<uint32_t>(), t->d_found.as<u64>(),
        t->d_tseq.as<u64>(), t->d_toff.as<u64>(), e->log_idx.as<u64>(), e->cfg.seq_base, static_cast<svdb_tie_first *>(t->d_out.p));
    TCK(cudaGetLastError());
    e->stats.kernels_launched++;
    TCK(cudaMemcpyAsync(out, t->d_out.p, na * sizeof(svdb_tie_first), cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));
    return SVDB_OK;
}

static int tie_split(void *ctx, size_t na, const uint32_t *ev, const uint32_t *depth, const double *pv, const uint8_t *ps,
                     size_t ld, const double *v, svdb_tie_split *out) {
    TieGpu *t = static_cast<TieGpu *>(ctx);
    svdb_engine *e = t->e;
    cudaStream_t st = e->stream;
    int rc = tie_upload_cells(t, na, ev, depth, pv, ps, ld);
    if (rc) return rc;
    TCK(cudaMemcpyAsync(t->d_v.p, v, na * 8, cudaMemcpyHostToDevice, st));
    tie_split_kernel<<<(unsigned)na, 128, 0, st>>>(e->kd_ptr(), e->K, e->kstride, t->d_ev.as<uint32_t>(), t->d_depth.as<uint32_t>(),
                                                   t->d_pv.as<double>(), t->d_ps.as<uint8_t>(), (int)ld, t->d_v.as<double>(),
                                                   t->d_tseq.as<u64>(), t->d_toff.as<u64>(), e->log_idx.as<u64>(),
                                                   e->cfg.seq_base, static_cast<svdb_tie_split *>(t->d_out.p));
    TCK(cudaGetLastError());
    e->stats.kernels_launched++;
    TCK(cudaMemcpyAsync(out, t->d_out.p, na * sizeof(svdb_tie_split), cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));
    return SVDB_OK;
}

struct XchAg {
    svdb_exchange *x;
    cudaStream_t st;
};
static int xch_allgather_cb(void *ctx, const void *send, void *recv, size_t bytes) {
    XchAg *a = static_cast<XchAg *>(ctx);
    return exchange_allgather_host(a->x, a->st, send, recv, bytes);
}

// callers hold e->mu
int resolve_ties_engine(svdb_engine *e, svdb_exchange *x, int rank, int world, svdb_allgather_fn ag, void *ag_ctx,
                        const double *Q, size_t nq, size_t ldq, svdb_candidate *merged, size_t k) {
    if (e->no_log) return e->fail(SVDB_ERR_ARG, "this engine was created without a log (SVDB_FLAG_NO_LOG)");
    if (!x && !ag) return e->fail(SVDB_ERR_ARG, "svdb_resolve_ties_sharded needs an exchange or an all-gather callback");
    int rc = e->flush();
    if (rc) return rc;
    cudaError_t ce = cudaSetDevice(e->device);
    if (ce != cudaSuccess) return e->fail_cuda("cudaSetDevice", ce);
    if (!e->tie) {
        e->tie = new (std::nothrow) TieGpu();
        if (!e->tie) return e->fail(SVDB_ERR_OOM, "tie walk state");
        e->tie->e = e;
    }
    XchAg xa{x, e->stream};
    svdb_tie_backend b{};
    b.ctx = e->tie;
    b.world = world;
    b.rank = rank;
    b.kd_dim = (size_t)e->K;
    b.allgather = x ? xch_allgather_cb : ag;
    b.allgather_ctx = x ? static_cast<void *>(&xa) : ag_ctx;
    b.collect = tie_collect;
    b.first_in_cell = tie_first_in_cell;
    b.split = tie_split;
    size_t events = 0;
    for (size_t i = 0; i < nq; i++) events += (merged[i * k].flags & SVDB_CAND_TIE) && merged[i * k].seq != ~0ull;
    uint64_t levels = 0;
    rc = svdb_tie_resolve(&b, Q, nq, ldq, merged, k, &levels);
    e->stats.tie_events += events;
    e->stats.tie_levels += levels;
    return rc;
}

}  // namespace svdb

extern "C" int svdb_resolve_ties_sharded(svdb_engine *e, svdb_exchange *x, int rank, int world, svdb_allgather_fn allgather,
                                         void *allgather_ctx, const double *Q, size_t nq, size_t ldq, svdb_candidate *merged,
                                         size_t k) {
    if (!e) return SVDB_ERR_ARG;
    if (nq == 0) return SVDB_OK;
    if (!Q || !merged || k < 1 || k > SVDB_MAX_K || ldq < (size_t)e->K || world < 1 || rank < 0 || rank >= world)
        return e->fail(SVDB_ERR_ARG, "bad argument to svdb_resolve_ties_sharded");
    std::lock_guard<std::mutex> g(e->mu);
    return resolve_ties_engine(e, x, rank, world, allgather, allgather_ctx, Q, nq, ldq, merged, k);
}
