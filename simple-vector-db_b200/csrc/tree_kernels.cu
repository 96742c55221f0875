// tree_kernels.cu -- K5 (build) and K6 (traversal) of the reference-shaped KD tree, sm_100a.
//
// K5 replaces kdtree_insert -> kdtree_insert_rec (src/kdtree.c:87-91, 47-62) for a whole
// batch of appended log entries at once.  Sequential insertion order matters for the shape,
// so the batch is inserted LEVEL-SYNCHRONOUSLY: every pending entry stands at a node of depth
// r in round r (all start at the root), computes its side (strictly-less / greater-or-equal on
// coordinate r % K) and bids for that child slot with atomicMin(seq).  The smallest sequence
// number wins the slot -- exactly the entry that sequential insertion would have put there --
// and everybody else steps down to the winner in the next round.  Rounds = depth of the
// deepest new entry (about 2.5 log2 N for random order).
//
// K6 replaces kdtree_nearest -> kdtree_nearest_rec (src/kdtree.c:171-178, 131-162): one
// thread per query, explicit stack, the same visit order, the same arithmetic, the same
// strict comparisons -- so the answer is the reference's, ties included.  It is the primary
// path for thin kd-points (the default kd_dim = 3), where the tree prunes to O(log N) visits.
#include <stdlib.h>

#include <algorithm>

#include "kernels.h"
#include "tree.cuh"

namespace svdb {

constexpr uint32_t PEND_DONE = 0xffffffffu;

// One round for entries [n0, n0+m).  pn[i]: node the entry stands at (PEND_DONE when placed);
// pds[i]: depth | side << 31 of its last bid.
template <bool SINGLE_CTA>
__global__ void __launch_bounds__(256) tree_insert_kernel(const double *__restrict__ pts, int stride, int K, uint32_t *child,
                                                          uint32_t n0, uint32_t m, uint32_t *pn, uint32_t *pds,
                                                          int first_round, int rounds, unsigned *still_pending) {
    for (int r = 0; r < rounds; r++) {
        bool mine_pending = false;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
            uint32_t node = pn[i];
            if (node == PEND_DONE) continue;
            const uint32_t e = n0 + i;
            uint32_t depth = pds[i] & 0x7fffffffu;
            if (!(first_round && r == 0)) {
                const uint32_t side = pds[i] >> 31;
                const uint32_t w = __ldcg(&child[2 * (size_t)node + side]);   // settled bids; L2 read (atomics live there)
                if (w == e) {
                    pn[i] = PEND_DONE;
                    continue;
                }
                node = w;
                depth++;
            }
            const int cd = depth % K;
            const uint32_t side = (pts[(size_t)e * stride + cd] < pts[(size_t)node * stride + cd]) ? 0u : 1u;   // kdtree.c:55
            atomicMin(&child[2 * (size_t)node + side], e);
            pn[i] = node;
            pds[i] = depth | (side << 31);
            mine_pending = true;
        }
        if (mine_pending) *still_pending = 1u;
        if (SINGLE_CTA) {
            __threadfence();
            if (!__syncthreads_or(mine_pending ? 1 : 0)) break;   // nobody bid: every entry is placed
        }
    }
}

__global__ void tree_init_kernel(uint32_t *child, uint32_t n0, uint32_t m, uint32_t *pn, uint32_t *pds) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        child[2 * (size_t)(n0 + i)] = NODE_NONE;
        child[2 * (size_t)(n0 + i) + 1] = NODE_NONE;
        pn[i] = (n0 == 0 && i == 0) ? PEND_DONE : 0u;   // the very first entry is the root (kdtree.c:49-51)
        pds[i] = 0u;
    }
}

cudaError_t launch_tree_insert(const double *pts, int stride, int K, uint32_t *child, u64 n0, u64 m, uint32_t *pn,
                               uint32_t *pds, unsigned *d_flag, unsigned *h_flag_pinned, int num_sms, cudaStream_t st,
                               int max_rounds, int *rounds_out) {
    if (m == 0) return cudaSuccess;
    const uint32_t M = (uint32_t)m, N0 = (uint32_t)n0;
    const int grid_all = (int)((m + 255) / 256 > (u64)num_sms * 8 ? (u64)num_sms * 8 : (m + 255) / 256);
    tree_init_kernel<<<grid_all, 256, 0, st>>>(child, N0, M, pn, pds);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int total_rounds = 0;
    int first = 1;
    const bool single = m <= 2048;
    const int per_check = single ? 64 : 16;
    for (;;) {
        e = cudaMemsetAsync(d_flag, 0, sizeof(unsigned), st);
        if (e != cudaSuccess) return e;
        if (single) {
            // one CTA walks many rounds per launch; only the last round's flag matters, so clear it
            // and run one extra probing round afterwards
            tree_insert_kernel<true><<<1, 256, 0, st>>>(pts, stride, K, child, N0, M, pn, pds, first, per_check, d_flag);
            first = 0;
            e = cudaMemsetAsync(d_flag, 0, sizeof(unsigned), st);
            if (e != cudaSuccess) return e;
            tree_insert_kernel<true><<<1, 256, 0, st>>>(pts, stride, K, child, N0, M, pn, pds, 0, 1, d_flag);
            total_rounds += per_check + 1;
        } else {
            for (int r = 0; r < per_check; r++) {
                if (r == per_check - 1) {
                    e = cudaMemsetAsync(d_flag, 0, sizeof(unsigned), st);
                    if (e != cudaSuccess) return e;
                }
                tree_insert_kernel<false><<<grid_all, 256, 0, st>>>(pts, stride, K, child, N0, M, pn, pds, first, 1, d_flag);
                first = 0;
            }
            total_rounds += per_check;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(h_flag_pinned, d_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (*h_flag_pinned == 0) break;
        if (total_rounds > max_rounds) {   // degenerate (e.g. sorted) insertion order: depth ~ N
            if (rounds_out) *rounds_out = -total_rounds;
            return cudaSuccess;
        }
    }
    if (rounds_out) *rounds_out = total_rounds;
    return cudaSuccess;
}

// ---- K6 ---------------------------------------------------------------------------------
constexpr int TREE_STACK = 160;

// One loop trip = one node visit (a lane whose near path ended first pops the deepest far subtree that can still
// matter and visits that), so the lanes of a warp stay in step whichever phase each of them is in.
// Traversal lengths have a heavy tail (profile: 5.6 of 32 lanes active on average), but handing finished lanes
// new queries from a counter was measured SLOWER on big calls (2^20 queries over 10M rows: 8.4-8.8 ms vs
// 7.0-7.3 ms): lanes that start together share the top of the tree, lanes that do not, do not.
__global__ void __launch_bounds__(128) tree_nearest_kernel(const double *__restrict__ pts, int stride, int K,
                                                           const uint32_t *__restrict__ child, u64 n,
                                                           const double *__restrict__ Q, int ldq, int nq,
                                                           const u64 *__restrict__ log_index, u64 seq_base,
                                                           svdb_candidate *out, const unsigned *__restrict__ only_marked) {
    struct Frame {
        uint32_t node, depth;
        double plane;
    };
    Frame st[TREE_STACK];
    // thin queries are copied once into thread-local storage (the host path hands them over in pinned
    // host memory, zero-copy); longer ones are read in place (L1-resident after the first node)
    double ql[16];
    const double *q = ql;                      // (no __restrict__: q may point at ql)
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    if (only_marked && !only_marked[qi]) return;   // K9 answered this one already
    const double *qg = Q + (size_t)qi * ldq;
    if (K <= 16)
        for (int i = 0; i < K; i++) ql[i] = qg[i];
    q = K <= 16 ? ql : qg;
    int sp = 0;
    bool overflow = false;
    double best = CUDART_INF;
    uint32_t best_node = NODE_NONE, cur = n ? 0u : NODE_NONE, depth = 0;
    for (;;) {
        if (cur == NODE_NONE) {
            // the near path ended: back to the deepest far subtree that can still hold something closer
            bool found = false;
            while (sp > 0) {
                sp--;
                if (st[sp].plane < best) {                           // :157 strict, tested after the near subtree
                    cur = st[sp].node;
                    depth = st[sp].depth;
                    found = true;
                    break;
                }
            }
            if (!found) {                      // this query is answered
                svdb_candidate c;
                if (best_node == NODE_NONE) {
                    c.dist = CUDART_INF;
                    c.seq = SEQ_NONE;
                    c.index = (u64)SVDB_NONE;
                } else {
                    c.dist = best;
                    c.seq = (u64)best_node + seq_base;
                    c.index = log_index[best_node];
                }
                c.flags = overflow ? SVDB_CAND_UNSAFE : 0ull;
                out[qi] = c;
                return;
            }
        }
        // ---- visit `cur` (kdtree.c:131-162) ----
        const double *p = pts + (size_t)cur * stride;
        double d = 0.0;
        double pcd = 0.0;
        const int cd = depth % K;
        for (int i = 0; i < K; i++) {
            const double x = __ldg(p + i);
            if (i == cd) pcd = x;
            const double t = __dsub_rn(x, q[i]);
            d = __dadd_rn(d, __dmul_rn(t, t));                    // kdtree.c:134-137
        }
        if (d < best) {                                          // :139 strict
            best = d;
            best_node = cur;
        }
        const bool left_near = q[cd] < pcd;                      // :147
        const uint32_t lo = child[2 * (size_t)cur], hi = child[2 * (size_t)cur + 1];
        const uint32_t near_c = left_near ? lo : hi, far_c = left_near ? hi : lo;
        if (far_c != NODE_NONE) {
            if (sp < TREE_STACK) {
                const double t = __dsub_rn(q[cd], pcd);
                st[sp].node = far_c;
                st[sp].depth = depth + 1;
                st[sp].plane = __dmul_rn(t, t);                   // :157
                sp++;
            } else {
                overflow = true;
            }
        }
        cur = near_c;
        depth++;
    }
}

// ---- K6 for k > 1: the same traversal keeping the k smallest (distance, seq) keys ----------------
// The reference defines only k = 1.  Position 0 is still ITS answer (first minimum in near-first
// order, strict '<'); the far side is visited whenever the plane is not beyond the current k-th key,
// a superset of what the reference visits, in the same order -- so the first minimum is the same node.
constexpr int KNN_MAX = SVDB_MAX_K;

__global__ void __launch_bounds__(128) tree_knn_kernel(const double *__restrict__ pts, int stride, int K,
                                                       const uint32_t *__restrict__ child, u64 n,
                                                       const double *__restrict__ Q, int ldq, int nq, int k,
                                                       const u64 *__restrict__ log_index, u64 seq_base,
                                                       svdb_candidate *out, const unsigned *__restrict__ only_marked) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    if (only_marked && !only_marked[qi]) return;                 // K9 answered this one already
    double ql[16];
    const double *qg = Q + (size_t)qi * ldq;
    if (K <= 16)
        for (int i = 0; i < K; i++) ql[i] = qg[i];
    const double *q = K <= 16 ? ql : qg;
    struct Frame {
        uint32_t node, depth;
        double plane;
    };
    Frame st[TREE_STACK];
    double kd[KNN_MAX];
    uint32_t ks[KNN_MAX];
    int m = 0, sp = 0;
    bool overflow = false;
    double best = CUDART_INF;
    uint32_t best_node = NODE_NONE;
    uint32_t cur = n ? 0u : NODE_NONE;
    uint32_t depth = 0;
    for (;;) {
        while (cur != NODE_NONE) {
            const double *p = pts + (size_t)cur * stride;
            double d = 0.0, pcd = 0.0;
            const int cd = depth % K;
            for (int i = 0; i < K; i++) {
                const double x = __ldg(p + i);
                if (i == cd) pcd = x;
                const double t = __dsub_rn(x, q[i]);
                d = __dadd_rn(d, __dmul_rn(t, t));
            }
            if (d < best) {
                best = d;
                best_node = cur;
            }
            if (d < CUDART_INF && (m < k || d < kd[m - 1] || (d == kd[m - 1] && cur < ks[m - 1]))) {
                int pos = m < k ? m : k - 1;
                while (pos > 0 && (d < kd[pos - 1] || (d == kd[pos - 1] && cur < ks[pos - 1]))) {
                    kd[pos] = kd[pos - 1];
                    ks[pos] = ks[pos - 1];
                    pos--;
                }
                kd[pos] = d;
                ks[pos] = cur;
                if (m < k) m++;
            }
            const bool left_near = q[cd] < pcd;
            const uint32_t lo = child[2 * (size_t)cur], hi = child[2 * (size_t)cur + 1];
            const uint32_t near_c = left_near ? lo : hi, far_c = left_near ? hi : lo;
            if (far_c != NODE_NONE) {
                if (sp < TREE_STACK) {
                    const double t = __dsub_rn(q[cd], pcd);
                    st[sp].node = far_c;
                    st[sp].depth = depth + 1;
                    st[sp].plane = __dmul_rn(t, t);
                    sp++;
                } else {
                    overflow = true;
                }
            }
            cur = near_c;
            depth++;
        }
        bool found = false;
        while (sp > 0) {
            sp--;
            const double bound = m == k ? kd[k - 1] : CUDART_INF;
            if (st[sp].plane <= bound) {
                cur = st[sp].node;
                depth = st[sp].depth;
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    // winner first, the rest in (distance, seq) order
    svdb_candidate *o = out + (size_t)qi * k;
    const u64 fl = overflow ? SVDB_CAND_UNSAFE : 0ull;
    int w = 0;
    if (best_node != NODE_NONE) {
        o[w].dist = best;
        o[w].seq = (u64)best_node + seq_base;
        o[w].index = log_index[best_node];
        o[w].flags = fl;
        w++;
    }
    for (int i = 0; i < m && w < k; i++) {
        if (ks[i] == best_node) continue;
        o[w].dist = kd[i];
        o[w].seq = (u64)ks[i] + seq_base;
        o[w].index = log_index[ks[i]];
        o[w].flags = fl;
        w++;
    }
    for (; w < k; w++) {
        o[w].dist = CUDART_INF;
        o[w].seq = SEQ_NONE;
        o[w].index = (u64)SVDB_NONE;
        o[w].flags = fl;
    }
}

cudaError_t launch_tree_nearest(const double *pts, int stride, int K, const uint32_t *child, u64 n, const double *Q,
                                int ldq, int nq, int k, const u64 *log_index, u64 seq_base, svdb_candidate *out,
                                cudaStream_t st, const unsigned *only_marked) {
    if (nq == 0) return cudaSuccess;
    if (k == 1) {
        tree_nearest_kernel<<<(nq + 127) / 128, 128, 0, st>>>(pts, stride, K, child, n, Q, ldq, nq, log_index, seq_base, out,
                                                              only_marked);
    } else
        tree_knn_kernel<<<(nq + 127) / 128, 128, 0, st>>>(pts, stride, K, child, n, Q, ldq, nq, k, log_index, seq_base, out,
                                                          only_marked);
    return cudaGetLastError();
}

}  // namespace svdb
