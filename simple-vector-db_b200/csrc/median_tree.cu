// median_tree.cu -- K8 (build) and K9 (traversal) of the BALANCED median KD tree over thin kd-points, sm_100a.
//
// The reference's tree (src/kdtree.c:47-62) is an insertion-order BST: its depth is ~2.5 log2 N for random
// input and N for sorted input, every node is a separate 32-byte fetch, and kdtree_nearest_rec (:131-162)
// walks it one dependent load at a time.  K5/K6 (tree_kernels.cu) keep exactly that shape because the
// reference's answer on exact ties between DISTINCT points depends on it.  Everything else -- every query
// whose minimum is unique, or tied only between copies of one point -- is answered by "the smallest
// (reference-order distance, log sequence number)", which does not depend on the tree shape.  This file
// builds the tree that answers THAT question fastest and flags the queries it may not answer itself:
//
//   * build (K8): median splits by MSD radix-select partitioning, all segments of a level at once.  The tree
//     is implicit: level l has 2^l segments, segment j covers positions [ (j*n)>>l, ((j+1)*n)>>l ) of the
//     permuted point array, node (l, j) has heap id 2^l + j and stores nothing but its split value; leaves
//     are the 2^L segments of <= 32 points at the last level, stored contiguously (coordinates + log seq).
//     Large segments: 8 passes of an 8-bit radix select find the key of median rank in every segment
//     (shared-memory histograms per 2048-element chunk, one pick per segment and pass), then a stable 3-way
//     partition (count, scan, scatter) moves the u32 permutation.  Segments of <= 2048 entries finish ALL
//     their remaining levels inside one CTA: a shared-memory bitonic sort by (sub-segment, key) per level.
//   * traversal (K9): LPQ lanes (32, 16 or 8) per query (k = 1), a full warp per query for k > 1.  The descent reads one split value per level (no
//     point fetch, no child links), a leaf visit is ONE coalesced read of up to 32 points, each lane forms
//     its point's distance in the reference's operation order (kdtree.c:134-137: rounded sub, mul, add, in
//     index order), and three REDUX min-reductions pick the leaf's smallest (distance, seq).  The far side of
//     a split is visited when plane <= best (NOT the reference's strict '<', :157): that is what makes the
//     lowest sequence number among equal distances reachable, and what makes tie detection complete.
//     Pruning is exact in rounded arithmetic: every term of the reference's sum is non-negative and rounding
//     is monotone, so d(p) >= fl(fl(p[cd]-q[cd])^2) >= fl(fl(s-q[cd])^2) for every p beyond the split s.
//   * ties: when two entries at the minimal distance have DIFFERENT coordinates the answer is flagged
//     SVDB_CAND_TIE and the caller reruns exactly those queries through K6, the reference's own traversal
//     (launch_tree_nearest with only_marked).  Copies of one point never flag: the earliest is the
//     reference's answer (the later copy descends the same path and becomes its descendant).
//   * the log is append-only; entries appended after the last build form a tail that the traversal scans
//     after the tree (the engine rebuilds once the tail outgrows "mtree.tail_max").
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace svdb {

constexpr int MT_BUCKET = 32;      // points per leaf (<=)
constexpr int MT_SMALL = 2048;     // segments up to this size finish inside one CTA
constexpr int MT_CHUNK = 2048;     // elements per CTA in the large-segment passes
constexpr int MT_THREADS = 256;
constexpr int MT_PER_THREAD = MT_CHUNK / MT_THREADS;

__host__ __device__ __forceinline__ u64 mt_bound(u64 j, u64 n, int level) { return (j * n) >> level; }

// Where node (level, j) keeps its split value.  blk = 1: heap order (2^level + j).  blk = 3: three levels per 64-byte
// block -- block (t, jt) holds the 7 nodes of the subtree rooted at node (3t, jt) in heap order (slot 7 unused), blocks
// of super-level t start at (8^t - 1) / 7 -- so that a descent pays one memory round trip per THREE levels.
__host__ __device__ __forceinline__ u64 mt_slot(int level, u64 j, int blk) {
    if (blk == 1) return (1ull << level) + j;
    const int t = level / 3, r = level - 3 * t;
    const u64 block = ((1ull << (3 * t)) - 1) / 7 + (j >> r);
    return block * 8 + ((1ull << r) - 1) + (j & ((1ull << r) - 1));
}
__device__ __forceinline__ void mt_prefetch(const void *p) {
#ifndef SVDB_CUSIM
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// double -> u64 whose unsigned order is the numeric order (-0 < +0, which is harmless: a partition
// that is valid in key order is valid in numeric order).  NaNs sort with +inf: their points never win.
__device__ __forceinline__ u64 mt_key(double x) {
    u64 u = (u64)__double_as_longlong(x);
    if (x != x) u = 0x7ff0000000000000ull;
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double mt_unkey(u64 k) {
    const u64 u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// inclusive scan over the 256 threads of a CTA; wt: 8 words of shared memory
__device__ __forceinline__ unsigned mt_block_scan(unsigned v, unsigned *wt) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();            // wt may still be read from a previous call
    if (lane == 31) wt[w] = v;
    __syncthreads();
    unsigned add = 0;
    for (int i = 0; i < w; i++) add += wt[i];
    return v + add;
}

// ---- K8, large segments -------------------------------------------------------------------------------
__global__ void mt_iota_kernel(uint32_t *perm, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) perm[i] = (uint32_t)i;
}

__global__ void mt_keys_kernel(const double *__restrict__ pts, int stride, int cd, const uint32_t *__restrict__ perm,
                               u64 *__restrict__ keys, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        keys[i] = mt_key(__ldg(pts + (size_t)perm[i] * stride + cd));
}

// rank of the median inside each segment; prefix starts empty
__global__ void mt_sel_init_kernel(u64 *prefix, uint32_t *rank, u64 n, int level) {
    const u64 S = 1ull << level;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < S; j += (u64)gridDim.x * blockDim.x) {
        prefix[j] = 0;
        rank[j] = (uint32_t)(mt_bound(2 * j + 1, n, level + 1) - mt_bound(j, n, level));
    }
}

// CTA (segment j, chunk c): histogram of digit `pass` (MSD first) over the keys that match the segment's prefix
__global__ void __launch_bounds__(MT_THREADS) mt_hist_kernel(const u64 *__restrict__ keys, u64 n, int level, unsigned cps,
                                                             int pass, const u64 *__restrict__ prefix, unsigned *hist) {
    __shared__ unsigned sh[256];
    const u64 j = blockIdx.x / cps, c = blockIdx.x % cps;
    const u64 lo = mt_bound(j, n, level), hi = mt_bound(j + 1, n, level);
    const u64 beg = lo + c * MT_CHUNK, end = min(hi, beg + (u64)MT_CHUNK);
    if (beg >= end) return;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const u64 pre = prefix[j];
    for (u64 i = beg + threadIdx.x; i < end; i += MT_THREADS) {
        const u64 k = keys[i];
        if (pass == 0 || (k >> (shift + 8)) == (pre >> (shift + 8))) atomicAdd(&sh[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[j * 256 + threadIdx.x], sh[threadIdx.x]);
}

// one CTA per segment: the digit whose bucket holds the wanted rank extends the prefix
__global__ void __launch_bounds__(MT_THREADS) mt_pick_kernel(unsigned *hist, u64 *prefix, uint32_t *rank, int pass) {
    __shared__ unsigned wt[8];
    const u64 j = blockIdx.x;
    const unsigned c = hist[j * 256 + threadIdx.x];
    hist[j * 256 + threadIdx.x] = 0;                       // ready for the next pass
    const unsigned r = rank[j];
    const unsigned incl = mt_block_scan(c, wt);
    __syncthreads();                                       // everybody has read rank[j]
    if (c && incl - c <= r && r < incl) {
        prefix[j] |= (u64)threadIdx.x << (56 - 8 * pass);
        rank[j] = r - (incl - c);
    }
}

// per chunk: how many keys are below / equal to the segment's median key
__global__ void __launch_bounds__(MT_THREADS) mt_count_kernel(const u64 *__restrict__ keys, u64 n, int level, unsigned cps,
                                                              const u64 *__restrict__ prefix, unsigned *cnt) {
    __shared__ unsigned sh[2];
    const u64 j = blockIdx.x / cps, c = blockIdx.x % cps;
    const u64 lo = mt_bound(j, n, level), hi = mt_bound(j + 1, n, level);
    const u64 beg = lo + c * MT_CHUNK, end = min(hi, beg + (u64)MT_CHUNK);
    if (threadIdx.x < 2) sh[threadIdx.x] = 0;
    __syncthreads();
    const u64 v = prefix[j];
    unsigned nl = 0, ne = 0;
    for (u64 i = beg + threadIdx.x; i < end; i += MT_THREADS) {
        const u64 k = keys[i];
        nl += k < v;
        ne += k == v;
    }
    nl = __reduce_add_sync(FULL, nl);
    ne = __reduce_add_sync(FULL, ne);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh[0], nl);
        atomicAdd(&sh[1], ne);
    }
    __syncthreads();
    if (threadIdx.x < 2) cnt[(u64)blockIdx.x * 2 + threadIdx.x] = sh[threadIdx.x];
}

// one CTA per segment: exclusive scan of the chunk counts, totals, and the node's split value
__global__ void __launch_bounds__(MT_THREADS) mt_scan_kernel(const unsigned *__restrict__ cnt, unsigned *offs, unsigned *tot,
                                                             unsigned cps, int level, const u64 *__restrict__ prefix,
                                                             double *split, int blk) {
    __shared__ unsigned wt[8];
    const u64 j = blockIdx.x;
    unsigned carry_l = 0, carry_e = 0;
    for (unsigned c0 = 0; c0 < cps; c0 += MT_THREADS) {
        const unsigned c = c0 + threadIdx.x;
        const unsigned l = c < cps ? cnt[(j * cps + c) * 2] : 0u;
        const unsigned e = c < cps ? cnt[(j * cps + c) * 2 + 1] : 0u;
        const unsigned il = mt_block_scan(l, wt);
        const unsigned ie = mt_block_scan(e, wt);
        if (c < cps) {
            offs[(j * cps + c) * 2] = carry_l + il - l;
            offs[(j * cps + c) * 2 + 1] = carry_e + ie - e;
        }
        __shared__ unsigned last[2];
        __syncthreads();
        if (threadIdx.x == MT_THREADS - 1) {
            last[0] = il;
            last[1] = ie;
        }
        __syncthreads();
        carry_l += last[0];
        carry_e += last[1];
    }
    if (threadIdx.x == 0) {
        tot[j * 2] = carry_l;
        tot[j * 2 + 1] = carry_e;
        split[mt_slot(level, j, blk)] = mt_unkey(prefix[j]);
    }
}

// stable 3-way partition of every segment around its median key: [less | equal ... equal | greater], with
// exactly rank-of-the-median entries on the left (eqL of the equal ones go left)
__global__ void __launch_bounds__(MT_THREADS) mt_scatter_kernel(const u64 *__restrict__ keys, const uint32_t *__restrict__ perm_in,
                                                                uint32_t *__restrict__ perm_out, u64 n, int level, unsigned cps,
                                                                const u64 *__restrict__ prefix, const uint32_t *__restrict__ rank,
                                                                const unsigned *__restrict__ offs, const unsigned *__restrict__ tot) {
    __shared__ unsigned wt[8];
    const u64 j = blockIdx.x / cps, c = blockIdx.x % cps;
    const u64 lo = mt_bound(j, n, level), hi = mt_bound(j + 1, n, level), mid = mt_bound(2 * j + 1, n, level + 1);
    const u64 beg = lo + c * MT_CHUNK, end = min(hi, beg + (u64)MT_CHUNK);
    const u64 v = prefix[j];
    const unsigned eqL = rank[j], lessT = tot[j * 2], eqT = tot[j * 2 + 1];
    const u64 first = beg + (u64)threadIdx.x * MT_PER_THREAD;
    u64 k[MT_PER_THREAD];
    unsigned nl = 0, ne = 0, mine = 0;
#pragma unroll
    for (int u = 0; u < MT_PER_THREAD; u++) {
        if (first + u < end) {
            k[u] = keys[first + u];
            nl += k[u] < v;
            ne += k[u] == v;
            mine++;
        }
    }
    const unsigned il = mt_block_scan(nl, wt);
    const unsigned ie = mt_block_scan(ne, wt);
    if (mine == 0) return;
    unsigned bl = offs[(u64)blockIdx.x * 2] + il - nl;
    unsigned be = offs[(u64)blockIdx.x * 2 + 1] + ie - ne;
    unsigned bg = (unsigned)(first - lo) - bl - be;
#pragma unroll
    for (int u = 0; u < MT_PER_THREAD; u++) {
        if (first + u < end) {
            u64 dst;
            if (k[u] < v) {
                dst = lo + bl++;
            } else if (k[u] == v) {
                const unsigned e = be++;
                dst = e < eqL ? lo + lessT + e : mid + (e - eqL);
            } else {
                dst = mid + (eqT - eqL) + bg++;
            }
            perm_out[dst] = perm_in[first + u];
        }
    }
}

// ---- K8, small segments: all remaining levels of one segment inside one CTA ---------------------------
__device__ __forceinline__ bool mt_after(u64 ka, u64 va, u64 kb, u64 vb) {     // (sub-segment, key, entry) order
    const uint32_t sa = (uint32_t)(va >> 32), sb = (uint32_t)(vb >> 32);
    if (sa != sb) return sa > sb;
    if (ka != kb) return ka > kb;
    return va > vb;
}

__global__ void __launch_bounds__(MT_THREADS) mt_small_kernel(const double *__restrict__ pts, int stride, int K,
                                                              const uint32_t *__restrict__ perm, u64 n, int level0, int L,
                                                              double *split, int blk, double *__restrict__ mpts,
                                                              uint32_t *__restrict__ mseq) {
    __shared__ u64 skey[MT_SMALL];
    __shared__ u64 sval[MT_SMALL];            // sub-segment id << 32 | log entry
    const u64 j = blockIdx.x;
    const u64 lo = mt_bound(j, n, level0), hi = mt_bound(j + 1, n, level0);
    const unsigned len = (unsigned)(hi - lo);
    for (unsigned p = threadIdx.x; p < MT_SMALL; p += MT_THREADS) sval[p] = p < len ? (u64)perm[lo + p] : ~0ull;
    __syncthreads();
    for (int lev = level0; lev < L; lev++) {
        const int cd = lev % K;
        for (unsigned p = threadIdx.x; p < MT_SMALL; p += MT_THREADS) {
            if (p < len) {
                const uint32_t e = (uint32_t)sval[p];
                const u64 sid = (((lo + p + 1) << lev) - 1) / n;       // segment of level lev that position lo+p lies in
                skey[p] = mt_key(__ldg(pts + (size_t)e * stride + cd));
                sval[p] = (sid << 32) | e;
            } else {
                skey[p] = ~0ull;
                sval[p] = ~0ull;
            }
        }
        __syncthreads();
        for (unsigned size = 2; size <= MT_SMALL; size <<= 1) {
            for (unsigned st = size >> 1; st > 0; st >>= 1) {
                for (unsigned i = threadIdx.x; i < MT_SMALL; i += MT_THREADS) {
                    const unsigned x = i ^ st;
                    if (x > i) {
                        const u64 ka = skey[i], va = sval[i], kb = skey[x], vb = sval[x];
                        const bool up = (i & size) == 0;
                        if (mt_after(ka, va, kb, vb) == up) {
                            skey[i] = kb;
                            sval[i] = vb;
                            skey[x] = ka;
                            sval[x] = va;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // every sub-segment is now sorted on coordinate cd: the entry at its median position is the split
        const u64 nsub = 1ull << (lev - level0), sub0 = j << (lev - level0);
        for (u64 x = threadIdx.x; x < nsub; x += MT_THREADS) {
            const u64 jj = sub0 + x;
            const u64 mpos = mt_bound(2 * jj + 1, n, lev + 1) - lo;
            split[mt_slot(lev, jj, blk)] = mt_unkey(skey[mpos]);
        }
        __syncthreads();
    }
    for (unsigned p = threadIdx.x; p < len; p += MT_THREADS) {
        const uint32_t e = (uint32_t)sval[p];
        mseq[lo + p] = e;
        for (int c = 0; c < K; c++) mpts[(lo + p) * K + c] = pts[(size_t)e * stride + c];
    }
}

int mtree_levels(u64 n) {
    int L = 0;
    while (((n + (1ull << L) - 1) >> L) > (u64)MT_BUCKET) L++;
    return L;
}

size_t mtree_split_count(u64 n, int blk) {
    const int L = mtree_levels(n);
    if (blk == 1 || L == 0) return (size_t)1 << L;
    const int T = (L + 2) / 3;                                   // super-levels; the last one may be partly used
    return (size_t)(((1ull << (3 * T)) - 1) / 7) * 8;
}

#define MT_CK(call)                         \
    do {                                    \
        cudaError_t e__ = (call);           \
        if (e__ != cudaSuccess) {           \
            for (void *p__ : blocks) cudaFree(p__); \
            return e__;                     \
        }                                   \
    } while (0)

cudaError_t launch_mtree_build(const double *pts, int stride, int K, u64 n, int blk, double *split, double *mpts, uint32_t *mseq,
                               int num_sms, cudaStream_t st, int *levels_out, int *launches_out) {
    if (blk != 1 && blk != 3) return cudaErrorInvalidValue;
    std::vector<void *> blocks;
    const int L = mtree_levels(n);
    if (levels_out) *levels_out = L;
    if (launches_out) *launches_out = 0;
    if (n == 0) return cudaSuccess;
    int launches = 0;
    // levels that still have segments larger than one CTA handles
    int levels_a = 0;
    while (levels_a < L && ((n + (1ull << levels_a) - 1) >> levels_a) > (u64)MT_SMALL) levels_a++;
    uint32_t *perm[2] = {nullptr, nullptr};
    u64 *keys = nullptr, *prefix = nullptr;
    unsigned *hist = nullptr, *cnt = nullptr, *offs = nullptr, *tot = nullptr;
    uint32_t *rank = nullptr;
    auto alloc = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
        if (e == cudaSuccess) blocks.push_back(*p);
        return e;
    };
    MT_CK(alloc((void **)&perm[0], n * 4));
    const int grid_all = (int)std::min<u64>((n + 255) / 256, (u64)num_sms * 16);
    mt_iota_kernel<<<grid_all, 256, 0, st>>>(perm[0], n);
    launches++;
    int cur = 0;
    if (levels_a > 0) {
        const u64 s_max = 1ull << (levels_a - 1);
        u64 chunks_max = 0;
        for (int l = 0; l < levels_a; l++) {
            const u64 S = 1ull << l, M = (n + S - 1) >> l;
            chunks_max = std::max(chunks_max, S * ((M + MT_CHUNK - 1) / MT_CHUNK));
        }
        MT_CK(alloc((void **)&perm[1], n * 4));
        MT_CK(alloc((void **)&keys, n * 8));
        MT_CK(alloc((void **)&prefix, s_max * 8));
        MT_CK(alloc((void **)&rank, s_max * 4));
        MT_CK(alloc((void **)&tot, s_max * 8));
        MT_CK(alloc((void **)&hist, s_max * 256 * 4));
        MT_CK(alloc((void **)&cnt, chunks_max * 8));
        MT_CK(alloc((void **)&offs, chunks_max * 8));
        MT_CK(cudaMemsetAsync(hist, 0, s_max * 256 * 4, st));
        for (int l = 0; l < levels_a; l++) {
            const u64 S = 1ull << l, M = (n + S - 1) >> l;
            const unsigned cps = (unsigned)((M + MT_CHUNK - 1) / MT_CHUNK);
            const unsigned grid = (unsigned)(S * cps);
            const int gs = (int)std::min<u64>((S + 255) / 256, (u64)num_sms * 16);
            mt_keys_kernel<<<grid_all, 256, 0, st>>>(pts, stride, l % K, perm[cur], keys, n);
            mt_sel_init_kernel<<<gs, 256, 0, st>>>(prefix, rank, n, l);
            for (int pass = 0; pass < 8; pass++) {
                mt_hist_kernel<<<grid, MT_THREADS, 0, st>>>(keys, n, l, cps, pass, prefix, hist);
                mt_pick_kernel<<<(unsigned)S, MT_THREADS, 0, st>>>(hist, prefix, rank, pass);
            }
            mt_count_kernel<<<grid, MT_THREADS, 0, st>>>(keys, n, l, cps, prefix, cnt);
            mt_scan_kernel<<<(unsigned)S, MT_THREADS, 0, st>>>(cnt, offs, tot, cps, l, prefix, split, blk);
            mt_scatter_kernel<<<grid, MT_THREADS, 0, st>>>(keys, perm[cur], perm[cur ^ 1], n, l, cps, prefix, rank, offs, tot);
            launches += 21;
            cur ^= 1;
            MT_CK(cudaGetLastError());
        }
    }
    mt_small_kernel<<<(unsigned)(1ull << levels_a), MT_THREADS, 0, st>>>(pts, stride, K, perm[cur], n, levels_a, L, split, blk, mpts, mseq);
    launches++;
    MT_CK(cudaGetLastError());
    MT_CK(cudaStreamSynchronize(st));
    for (void *p : blocks) cudaFree(p);
    if (launches_out) *launches_out = launches;
    return cudaSuccess;
}

// ---- K9: traversal ----------------------------------------------------------------------------------
struct MtBest {
    double d;
    uint32_t s;
    const double *p;     // coordinates of the best entry
    bool tie;
};

// One coalesced visit of up to LPQ entries (one per lane of the group): base + (pos0 + lane) * stride.
// KNN (full warps only): every finite candidate is also offered to the warp's sorted list of the klim smallest keys.
template <int LPQ, bool KNN>
__device__ __forceinline__ void mt_visit(const unsigned mask, const int gl, const double *__restrict__ base, int stride, u64 pos0,
                                         unsigned count, const uint32_t *__restrict__ seqs, const double *q, int K, MtBest &b,
                                         WarpList *list = nullptr, int klim = 0) {
    const bool has = (unsigned)gl < count;
    const double *p = base + (pos0 + (has ? gl : 0)) * (size_t)stride;
    double d = 0.0;
    for (int i = 0; i < K; i++) {
        const double t = __dsub_rn(__ldg(p + i), q[i]);
        d = __dadd_rn(d, __dmul_rn(t, t));                         // kdtree.c:134-137
    }
    const uint32_t s = seqs ? __ldg(seqs + pos0 + (has ? gl : 0)) : (uint32_t)(pos0 + gl);
    const bool valid = has && d < CUDART_INF;                      // NaN and +inf never win (kdtree.c:139 against INFINITY)
    if (KNN) list->offer(valid, d, (u64)s, gl, klim);
    const unsigned hi = valid ? (unsigned)__double2hiint(d) : 0xffffffffu;
    const unsigned m_hi = __reduce_min_sync(mask, hi);
    if (m_hi == 0xffffffffu) return;
    const unsigned lo = (valid && hi == m_hi) ? (unsigned)__double2loint(d) : 0xffffffffu;
    const unsigned m_lo = __reduce_min_sync(mask, lo);
    const bool match = valid && hi == m_hi && (unsigned)__double2loint(d) == m_lo;
    const unsigned m_s = __reduce_min_sync(mask, match ? s : 0xffffffffu);
    const double dl = __hiloint2double((int)m_hi, (int)m_lo);
    if (dl > b.d) return;
    const unsigned mm = __ballot_sync(mask, match);
    bool check = false;
    if (dl < b.d) {
        const int src = __ffs(__ballot_sync(mask, match && s == m_s)) - 1;            // absolute lane of the winner
        b.d = dl;
        b.s = m_s;
        b.p = base + (pos0 + (unsigned)(src & (LPQ - 1))) * (size_t)stride;
        b.tie = false;
        check = __popc(mm) > 1;
    } else {                                                       // equals the running best
        check = true;
    }
    if (check) {
        // do the entries at the minimum differ from the best entry's coordinates?  (copies of one point do not)
        bool differs = false;
        if (match)
            for (int i = 0; i < K; i++) differs |= __ldg(p + i) != __ldg(b.p + i);
        if (__ballot_sync(mask, differs)) b.tie = true;
        if (m_s < b.s) {
            const int src = __ffs(__ballot_sync(mask, match && s == m_s)) - 1;
            b.s = m_s;
            b.p = base + (pos0 + (unsigned)(src & (LPQ - 1))) * (size_t)stride;
        }
    }
}

// KNN = false: k = 1.  KNN = true (LPQ = 32): the k smallest (distance, seq) in a warp-distributed list; the far side
// of a split is visited while its plane is not beyond the k-th key; the minimum and its tie flag are tracked as for k = 1.
template <int LPQ, bool KNN>
__global__ void __launch_bounds__(128, KNN ? 8 : 12) mtree_nearest_kernel(const double *__restrict__ split, int blk, const double *__restrict__ mpts,
                                                            const uint32_t *__restrict__ mseq, u64 nb, int L,
                                                            const double *__restrict__ pts, int stride, u64 n, int K,
                                                            const double *__restrict__ Q, int ldq, int nq, int k,
                                                            const u64 *__restrict__ log_index, u64 seq_base, int mark_ties,
                                                            unsigned *__restrict__ marks, svdb_candidate *out) {
    static_assert(!KNN || LPQ == 32, "the k-smallest list is one key per lane of a full warp");
    const int qi = (int)((blockIdx.x * blockDim.x + threadIdx.x) / LPQ);
    if (qi >= nq) return;                                          // whole groups leave together
    const int lane = threadIdx.x & 31, gl = lane & (LPQ - 1);
    const unsigned mask = LPQ == 32 ? FULL : (((1u << LPQ) - 1u) << (lane & ~(LPQ - 1)));
    double q[8];
    for (int i = 0; i < K; i++) q[i] = Q[(size_t)qi * ldq + i];
    MtBest b;
    b.d = CUDART_INF;
    b.s = 0xffffffffu;
    b.p = nullptr;
    b.tie = false;
    WarpList list;
    list.reset();
    if (nb) {
        // the far-side stack is the same in every lane of the group: one copy per group in shared memory (every lane
        // writes the same value and reads back what it wrote; the __syncwarp keeps a lane that runs ahead from
        // pushing over frames a slower lane has not popped yet); thread-local arrays cost 32 copies of the traffic
        __shared__ uint32_t s_node[128 / LPQ][32];
        __shared__ double s_plane[128 / LPQ][32];
        uint32_t *st_node = s_node[threadIdx.x / LPQ];
        double *st_plane = s_plane[threadIdx.x / LPQ];
        int sp = 0;
        uint32_t h = 1;
        for (;;) {
            __syncwarp(mask);
            int lev = 31 - __clz(h);
            int cd = lev % K;
            while (lev < L) {
                const double *sp_ = split + mt_slot(lev, h - (1u << lev), blk);
                if (blk == 3 && lev % 3 == 0) mt_prefetch(sp_ + 4);   // second sector of the block: levels +1/+2 hit L1
                const double s = __ldg(sp_);
                const double qc = q[cd];
                const bool left = qc < s;
                const double t = __dsub_rn(qc, s);
                st_node[sp] = 2 * h + (left ? 1u : 0u);            // the far child
                st_plane[sp] = __dmul_rn(t, t);
                sp++;
                h = 2 * h + (left ? 0u : 1u);
                lev++;
                cd = cd + 1 == K ? 0 : cd + 1;
            }
            const u64 j = h - (1u << L);
            const u64 lo = mt_bound(j, nb, L);
            const unsigned cnt = (unsigned)(mt_bound(j + 1, nb, L) - lo);
            for (unsigned off = 0; off < cnt; off += LPQ)
                mt_visit<LPQ, KNN>(mask, gl, mpts, K, lo + off, min((unsigned)LPQ, cnt - off), mseq, q, K, b, &list, k);
            double bound = b.d;
            if (KNN) {                                             // k-th smallest key so far (+inf while fewer than k)
                u64 ts;
                list.key_at(k - 1, bound, ts);
            }
            bool found = false;
            while (sp > 0) {
                sp--;
                if (st_plane[sp] <= bound) {                       // '<=': equal distances must be seen (lowest seq, tie detection)
                    h = st_node[sp];
                    found = true;
                    break;
                }
            }
            if (!found) break;
        }
    }
    for (u64 pos = nb; pos < n; pos += LPQ)                        // entries appended since the build
        mt_visit<LPQ, KNN>(mask, gl, pts, stride, pos, (unsigned)min((u64)LPQ, n - pos), nullptr, q, K, b, &list, k);
    if (KNN) {
        // lane i holds the i-th smallest (distance, seq); position 0 is the minimum, i.e. the reference's answer unless flagged
        const u64 fl = (b.tie && mark_ties) ? SVDB_CAND_TIE : 0ull;
        if (gl < k) {
            svdb_candidate c;
            const bool none = list.seq == SEQ_NONE;
            c.dist = none ? CUDART_INF : list.d;
            c.seq = none ? SEQ_NONE : list.seq + seq_base;
            c.index = none ? (u64)SVDB_NONE : log_index[list.seq];
            c.flags = none ? 0ull : fl;
            out[(size_t)qi * k + gl] = c;
        }
        if (gl == 0 && marks) marks[qi] = fl ? 1u : 0u;
        return;
    }
    if (gl == 0) {
        svdb_candidate c;
        if (b.s == 0xffffffffu) {
            c.dist = CUDART_INF;
            c.seq = SEQ_NONE;
            c.index = (u64)SVDB_NONE;
            c.flags = 0;
        } else {
            c.dist = b.d;
            c.seq = (u64)b.s + seq_base;
            c.index = log_index[b.s];
            c.flags = (b.tie && mark_ties) ? SVDB_CAND_TIE : 0ull;
        }
        out[qi] = c;
        if (marks) marks[qi] = c.flags ? 1u : 0u;
    }
}

cudaError_t launch_mtree_nearest(const MtreeView &t, const double *pts, int stride, int K, u64 n, const double *Q, int ldq,
                                 int nq, int k, const u64 *log_index, u64 seq_base, int mark_ties, int lanes, unsigned *marks,
                                 svdb_candidate *out, cudaStream_t st) {
    if (nq == 0) return cudaSuccess;
    if (K > 8 || k < 1 || k > 32) return cudaErrorInvalidValue;
    if (k > 1) lanes = 32;
    const int gpb = 128 / lanes;
    const unsigned grid = (unsigned)((nq + gpb - 1) / gpb);
#define MT_ARGS t.split, t.block_levels, t.mpts, t.mseq, t.n_built, t.levels, pts, stride, n, K, Q, ldq, nq, k, log_index, seq_base, mark_ties, marks, out
    if (k > 1)
        mtree_nearest_kernel<32, true><<<grid, 128, 0, st>>>(MT_ARGS);
    else if (lanes == 32)
        mtree_nearest_kernel<32, false><<<grid, 128, 0, st>>>(MT_ARGS);
    else if (lanes == 16)
        mtree_nearest_kernel<16, false><<<grid, 128, 0, st>>>(MT_ARGS);
    else if (lanes == 8)
        mtree_nearest_kernel<8, false><<<grid, 128, 0, st>>>(MT_ARGS);
    else
        return cudaErrorInvalidValue;
#undef MT_ARGS
    return cudaGetLastError();
}

}  // namespace svdb
