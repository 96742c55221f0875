// tail.cuh -- what follows a distance scan, as device functions shared by the stand-alone kernels and by the scans'
// fused tail (the last CTA of a scan launch runs them itself, kernels.h: TailArgs):
//   finalize_query   per query: merge the per-CTA lists, recompute the survivors in the reference's operation order
//                    (kdtree.c:134-137), rank by (d, seq), emit top-k, prove that nothing outside the candidate set
//                    can belong to it, resolve exact ties the way the reference's tree does (kdtree.c:139, :147-159)
//   xch_push / xch_wait / merge_gathered
//                    the cross-shard exchange over NVLink peer memory (exchange.cu) and the K7 merge
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "tree.cuh"

namespace svdb {

constexpr size_t XCH_FLAG_BYTES = 32 * 8;
constexpr int XCH_EPOCH_SLOT = 31;

__device__ __forceinline__ void st_release_sys_u64(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_sys_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// lists are written by other CTAs of the same launch (fused tail): read them past L1
__device__ __forceinline__ Cand ldcg_cand(const Cand *p) {
    Cand c;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<u64 *>(&c.d)), "=l"(c.seq) : "l"(p));
    return c;
}

// shared memory finalize_query needs besides the staging / re-rank buffer: [nw][32] Cand, [nw] bounds, FIN_NC candidate
// entries + 8 words, FIN_NC candidates, and the mbarrier its bulk copies complete on (the caller initialises it: fin_bar_init)
// (every part a multiple of 16 bytes: the buffer behind it is the destination of bulk copies)
constexpr int FIN_NC = 128;                     // candidate slots of the selection path (one thread re-ranks one candidate)
__host__ __device__ constexpr size_t fin_head_bytes(int nw) {
    return (size_t)nw * 32 * 16 + (size_t)((nw + 1) & ~1) * 8 + (size_t)(FIN_NC + 8) * 8 + (size_t)FIN_NC * 16 + 16;
}
__device__ __forceinline__ uint32_t fin_bar_addr(unsigned char *fsm, int nw) { return smem_u32(fsm + fin_head_bytes(nw) - 16); }
// one thread, followed by a __syncthreads() before the first finalize_query
__device__ __forceinline__ void fin_bar_init(unsigned char *fsm, int nw) {
    mbar_init(fin_bar_addr(fsm, nw), 1);
    mbar_fence_init();
}
constexpr int FIN_MIN_TBUF = 32 * 33 * 8;      // room for 32 candidates x 32 coordinates per round

// =====================================================================================
// finalize of ONE query by the whole CTA (blockDim.x = nw * 32, nw <= 16).  fsm: fsm_bytes of shared memory the
// caller does not need any more (>= fin_head_bytes(nw) + FIN_MIN_TBUF, 128-byte aligned, its mbarrier initialised by
// fin_bar_init; `phase` = that barrier's phase, 0 at first).  Ends with a __syncthreads().
//   1. the candidates: either (short lists of approximate keys that fit shared memory together -- the single-plane scans)
//      SELECTED from all lists at once: an upper bound of the k-th smallest key from one pass, then every entry inside the
//      error window of it, up to FIN_NC of them; or (long lists, exact keys) MERGED: every warp merges a slice of the
//      per-CTA lists into its register list, the slices are merged through shared memory -> the 32 best approximate keys.
//      Either way `bound` is the smallest approximate key any list may have dropped;
//   2. only candidates whose approximate key is within the error margin of the k-th can belong to the exact top-k;
//      those are recomputed in the reference's operation order: the warps fetch each candidate row coalesced and form
//      the rounded squares t_i = (x_i - q_i)^2 in shared memory, then one thread per candidate adds them strictly in
//      index order (the serial chain the reference has); with more than 32 candidates the best 32 by exact distance go on;
//   3. rank by (exact distance, seq), emit top-k, prove completeness against `bound`.
// =====================================================================================
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

static __device__ __noinline__ void finalize_query(const FinalArgs &p, int qi, unsigned char *fsm, int fsm_bytes, uint32_t &phase,
                                                   unsigned long long *dbg = nullptr) {
    const int NW = blockDim.x >> 5;
    Cand *mrg = reinterpret_cast<Cand *>(fsm);                                        // [NW][32]
    double *wbound = reinterpret_cast<double *>(fsm + (size_t)NW * 32 * sizeof(Cand)); // [NW]
    u64 *cseq = reinterpret_cast<u64 *>(wbound + ((NW + 1) & ~1));                    // [FIN_NC] entries, [FIN_NC] counter, [FIN_NC + 1] k-th key
    double *red = reinterpret_cast<double *>(cseq + FIN_NC + 2);                      // [4] block reductions
    Cand *cand = reinterpret_cast<Cand *>(cseq + FIN_NC + 8);                         // [FIN_NC] (selection path)
    double *tbuf = reinterpret_cast<double *>(fsm + fin_head_bytes(NW));
    const int tcap = (fsm_bytes - (int)fin_head_bytes(NW)) / 8;                       // doubles in tbuf
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Cand *L = p.lists + (size_t)qi * p.nlists * p.cap;
    const double *qv = p.q + (size_t)qi * p.ldq;

    const bool approx = p.sq_mode || p.eps >= 0.0;
    const double eps64 = 4.0 * (double)(p.K + 2) * 1.1102230246251565e-16;    // reference-order sum vs the real-number sum
    // the store-wide terms of the error bounds: fetched now, needed once the k-th smallest key is known
    const unsigned long long plane_err_b = p.plane_err_bits ? __ldg(p.plane_err_bits) : 0ull;
    const unsigned long long xn_max_b = p.xn_max_bits ? __ldg(p.xn_max_bits) : 0ull;
    const double qnorm_in = (!p.sq_mode && p.eabs_coef > 0.0) ? __ldg(p.qnorm + qi) : 0.0;
    const double p8_lo = p.sq_mode == 2 ? __ldg(&p.plane8->lo) : 0.0, p8_step = p.sq_mode == 2 ? __ldg(&p.plane8->step) : 1.0;
    const uint32_t bar = fin_bar_addr(fsm, NW);
    const Cand *sl = reinterpret_cast<const Cand *>(tbuf);
    const int T = blockDim.x;
    double bound = CUDART_INF;                           // the smallest approximate key any scan list may have dropped
    double eq2 = 0.0, qn2 = 0.0, eabs = 0.0, E = 0.0, scale = 0.0;
    double cd = CUDART_INF;                              // thread j < nneed: approximate key and entry of candidate j
    u64 cs = SEQ_NONE;
    int nneed = 0;
    bool overflow = false;
    double d_next = CUDART_INF;                          // warp 0: exact distance of the best candidate that got no lane (rank 32)

    // |key - d| bounds -> the largest approximate key a member of the true top-k can have, given the k-th smallest key dk
    auto window = [&](double dk) -> double {
        if (p.sq_mode) {
            // the k rows with the smallest keys have sqrt(d) <= U; a row with sqrt(d) <= U has sqrt(key) <= (U + E)(1 + gamma)
            const double U = sqrt(dk) / (1.0 - p.sq_gamma) + E;
            const double lim = (U * (1.0 + eps64) + E) * (1.0 + p.sq_gamma);
            return lim * lim * (1.0 + 1e-12);
        }
        // |key - d| <= eps d + eabs: a member of the true top-k has key <= (dk + eabs)(1 + eps)/(1 - eps) + eabs
        return (dk + eabs) * ((1.0 + p.eps) / (1.0 - p.eps)) * (1.0 + 1e-15) + eabs;
    };
    // sqrt-form keys: |q - fl32(q)|^2 and |q|^2 of this query are formed here (no prep launch in front of the scan), while
    // the first copy of lists is in flight; eight loads in flight per thread.  Per-warp sums.
    auto query_norms = [&]() {
        for (int i0 = 0; i0 < p.K; i0 += 8 * T) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * T + (int)threadIdx.x;
                v[u] = i < p.K ? __ldg(qv + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                // q^: K12 scans against fl32(q); K13 against q on the byte plane's query grid (the scan's own expression)
                const double qh = p.sq_mode == 2 ? p8_lo + p8_step * ((double)p8_quant_q(v[u], p8_lo, p8_step) / 256.0)
                                                 : (double)__double2float_rn(v[u]);
                const double r = i0 + u * T + (int)threadIdx.x < p.K ? v[u] - qh : 0.0;     // slots beyond K are not coordinates
                eq2 = fma(r, r, eq2);
                qn2 = fma(v[u], v[u], qn2);
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            eq2 += shfl_xor_f64(eq2, m);
            qn2 += shfl_xor_f64(qn2, m);
        }
    };
    auto error_terms = [&]() {                           // after red[] is complete
        if (p.sq_mode) {
            // E bounds |x - x^| + |q - q^| (2-norms) plus what squares of tiny differences lose to fp32 underflow
            E = __longlong_as_double((long long)plane_err_b) + sqrt(red[0]) * (1.0 + 1e-9) + sqrt((double)p.K) * 1e-22;
            scale = __longlong_as_double((long long)xn_max_b) + red[1];
            // K13: x^ and q^ are real numbers lo + step * (integer); their fp64 evaluations (plane build, above) are off by
            // 2^-52 of their size -- far below anything the window notices, but it belongs to E
            if (p.sq_mode == 2) E += 1e-14 * sqrt(scale);
        } else if (p.eabs_coef > 0.0) {
            // GEMM-form keys (K2, K10) carry an absolute error
            scale = __longlong_as_double((long long)xn_max_b) + qnorm_in;
            eabs = p.eabs_coef * scale;
        }
    };

    const int total_c = p.nlists * p.cap;
    const bool select_path = approx && p.cap < 16 && total_c > 0 && (size_t)total_c * sizeof(Cand) <= (size_t)tcap * 8;
    if (select_path) {
        // ---- 1+2 (approximate keys, short lists, all of them fit the buffer -- the single-query tail): SELECTION, not a merge.
        // One bulk async copy brings every list into shared memory.  ONE strided pass gives every thread the smallest key
        // among its entries; the k-th smallest of those T minima (rank counting in shared memory) is an upper bound of the
        // k-th smallest key overall -- they are k distinct entries -- and for k = 1 it is the minimum itself.  Any upper
        // bound serves: the window only has to contain the true top-k.  The candidates are the entries <= window(that key),
        // compacted with one ballot per 32 entries into up to FIN_NC slots.  (Measured before: walking the 296 lists of a
        // single-query scan by warp-list insertion cost 11-18 of the tail's 27 us; k rounds of min-extraction 2.6 us each.)
        if (threadIdx.x == 0) {
            // the lists were written through the generic proxy (by other SMs), the bulk copy reads through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
            const uint32_t bytes = (uint32_t)total_c * sizeof(Cand);
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(tbuf), L, bytes, bar);
            cseq[FIN_NC] = 0;                            // candidate counter
            cseq[FIN_NC + 1] = ~0ull;                    // k-th smallest thread minimum (ord), none yet
        }
        if (p.sq_mode) query_norms();
        mbar_wait(bar, phase);
        phase ^= 1;
        if (dbg && threadIdx.x == 0) dbg[10] = global_timer_ns();
        // keys are compared as integers: ord() maps a double's bits to an unsigned with the same order (negative keys
        // happen with GEMM-form keys), so the pass costs integer compares only -- this CTA runs alone on its SM with one warp
        // per scheduler, every instruction's latency shows
        auto ord = [](u64 bits) -> u64 { return bits ^ ((u64)((long long)bits >> 63) | 0x8000000000000000ull); };
        const ulonglong2 *se = reinterpret_cast<const ulonglong2 *>(sl);       // .x = key bits, .y = entry
        u64 *smin = reinterpret_cast<u64 *>(mrg);                              // [T] thread minima
        u64 bo = ~0ull;
        {
            int slot = (int)threadIdx.x % p.cap;
            const int slot_step = T % p.cap;
#pragma unroll 4
            for (int i = threadIdx.x; i < total_c; i += T) {
                const ulonglong2 c = se[i];
                const u64 o = ord(c.x);
                const bool valid = c.y != SEQ_NONE;
                if (valid && slot == p.cap - 1) bound = fmin(bound, __longlong_as_double((long long)c.x));   // that list was full
                if (valid && o < bo) bo = o;
                slot += slot_step;
                if (slot >= p.cap) slot -= p.cap;
            }
        }
        smin[threadIdx.x] = bo;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) bound = fmin(bound, shfl_xor_f64(bound, m));
        if (lane == 0) {
            wbound[warp] = bound;
            cand[warp] = Cand{eq2, (u64)__double_as_longlong(qn2)};
        }
        __syncthreads();
        {
            int rk = 0;
            for (int j = 0; j < T; j++) {
                const u64 o = smin[j];
                rk += (o < bo || (o == bo && j < (int)threadIdx.x)) ? 1 : 0;
            }
            if (rk == p.k - 1) cseq[FIN_NC + 1] = bo;    // exactly one thread
            double a = 0.0, b2 = 0.0;
            bound = CUDART_INF;
            for (int w = 0; w < NW; w++) {
                bound = fmin(bound, wbound[w]);
                a += cand[w].d;
                b2 += __longlong_as_double((long long)cand[w].seq);
            }
            if (threadIdx.x == 0) {
                red[0] = a;
                red[1] = b2;
            }
        }
        __syncthreads();
        const u64 ko = cseq[FIN_NC + 1];
        const u64 sk = ko != ~0ull ? 0ull : SEQ_NONE;    // fewer than k threads saw an entry: every entry is a candidate
        // the k-th smallest key as a distance again (ord is an involution up to the sign test)
        const u64 kbits = (ko >> 63) ? (ko ^ 0x8000000000000000ull) : ~ko;
        const double dk = __longlong_as_double((long long)kbits);
        if (dbg && threadIdx.x == 0) dbg[11] = global_timer_ns();
        error_terms();
        const double lim = sk != SEQ_NONE ? window(dk) : CUDART_INF;
        const unsigned below = (1u << lane) - 1u;
        // one thread re-ranks one candidate; the re-rank buffer must hold at least 32 coordinates of each per round
        const int nc_max = min(min(FIN_NC, T), tcap / 33);
        for (int base = warp * 32; base < total_c; base += NW * 32) {
            const int i = base + lane;
            Cand c = Cand{CUDART_INF, SEQ_NONE};
            if (i < total_c) c = sl[i];
            const bool in = c.seq != SEQ_NONE && !(c.d > lim);
            const unsigned m = __ballot_sync(FULL, in);
            if (m) {
                unsigned at = 0;
                if (lane == 0) at = (unsigned)atomicAdd(reinterpret_cast<unsigned long long *>(cseq + FIN_NC), (unsigned long long)__popc(m));
                at = __shfl_sync(FULL, at, 0) + __popc(m & below);
                if (in && at < (unsigned)nc_max) {
                    cand[at] = c;
                    cseq[at] = c.seq;
                }
            }
        }
        __syncthreads();
        if (dbg && threadIdx.x == 0) {                   // diagnostics: what the selection saw
            dbg[16] = (unsigned long long)__double_as_longlong(E);
            dbg[17] = (unsigned long long)__double_as_longlong(lim);
            dbg[18] = (unsigned long long)__double_as_longlong(dk);
            dbg[19] = cseq[FIN_NC];
            dbg[20] = (unsigned long long)__double_as_longlong(bound);
        }
        const unsigned found = (unsigned)cseq[FIN_NC];
        overflow = found > (unsigned)nc_max;             // more near-ties than slots: not provable here (-> fp64 / exact rerun)
        nneed = (int)min(found, (unsigned)nc_max);
        if ((int)threadIdx.x < nneed) {
            cd = cand[threadIdx.x].d;
            cs = cand[threadIdx.x].seq;
        }
        if (dbg && threadIdx.x == 0) dbg[1] = dbg[12] = dbg[13] = global_timer_ns();
    } else {
        // ---- 1. merge (long lists, exact keys, or lists larger than the buffer) ----
        // The lists come into shared memory with one bulk async copy per chunk of whole lists that fits the buffer; every
        // warp merges a slice into its register list, the slices are merged through shared memory -> the 32 best
        // approximate keys, plus `bound`.
        WarpList wl;
        wl.reset();
        const int per_chunk = max(1, (int)(((size_t)tcap * 8) / ((size_t)p.cap * sizeof(Cand))));
        for (int l0 = 0; l0 < p.nlists; l0 += per_chunk) {
            const int nl = min(per_chunk, p.nlists - l0);
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                const uint32_t bytes = (uint32_t)nl * p.cap * sizeof(Cand);
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(smem_u32(tbuf), L + (size_t)l0 * p.cap, bytes, bar);
            }
            if (l0 == 0 && p.sq_mode) query_norms();
            mbar_wait(bar, phase);
            phase ^= 1;
            if (dbg && threadIdx.x == 0 && l0 == 0) dbg[10] = global_timer_ns();
            if (p.cap >= 16) {
                // long lists (k >= 8): most keys of a list qualify while the running list fills up, and every one of them
                // would be a serial insert -- merge list by list with the fixed-cost bitonic network instead
                for (int l = warp; l < nl; l += NW) {
                    Cand c = Cand{CUDART_INF, SEQ_NONE};
                    if (lane < p.cap) c = sl[(size_t)l * p.cap + lane];
                    if (lane == p.cap - 1 && c.seq != SEQ_NONE) bound = fmin(bound, c.d);   // that list was full
                    wl.merge_sorted(c.d, c.seq, lane);
                }
            } else {
                const int tc = nl * p.cap;
                for (int base = warp * 32; base < tc; base += NW * 32) {
                    const int i = base + lane;
                    Cand c = Cand{CUDART_INF, SEQ_NONE};
                    if (i < tc) c = sl[i];
                    const bool has = c.seq != SEQ_NONE;
                    if (has && (i % p.cap) == p.cap - 1) bound = fmin(bound, c.d);       // that list was full
                    wl.offer(has, c.d, c.seq, lane, p.cap);
                }
            }
            __syncthreads();                             // the buffer is refilled (or reused by the re-rank) next
        }
        if (dbg && threadIdx.x == 0) dbg[11] = global_timer_ns();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) bound = fmin(bound, shfl_xor_f64(bound, m));
        mrg[warp * 32 + lane] = Cand{wl.d, wl.seq};
        if (lane == 0) {
            wbound[warp] = bound;
            if (p.sq_mode) {
                tbuf[2 * warp] = eq2;
                tbuf[2 * warp + 1] = qn2;
            }
        }
        __syncthreads();
        if (dbg && threadIdx.x == 0) dbg[12] = global_timer_ns();
        if (p.sq_mode && threadIdx.x == 0) {
            double a = 0.0, b2 = 0.0;
            for (int w = 0; w < NW; w++) {
                a += tbuf[2 * w];
                b2 += tbuf[2 * w + 1];
            }
            red[0] = a;
            red[1] = b2;
        }
        if (warp == 0) {
            // only the best `cap` keys of a slice list are meaningful; the slice lists are sorted: bitonic merges
            if (lane >= p.cap) wl.reset();
            for (int w = 1; w < NW; w++) {
                Cand c = mrg[w * 32 + lane];
                // a slice list that is full may itself have dropped keys >= its last one
                if (lane == p.cap - 1 && c.seq != SEQ_NONE) bound = fmin(bound, c.d);
                if (lane >= p.cap) c = Cand{CUDART_INF, SEQ_NONE};
                wl.merge_sorted(c.d, c.seq, lane);
                bound = fmin(bound, wbound[w]);
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) bound = fmin(bound, shfl_xor_f64(bound, m));
            double dl;
            u64 sl2;
            wl.key_at(p.cap - 1, dl, sl2);
            if (sl2 != SEQ_NONE) bound = fmin(bound, dl);     // the final list is full too
        }
        __syncthreads();                                     // red[] is complete; tbuf may be reused
        if (dbg && threadIdx.x == 0) dbg[1] = global_timer_ns();

        // ---- 2. which candidates can still belong to the exact top-k ----
        error_terms();
        if (warp == 0) {
            const bool valid = wl.seq != SEQ_NONE && lane < p.cap;
            bool need = valid;
            double dk;
            u64 sk;
            wl.key_at(p.k - 1, dk, sk);
            if (approx && sk != SEQ_NONE) need = valid && !(wl.d > window(dk));
            nneed = __popc(__ballot_sync(FULL, need));      // the list is sorted: a prefix of the lanes
            cd = wl.d;
            cs = wl.seq;
            cseq[lane] = wl.seq;
            if (lane == 0) cseq[FIN_NC] = (u64)nneed;
        }
        __syncthreads();
        nneed = (int)cseq[FIN_NC];
        if (dbg && threadIdx.x == 0) dbg[13] = global_timer_ns();
    }

    double dex = CUDART_INF;
    if (!approx) {
        dex = cd;                                       // keys are reference-order already
    } else {
        dex = 0.0;
        // as many coordinates per round as the buffer holds for `nneed` candidates (usually all K)
        int ch = nneed > 0 ? (tcap / nneed - 1) & ~31 : 256;
        if (ch > p.K) ch = (p.K + 31) & ~31;
        const int ld = ch + 1;                          // odd stride: conflict-free column walks
        for (int c0 = 0; c0 < p.K; c0 += ch) {
            const int len = min(ch, p.K - c0);
            // work items = (candidate, 256-coordinate segment), dealt round-robin to the warps; every
            // lane keeps 8 row loads and 8 query loads in flight
            const int nseg = (len + 255) >> 8;
            // two items per trip: 16 row loads + 16 query loads in flight per lane (the loop is DRAM-latency bound)
            for (int w = warp; w < nneed * nseg; w += 2 * NW) {
                double x[2][8], y[2][8];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int wi = w + h * NW;
                    const bool on = wi < nneed * nseg;
                    const int j = on ? wi / nseg : 0, s0 = on ? (wi % nseg) << 8 : 0;
                    const double *row = p.pts + cseq[j] * (u64)p.stride + c0 + s0;
                    const double *qq = qv + c0 + s0;
                    const int slen = on ? min(256, len - s0) : 0;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = u * 32 + lane;
                        x[h][u] = i < slen ? __ldg(row + i) : 0.0;
                        y[h][u] = i < slen ? __ldg(qq + i) : 0.0;
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int wi = w + h * NW;
                    if (wi >= nneed * nseg) break;
                    const int j = wi / nseg, s0 = (wi % nseg) << 8;
                    double *t = tbuf + j * ld + s0;
                    const int slen = min(256, len - s0);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = u * 32 + lane;
                        const double df = __dsub_rn(x[h][u], y[h][u]);
                        if (i < slen) t[i] = __dmul_rn(df, df);
                    }
                }
            }
            __syncthreads();
            if (dbg && threadIdx.x == 0 && c0 == 0) dbg[14] = global_timer_ns();
            if ((int)threadIdx.x < nneed) {
                // kdtree.c:136, strictly in index order: a chain of `len` dependent rounded adds (the one part of the
                // reference's loop that cannot be parallelised).  Sixteen squares are fetched ahead of the chain.
                const uint32_t ta = smem_u32(tbuf + threadIdx.x * ld);
                int i = 0;
                for (; i + 16 <= len; i += 16) {
                    double v[16];
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(ta + (uint32_t)(i + u) * 8u));
#pragma unroll
                    for (int u = 0; u < 16; u++) dex = __dadd_rn(dex, v[u]);
                }
                for (; i < len; i++) dex = __dadd_rn(dex, tbuf[threadIdx.x * ld + i]);
            }
            __syncthreads();
        }
        if (nneed > 32) {
            // more candidates than warp 0 has lanes (selection path, wide windows): rank all of them by (exact distance,
            // entry) and hand the best 32 to warp 0 -- k <= 24 of them are the answer.  The 33rd is kept for the tie test.
            Cand *all = reinterpret_cast<Cand *>(tbuf), *best = all + FIN_NC;       // [nneed], [33]
            Cand mine = Cand{CUDART_INF, SEQ_NONE};
            if ((int)threadIdx.x < nneed && cs != SEQ_NONE && dex < CUDART_INF) mine = Cand{dex, cs};   // kdtree.c:139: non-finite never wins
            if ((int)threadIdx.x < nneed) all[threadIdx.x] = mine;
            __syncthreads();
            if ((int)threadIdx.x < nneed) {
                int rk = 0;
                for (int j = 0; j < nneed; j++) {
                    const Cand c = all[j];
                    rk += (key_less(c.d, c.seq, mine.d, mine.seq) || (c.d == mine.d && c.seq == mine.seq && j < (int)threadIdx.x)) ? 1 : 0;
                }
                if (rk <= 32) best[rk] = mine;
            }
            __syncthreads();
            if (warp == 0) {
                dex = best[lane].d;
                cs = best[lane].seq;
                d_next = best[32].d;
            }
            nneed = 32;
            __syncthreads();
        }
    }

    if (warp == 0) {
        // ---- 3. rank and emit ----
        bool valid = lane < nneed && cs != SEQ_NONE;
        u64 seq = cs;
        if (valid && !(dex < CUDART_INF)) valid = false;    // kdtree.c:139 strict <: non-finite never wins
        if (!valid) {
            dex = CUDART_INF;
            seq = SEQ_NONE;
        }
        int rank = 0;
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const double dj = __shfl_sync(FULL, dex, j);
            const u64 sj = __shfl_sync(FULL, seq, j);
            rank += key_less(dj, sj, dex, seq) ? 1 : 0;
        }
        const int nvalid = __popc(__ballot_sync(FULL, valid));
        const unsigned mk = __ballot_sync(FULL, valid && rank == p.k - 1);
        const double ek = mk ? __shfl_sync(FULL, dex, __ffs(mk) - 1) : CUDART_INF;

        bool unsafe = overflow;                          // more entries inside the window than candidate slots: not provable here
        {
            // an exact tie at the minimum that reaches beyond the 32 candidates warp 0 holds cannot be resolved here
            const unsigned m0 = __ballot_sync(FULL, valid && rank == 0);
            if (m0 && d_next < CUDART_INF && d_next == __shfl_sync(FULL, dex, __ffs(m0) - 1)) unsafe = true;
        }
        if (approx && bound < CUDART_INF) {
            if (p.sq_mode) {
                // entries outside the candidate set have key >= bound, hence sqrt(d) >= sqrt(bound)/(1 + gamma) - E
                const double lb = sqrt(bound) / (1.0 + p.sq_gamma) - E;
                unsafe = unsafe || nvalid < p.k || !(lb > 0.0 && ek < lb * lb * (1.0 - eps64 - 1e-12));
            } else {
                // entries outside the candidate set have approximate key >= bound, hence reference
                // distance >= bound * (1 - eps); they cannot enter the top-k iff ek is strictly below
                unsafe = unsafe || nvalid < p.k || !(ek < bound * (1.0 - p.eps) - eabs);
            }
        }
        if (p.scale_hi > 0.0) {
            // fp32 keys (K10, K11, K12): their error bound only holds while neither squares overflow nor products underflow
            if (!(scale >= p.scale_lo && scale <= p.scale_hi)) unsafe = true;
        }

        // ---- exact ties at the minimum: the reference keeps whichever its tree reaches first ----
        bool tie_flag = false;
        if ((p.child != nullptr || p.mark_ties) && nvalid >= 2) {
            const unsigned m0 = __ballot_sync(FULL, valid && rank == 0);
            const int l0 = __ffs(m0) - 1;
            const double e1 = __shfl_sync(FULL, dex, l0);
            const u64 seq0 = __shfl_sync(FULL, seq, l0);
            const bool tied = valid && dex == e1;
            unsigned tmask = __ballot_sync(FULL, tied);
            const int nt = __popc(tmask);
            if (nt >= 2) {
                // a dropped entry could tie as well: exact keys -> bound <= e1; approximate keys are
                // already covered by the completeness proof above (e1 <= ek < what a dropped entry can have)
                const bool more = !approx && bound <= e1;
                if (more && p.child != nullptr) unsafe = true;
                // identical kd-points? then the earliest insert is an ancestor of the others and wins
                bool differs = false;
                const double *r0 = p.pts + seq0 * (u64)p.stride;
                for (unsigned tm = tmask; tm; tm &= tm - 1) {
                    const u64 st = __shfl_sync(FULL, seq, __ffs(tm) - 1);
                    const double *rt = p.pts + st * (u64)p.stride;
                    for (int i = lane; i < p.K; i += 32) differs |= rt[i] != r0[i];
                }
                differs = __any_sync(FULL, differs);
                if (p.child == nullptr) {
                    // one shard of a larger log: the order of the GLOBAL tree decides (tie_protocol.cu); say so
                    tie_flag = differs || more;
                } else if (differs) {
                    if (tied) cseq[rank] = seq;            // tied entries hold ranks 0..nt-1
                    __syncwarp();
                    u64 w = 0;
                    if (lane == 0) w = resolve_tie(p.pts, p.stride, p.K, p.child, qv, cseq, nt);
                    w = __shfl_sync(FULL, w, 0);
                    const unsigned mw = __ballot_sync(FULL, tied && seq == w);
                    const int wr = __shfl_sync(FULL, rank, __ffs(mw) - 1);
                    if (tied) {
                        if (seq == w) rank = 0;
                        else if (rank < wr) rank++;
                    }
                }
            }
        }
        const u64 oflags = (unsafe ? SVDB_CAND_UNSAFE : 0ull) | (tie_flag ? SVDB_CAND_TIE : 0ull);
        svdb_candidate *out = p.out + (size_t)qi * p.k;
        if (valid && rank < p.k) {
            svdb_candidate c;
            c.dist = dex;
            c.seq = seq + p.seq_base;
            c.index = p.log_index[seq];
            c.flags = oflags;
            out[rank] = c;
        }
        if (lane < p.k && lane >= nvalid) {
            svdb_candidate c;
            c.dist = CUDART_INF;
            c.seq = SEQ_NONE;
            c.index = (u64)SVDB_NONE;
            c.flags = oflags;
            out[lane] = c;
        }
    }
    __syncthreads();
}

// =====================================================================================
// Cross-shard exchange (exchange.cu describes the buffers).  All by one CTA.
// =====================================================================================
// local results (nrec candidates) -> slot `rank` of every peer's gather buffer, then the epoch flags.  Returns the epoch.
__device__ __forceinline__ u64 xch_push(const svdb_candidate *local, int nrec, const PeerPtrs &peers, int rank, int world,
                                        size_t max_rec, u64 *s_epoch) {
    if (threadIdx.x == 0) {
        u64 *mine = reinterpret_cast<u64 *>(peers.p[rank]);
        *s_epoch = mine[XCH_EPOCH_SLOT] + 1;
        mine[XCH_EPOCH_SLOT] = *s_epoch;              // read back by whoever waits for this epoch
    }
    __syncthreads();
    const u64 epoch = *s_epoch;
    const size_t parity_off = XCH_FLAG_BYTES + (size_t)(epoch & 1) * world * max_rec * sizeof(svdb_candidate);
    const int words = nrec * 4;                                    // 8-byte words
    const u64 *src = reinterpret_cast<const u64 *>(local);
    for (int r = 0; r < world; r++) {
        u64 *dst = reinterpret_cast<u64 *>(peers.p[r] + parity_off + (size_t)rank * max_rec * sizeof(svdb_candidate));
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys_u64(reinterpret_cast<u64 *>(peers.p[threadIdx.x]) + rank, epoch);
    return epoch;
}
// lanes 0..world-1 of the calling warp spin until every shard's data of `epoch` has landed in the local buffer
__device__ __forceinline__ void xch_wait(const unsigned char *mine, int world, u64 epoch, int lane) {
    const u64 *flags = reinterpret_cast<const u64 *>(mine);
    if (lane < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(flags + lane) < epoch) {
            if (clock64() - t0 > 20000000000ll) __trap();          // a peer died: do not hang the GPU
        }
    }
    __syncwarp();
}
// K7 for one query by one warp: in = the gathered blocks of this epoch ([world] blocks, max_rec records apart)
__device__ __forceinline__ void merge_gathered(const svdb_candidate *in, int world, size_t max_rec, int qi, int k, int lane,
                                               svdb_candidate *out) {
    WarpList wl;
    wl.reset();
    u64 fl = 0;
    const int total = world * k;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        double d = CUDART_INF;
        u64 s = SEQ_NONE;
        if (i < total) {
            const svdb_candidate *c = in + (size_t)(i / k) * max_rec + (size_t)qi * k + (i % k);
            d = __ldcg(&c->dist);
            s = __ldcg(&c->seq);
            fl |= __ldcg(&c->flags) & ~SVDB_CAND_TIE;
        }
        wl.offer(s != SEQ_NONE, d, s, lane);
    }
    // SVDB_CAND_TIE of the merged answer: >= 2 entries at the merged minimum, or one that its shard flagged
    double dmin;
    u64 smin;
    wl.key_at(0, dmin, smin);
    int at_min = 0;
    if (smin != SEQ_NONE) {
        for (int base = 0; base < total; base += 32) {
            const int i = base + lane;
            if (i < total) {
                const svdb_candidate *c = in + (size_t)(i / k) * max_rec + (size_t)qi * k + (i % k);
                if (__ldcg(&c->seq) != SEQ_NONE && __ldcg(&c->dist) == dmin) at_min += (__ldcg(&c->flags) & SVDB_CAND_TIE) ? 2 : 1;
            }
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        fl |= __shfl_xor_sync(FULL, fl, m);
        at_min += __shfl_xor_sync(FULL, at_min, m);
    }
    if (at_min >= 2) fl |= SVDB_CAND_TIE;
    if (lane < k) {
        svdb_candidate c;
        c.dist = wl.d;
        c.seq = wl.seq;
        c.index = (u64)SVDB_NONE;
        c.flags = fl;
        if (wl.seq != SEQ_NONE) {
            for (int i = 0; i < total; i++) {
                const svdb_candidate *src = in + (size_t)(i / k) * max_rec + (size_t)qi * k + (i % k);
                if (__ldcg(&src->seq) == wl.seq) {
                    c.index = __ldcg(&src->index);
                    break;
                }
            }
        }
        out[(size_t)qi * k + lane] = c;
    }
}

// =====================================================================================
// The fused tail of a scan launch.  Every CTA calls it after its lists are written (all threads); the last one to
// arrive finalizes the launch's queries and, on a sharded store, exchanges and merges.  smem: the CTA's dynamic
// shared memory, free by now.
// =====================================================================================
__device__ __forceinline__ void scan_tail(const TailArgs &t, unsigned char *smem, int smem_bytes) {
    if (t.ticket == nullptr) return;
    __shared__ unsigned s_last;
    __shared__ u64 s_epoch;
    __threadfence();                                   // this CTA's lists are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(t.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;                     // ... before its ticket is
        if (t.dbg) t.dbg[32 + blockIdx.x] = global_timer_ns();
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ unsigned s_stale;
    if (threadIdx.x == 0) {
        unsigned stale;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(stale) : "l"(t.ticket + 3));
        s_stale = stale;                               // a scan that started early (kernels.h: pdl) saw its query change
        t.ticket[0] = 0;                               // re-armed for the next launch on the stream ...
        t.ticket[2] = 0;                               // ... and so is the tile counter of the scans that hand tiles out dynamically
        t.ticket[3] = 0;
        if (t.dbg) t.dbg[0] = global_timer_ns();
    }
    if (threadIdx.x == 0) fin_bar_init(smem, blockDim.x >> 5);
    __syncthreads();
    if (t.dbg && threadIdx.x == 0) t.dbg[9] = global_timer_ns();
    uint32_t phase = 0;
    for (int qi = 0; qi < t.fin.nq; qi++) finalize_query(t.fin, qi, smem, smem_bytes, phase, qi == 0 ? t.dbg : nullptr);
    if (s_stale) {                                     // answers for a query that is not the caller's: never hand them out as good
        for (int i = threadIdx.x; i < t.fin.nq * t.fin.k; i += blockDim.x) t.fin.out[i].flags |= SVDB_CAND_UNSAFE;
        __syncthreads();
    }
    if (t.dbg && threadIdx.x == 0) t.dbg[2] = global_timer_ns();
    if (t.world <= 1) return;
    const int nrec = t.fin.nq * t.fin.k;
    const u64 epoch = xch_push(t.fin.out, nrec, t.peers, t.rank, t.world, t.max_rec, &s_epoch);
    if (t.dbg && threadIdx.x == 0) t.dbg[3] = global_timer_ns();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const unsigned char *mine = t.peers.p[t.rank];
    if (warp < t.fin.nq) xch_wait(mine, t.world, epoch, lane);
    if (t.dbg && threadIdx.x == 0) t.dbg[4] = global_timer_ns();
    const svdb_candidate *in = reinterpret_cast<const svdb_candidate *>(
        mine + XCH_FLAG_BYTES + (size_t)(epoch & 1) * t.world * t.max_rec * sizeof(svdb_candidate));
    for (int qi = warp; qi < t.fin.nq; qi += NW) merge_gathered(in, t.world, t.max_rec, qi, t.fin.k, lane, t.xout);
    if (t.dbg && threadIdx.x == 0) t.dbg[5] = global_timer_ns();
}

}  // namespace svdb
