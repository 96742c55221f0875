// warplist.cuh -- the (distance, sequence) candidate key and the warp-distributed sorted top-32 list.
// Split from common.cuh so that the CPU emulation build of median_tree.cu (tests/cusim/) can include it too:
// nothing here needs PTX.  Include through common.cuh (device build) or tests/cusim/cusim_common.h.
#pragma once

namespace svdb {

// Candidate as produced by the scans: 16 bytes, key = (d, seq).
struct __align__(16) Cand {
    double d;
    u64 seq;
};

__device__ __forceinline__ bool key_less(double d1, u64 s1, double d2, u64 s2) {
    return d1 < d2 || (d1 == d2 && s1 < s2);
}

// ---- warp-distributed sorted list: lane i holds the i-th smallest key seen so far ----
struct WarpList {
    double d;   // +inf  = empty slot
    u64 seq;    // SEQ_NONE = empty slot
    __device__ __forceinline__ void reset() {
        d = CUDART_INF;
        seq = SEQ_NONE;
    }
    // All lanes call with the same (nd, ns); nd must be finite.
    __device__ __forceinline__ void insert(double nd, u64 ns, int lane) {
        const bool before = key_less(nd, ns, d, seq);       // monotone over lanes: F..F T..T
        const unsigned m = __ballot_sync(FULL, before);
        const double pd = __shfl_up_sync(FULL, d, 1);
        const u64 ps = __shfl_up_sync(FULL, seq, 1);
        if (before) {
            const bool prev_before = lane > 0 && ((m >> (lane - 1)) & 1u);
            d = prev_before ? pd : nd;
            seq = prev_before ? ps : ns;
        }
    }
    // Key at position pos (warp-uniform), broadcast to all lanes.
    __device__ __forceinline__ void key_at(int pos, double &kd, u64 &ks) const {
        kd = __shfl_sync(FULL, d, pos);
        ks = __shfl_sync(FULL, seq, pos);
    }
    // Merge another ascending list (one key per lane, empty slots = (+inf, SEQ_NONE)): afterwards this list holds
    // the 32 smallest of the 64 keys.  Bitonic: min(mine[i], other[31-i]) is a bitonic sequence of exactly those
    // keys, five compare-exchange stages sort it.  Cost is fixed (24 SHFL), unlike offer(), whose serial inserts
    // cost ~100 cycles per key that gets in -- the better choice when many keys of the other list qualify.
    __device__ __forceinline__ void merge_sorted(double od, u64 os, int lane) {
        const double rd = __shfl_sync(FULL, od, 31 - lane);
        const u64 rs = __shfl_sync(FULL, os, 31 - lane);
        if (key_less(rd, rs, d, seq)) {
            d = rd;
            seq = rs;
        }
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            const double pd = __shfl_xor_sync(FULL, d, j);
            const u64 ps = __shfl_xor_sync(FULL, seq, j);
            const bool upper = (lane & j) != 0;
            // the lower lane of a pair keeps the smaller key, the upper lane the larger one
            if (key_less(pd, ps, d, seq) != upper) {
                d = pd;
                seq = ps;
            }
        }
    }
    // Offer one candidate per lane (has = this lane holds one).  Only the best `lim` keys are
    // maintained (tau = key at lane lim-1); lanes beyond hold sorted leftovers nobody reads.
    __device__ __forceinline__ void offer(bool has, double cd, u64 cs, int lane, int lim = 32) {
        double td;
        u64 ts;
        key_at(lim - 1, td, ts);
        unsigned m = __ballot_sync(FULL, has && cd < CUDART_INF && key_less(cd, cs, td, ts));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const double nd = __shfl_sync(FULL, cd, src);
            const u64 ns = __shfl_sync(FULL, cs, src);
            insert(nd, ns, lane);
        }
    }
};

}  // namespace svdb
