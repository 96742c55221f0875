// umma_select.cuh -- the selection half of K10 (umma_filter.cu): candidate-buffer entries, the conservative threshold,
// pruning a (CTA, query) buffer to its `keep` smallest keys by bitwise selection, and the final 32-key bitonic sort.
// Split from umma_filter.cu so that the CPU emulation build (tests/cusim/, tests/test_umma_select_sim.py) can run the
// same source: nothing here needs PTX.  Include after common.cuh (device build) or tests/cusim/cusim_common.h.
#pragma once

namespace svdb {

constexpr int UF_BUF = 256;                     // append-buffer entries per (CTA, query)

struct UfEntry {
    float key;
    uint32_t row;
};

// float <-> u32 whose unsigned order is the numeric order (keys are finite: non-finite ones were clamped to -FLT_MAX)
__device__ __forceinline__ uint32_t uf_ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float uf_unord(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// The epilogue tests |x|^2 - 2 acc < thr instead of forming the key first.  thr = tau - |q|^2 plus a slack that covers
// the roundings of both sides (2^-23 relative each), so that every row whose key is below tau passes; a row that
// passes with a key slightly above tau is just one more candidate.  tau = +inf gives +inf.
__device__ __forceinline__ float uf_thr(float tau, float qn) { return (tau - qn) + ldexpf(fabsf(tau) + fabsf(qn), -20); }

// Keep the `keep` smallest of the cnt (<= UF_BUF) entries of one (CTA, query) buffer, compacted to its front in no
// particular order; returns how many are left.  When something was dropped, tau = the largest key kept (every dropped
// key is >= tau).  One warp, entries in registers (8 per lane); the keep-th smallest key is found by a 32-step
// bitwise search with a warp-wide count per step -- a fixed ~2k cycles, where sorted insertion paid ~200 cycles of
// dependent shuffles for every key that entered the list.
__device__ __forceinline__ unsigned uf_prune(UfEntry *b, unsigned cnt, int keep, int lane, float &tau, bool &dropped) {
    constexpr int PER = UF_BUF / 32;
    dropped = false;
    if (cnt <= (unsigned)keep) return cnt;
    uint32_t u[PER], r[PER];
#pragma unroll
    for (int s = 0; s < PER; s++) {
        const unsigned i = s * 32 + lane;
        u[s] = 0xffffffffu;                                  // never a valid key (that would be a NaN pattern)
        r[s] = 0;
        if (i < cnt) {
            const UfEntry e = b[i];
            u[s] = uf_ord(e.key);
            r[s] = e.row;
        }
    }
    uint32_t T = 0;                                          // becomes the keep-th smallest key
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t t = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int s = 0; s < PER; s++) c += u[s] < t;
        if (__reduce_add_sync(FULL, c) < keep) T = t;
    }
    int nlt = 0;
#pragma unroll
    for (int s = 0; s < PER; s++) nlt += u[s] < T;
    unsigned eq_left = (unsigned)(keep - __reduce_add_sync(FULL, nlt));   // keys equal to T that still fit
    __syncwarp(FULL);                                        // every lane holds its entries: the buffer may be overwritten
    const unsigned below = (1u << lane) - 1u;
    unsigned base = 0;
#pragma unroll
    for (int s = 0; s < PER; s++) {
        const bool eq = u[s] == T;
        const unsigned me = __ballot_sync(FULL, eq);
        const bool k = u[s] < T || (eq && (unsigned)__popc(me & below) < eq_left);
        const unsigned mk = __ballot_sync(FULL, k);
        if (k) b[base + __popc(mk & below)] = UfEntry{uf_unord(u[s]), r[s]};
        base += __popc(mk);
        eq_left -= min(eq_left, (unsigned)__popc(me));
    }
    tau = uf_unord(T);
    dropped = true;
    return base;
}

// Ascending bitonic sort of one 64-bit key per lane.
__device__ __forceinline__ u64 uf_sort32(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 o = __shfl_xor_sync(FULL, v, j);
            const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
            v = take_min ? (o < v ? o : v) : (o > v ? o : v);
        }
    }
    return v;
}

}  // namespace svdb
