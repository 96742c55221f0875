// common.cuh -- shared device helpers: PTX wrappers (mbarrier, bulk async copy),
// the (distance, sequence) candidate key and the warp-distributed top-32 list.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace svdb {

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;
constexpr u64 SEQ_NONE = ~0ull;

// Candidate as produced by the scans: 16 bytes, key = (d, seq).
struct __align__(16) Cand {
    double d;
    u64 seq;
};

__device__ __forceinline__ bool key_less(double d1, u64 s1, double d2, u64 s2) {
    return d1 < d2 || (d1 == d2 && s1 < s2);
}

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waits are bounded: a transaction that never completes (a bug, never a data condition)
// traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); completes on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 ldg128_stream(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(FULL, v, m); }

// ---- reference-order arithmetic (kdtree.c:134-137): rounded sub, mul, add; never fused ----
__device__ __forceinline__ double exact_sqdist(const double *__restrict__ p, const double *__restrict__ q, int K) {
    double d = 0.0;
    for (int i = 0; i < K; i++) {
        const double t = __dsub_rn(p[i], q[i]);
        d = __dadd_rn(d, __dmul_rn(t, t));
    }
    return d;
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
