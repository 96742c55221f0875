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

}  // namespace svdb
#include "warplist.cuh"
namespace svdb {

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Programmatic dependent launch (griddepcontrol, sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start once every CTA of the kernel in front of it on the stream has executed pdl_launch_dependents() (or exited);
// pdl_wait() returns when that kernel has COMPLETED and its memory is visible.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// order-sensitive 64-bit mix of a query's raw bits: what a scan that started early re-checks after pdl_wait()
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long bits) {
    return (h ^ bits) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
}
__device__ __forceinline__ double ld_cv_f64(const double *p) {
    double v;
    asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

// ---- K13's one-byte plane (plane_scan.cu): x^ = lo + step * u, u in 0..255; queries on a 16-bit grid, q^ = lo + step * Q / 256.
// The SAME expressions quantise rows (build), queries (scan) and form |q - q^| (finalize): the bound needs them identical.
__device__ __forceinline__ unsigned p8_quant_x(double v, double lo, double step) {
    const double t = rint((v - lo) / step);
    return t >= 255.0 ? 255u : (t > 0.0 ? (unsigned)t : 0u);          // NaN -> 0
}
__device__ __forceinline__ unsigned p8_quant_q(double v, double lo, double step) {
    const double t = rint(256.0 * ((v - lo) / step));
    return t >= 65535.0 ? 65535u : (t > 0.0 ? (unsigned)t : 0u);
}

// ---- reference-order arithmetic (kdtree.c:134-137): rounded sub, mul, add; never fused ----
__device__ __forceinline__ double exact_sqdist(const double *__restrict__ p, const double *__restrict__ q, int K) {
    double d = 0.0;
    for (int i = 0; i < K; i++) {
        const double t = __dsub_rn(p[i], q[i]);
        d = __dadd_rn(d, __dmul_rn(t, t));
    }
    return d;
}

}  // namespace svdb
