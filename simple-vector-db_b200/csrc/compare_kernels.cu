// compare_kernels.cu -- K3 (/compare metrics over stored rows) and K4 (insert-time norms), sm_100a.
//
// Replaces cosine_similarity / euclidean_distance / dot_product,
// src/vector_database.c:301-313, 322-333, 342-352.  Those accumulate in FLOAT:
//     dot : acc = (float)((double)acc + a[i]*b[i])              (:308, :349)
//     na  : the same with a[i]*a[i] -- depends on one row only -> precomputed at insert (K4)
//     euc : diff = (float)(a[i]-b[i]); sum = sum +f diff *f diff   (:329-330)
// and the result depends on that exact operation order, so each pair is a strictly
// sequential chain over i.  Parallelism comes from the pairs: one pair per lane.
//
// Memory plan (HBM-bound, 2*D*8 bytes per pair): a warp owns 32 pairs and a ring of
// shared-memory stages.  A stage holds one 128-byte chunk of both rows of every pair; it is
// filled with 16-byte cp.async copies in which 8 consecutive lanes fetch one contiguous
// 128-byte row segment (full-sector, coalesced), and XOR-swizzled so that each lane can then
// read ITS pair's chunk with conflict-free 128-bit shared loads.
#include "common.cuh"
#include "kernels.h"

namespace svdb {

constexpr int CMP_WARPS = 4;
constexpr int CMP_STAGES = 3;
constexpr int CMP_STAGE_BYTES = 32 * 2 * 128;  // 32 pairs x (a, b) x 128 B

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct MetricAcc {
    float dot, sum;
    __device__ __forceinline__ void reset() { dot = 0.0f, sum = 0.0f; }
    template <int MODE>
    __device__ __forceinline__ void step(double a, double b) {
        if (MODE != 1) dot = __double2float_rn(__dadd_rn((double)dot, __dmul_rn(a, b)));
        if (MODE == 1 || MODE == 3) {
            const float df = __double2float_rn(__dsub_rn(a, b));
            sum = __fadd_rn(sum, __fmul_rn(df, df));
        }
    }
};

__device__ __forceinline__ float finish_cosine(float dot, float na, float nb) {
    // vector_database.c:312: dot / (sqrt(na) * sqrt(nb)) evaluated in double, returned as float
    return __double2float_rn(__ddiv_rn((double)dot, __dmul_rn(__dsqrt_rn((double)na), __dsqrt_rn((double)nb))));
}
__device__ __forceinline__ float finish_euclid(float sum) { return __double2float_rn(__dsqrt_rn((double)sum)); }

// MODE 0 cosine, 1 euclidean, 2 dot, 3 all three, 4 self-dot of one row (K4)
template <int MODE>
__global__ void __launch_bounds__(CMP_WARPS * 32) compare_kernel(CompareArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = smem_u32(smem) + warp * (CMP_STAGES * CMP_STAGE_BYTES);
    const int nchunks = (p.D + 15) >> 4;
    const u64 ngroups = (p.n + 31) >> 5;
    const u64 gw = (u64)blockIdx.x * CMP_WARPS + warp, GW = (u64)gridDim.x * CMP_WARPS;
    const int sub = lane >> 3, piece = lane & 7;

    for (u64 g = gw; g < ngroups; g += GW) {
        const u64 pid = g * 32 + lane;
        bool ok = pid < p.n;
        u64 ra = 0, rb = 0;  // version rows
        if (ok) {
            if (MODE == 4) {
                ra = rb = p.first + pid;
            } else {
                const u64 x = p.i1[pid], y = p.i2[pid];
                ok = x < p.nrows && y < p.nrows;
                if (ok) {
                    ra = p.cur ? p.cur[x] : x;
                    rb = p.cur ? p.cur[y] : y;
                }
            }
        }
        const double *pa = p.rows + ra * (u64)p.ldr;
        const double *pb = p.rows + rb * (u64)p.ldr;

        auto issue = [&](int c) {
            const uint32_t st = ring + (c % CMP_STAGES) * CMP_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int pr = i * 4 + sub;  // pair whose segment this lane helps to fetch
                const double *sa = reinterpret_cast<const double *>(
                    __shfl_sync(FULL, reinterpret_cast<unsigned long long>(pa), pr));
                const uint32_t col = (uint32_t)((piece ^ (pr & 7)) * 16);
                cp_async16(st + (pr * 2 + 0) * 128 + col, sa + c * 16 + piece * 2);
                if (MODE != 4) {
                    const double *sb = reinterpret_cast<const double *>(
                        __shfl_sync(FULL, reinterpret_cast<unsigned long long>(pb), pr));
                    cp_async16(st + (pr * 2 + 1) * 128 + col, sb + c * 16 + piece * 2);
                }
            }
        };

        MetricAcc acc;
        acc.reset();
#pragma unroll
        for (int c = 0; c < CMP_STAGES - 1; c++) {
            if (c < nchunks) issue(c);
            cp_async_commit();
        }
        for (int c = 0; c < nchunks; c++) {
            if (c + CMP_STAGES - 1 < nchunks) issue(c + CMP_STAGES - 1);
            cp_async_commit();
            cp_async_wait<CMP_STAGES - 1>();
            __syncwarp();
            const uint32_t st = ring + (c % CMP_STAGES) * CMP_STAGE_BYTES + lane * 256;
            const int e0 = c * 16;
            if (e0 + 16 <= p.D) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t col = (uint32_t)((j ^ (lane & 7)) * 16);
                    const double2 a = lds128(st + col);
                    const double2 b = MODE == 4 ? a : lds128(st + 128 + col);
                    acc.step<MODE>(a.x, b.x);
                    acc.step<MODE>(a.y, b.y);
                }
            } else {
                for (int j = 0; j < 8; j++) {
                    const uint32_t col = (uint32_t)((j ^ (lane & 7)) * 16);
                    const double2 a = lds128(st + col);
                    const double2 b = MODE == 4 ? a : lds128(st + 128 + col);
                    if (e0 + 2 * j < p.D) acc.step<MODE>(a.x, b.x);
                    if (e0 + 2 * j + 1 < p.D) acc.step<MODE>(a.y, b.y);
                }
            }
            __syncwarp();
        }
        cp_async_wait<0>();

        if (pid < p.n) {
            if (MODE == 3) {
                float *o = p.out + pid * 3;
                if (ok) {
                    o[0] = finish_cosine(acc.dot, p.norm[ra], p.norm[rb]);
                    o[1] = finish_euclid(acc.sum);
                    o[2] = acc.dot;
                } else {
                    o[0] = o[1] = o[2] = -1.0f;
                }
            } else {
                float r = -1.0f;  // vector_database.c:302-305 sentinel
                if (ok) {
                    if (MODE == 0) r = finish_cosine(acc.dot, p.norm[ra], p.norm[rb]);
                    if (MODE == 1) r = finish_euclid(acc.sum);
                    if (MODE == 2 || MODE == 4) r = acc.dot;
                }
                p.out[pid] = r;
            }
        }
    }
}

cudaError_t launch_compare(const CompareArgs &a, int num_sms, cudaStream_t st) {
    if (a.n == 0) return cudaSuccess;
    if (a.ldr % 16 != 0) return cudaErrorInvalidValue;
    const size_t smem = (size_t)CMP_WARPS * CMP_STAGES * CMP_STAGE_BYTES;  // 96 KB: two CTAs per SM
    const u64 ngroups = (a.n + 31) / 32;
    u64 grid = (ngroups + CMP_WARPS - 1) / CMP_WARPS;
    const u64 maxgrid = (u64)num_sms * 2;
    if (grid > maxgrid) grid = maxgrid;
#define SVDB_CMP_LAUNCH(M)                                                                                     \
    case M: {                                                                                                  \
        static bool configured = false;                                                                        \
        if (!configured) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(compare_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                    \
            configured = true;                                                                                 \
        }                                                                                                      \
        compare_kernel<M><<<(unsigned)grid, CMP_WARPS * 32, smem, st>>>(a);                                    \
        break;                                                                                                 \
    }
    switch (a.mode) {
        SVDB_CMP_LAUNCH(0)
        SVDB_CMP_LAUNCH(1)
        SVDB_CMP_LAUNCH(2)
        SVDB_CMP_LAUNCH(3)
        SVDB_CMP_LAUNCH(4)
        default:
            return cudaErrorInvalidValue;
    }
#undef SVDB_CMP_LAUNCH
    return cudaGetLastError();
}

}  // namespace svdb
