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
// shared-memory stages.  A stage holds one CH-coordinate segment (default 32 = 256 bytes) of
// both rows of every pair; it is filled with 16-byte cp.async copies in which 8 consecutive
// lanes fetch one contiguous 128-byte line (full sectors) and the lines of a segment are
// requested back to back (one DRAM burst per row instead of two), XOR-swizzled so that each
// lane can then read ITS pair's segment with conflict-free 128-bit shared loads.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "smem_optin.h"

namespace svdb {

constexpr int CMP_STAGES = 3;
// CH = coordinates of each row fetched per stage (16 / 32 / 64 -> 128 / 256 / 512-byte bursts per row);
// a warp-stage holds 32 pairs x (a, b) x CH doubles

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
template <int MODE, int CH, int CMP_WARPS>
__global__ void __launch_bounds__(CMP_WARPS * 32) compare_kernel(CompareArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ROWB = CH * 8;                       // bytes of one row segment
    constexpr int CMP_STAGE_BYTES = 32 * 2 * ROWB;
    constexpr int PIECES = CH / 2;                     // 16-byte pieces per row segment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = smem_u32(smem) + warp * (CMP_STAGES * CMP_STAGE_BYTES);
    const int nchunks = (p.D + CH - 1) / CH;
    const u64 ngroups = (p.n + 31) >> 5;
    const u64 gw = (u64)blockIdx.x * CMP_WARPS + warp, GW = (u64)gridDim.x * CMP_WARPS;
    const int sub = lane >> 3, piece = lane & 7;

    // Which rows a pair needs is two dependent loads deep (pair index -> current version) before the rows
    // themselves can be requested.  Both are taken off a group's critical path: the indices are loaded two
    // groups ahead, the versions looked up one group ahead, in registers.
    constexpr u64 NO_ROW = ~0ull;
    auto load_pair = [&](u64 gg, u64 &x, u64 &y) {
        x = y = NO_ROW;
        const u64 pid = gg * 32 + lane;
        if (gg < ngroups && pid < p.n) {
            if (MODE == 4) {
                x = y = p.first + pid;
            } else {
                x = p.i1[pid];
                y = p.i2[pid];
            }
        }
    };
    auto look_up = [&](u64 x, u64 y, u64 &r1, u64 &r2) {
        r1 = r2 = NO_ROW;
        if (MODE == 4) {
            r1 = r2 = x;
        } else if (x < p.nrows && y < p.nrows) {       // NO_ROW fails this too
            r1 = p.cur ? p.cur[x] : x;
            r2 = p.cur ? p.cur[y] : y;
        }
    };
    u64 nx, ny, nra, nrb;                              // indices of the group after next / rows of the next group
    load_pair(gw, nx, ny);
    look_up(nx, ny, nra, nrb);
    load_pair(gw + GW, nx, ny);

    for (u64 g = gw; g < ngroups; g += GW) {
        const u64 pid = g * 32 + lane;
        const bool ok = nra != NO_ROW;
        const u64 ra = ok ? nra : 0, rb = ok ? nrb : 0;  // out-of-range pairs fetch row 0 and answer the sentinel
        look_up(nx, ny, nra, nrb);                      // next group: in flight while this one is processed
        load_pair(g + 2 * GW, nx, ny);
        const double *pa = p.rows + ra * (u64)p.ldr;
        const double *pb = p.rows + rb * (u64)p.ldr;
        // cosine: the two norms are one more dependent random load each -- ask for them now, use them at the end
        float na = 0.0f, nb = 0.0f;
        if ((MODE == 0 || MODE == 3) && ok) {
            na = __ldg(p.norm + ra);
            nb = __ldg(p.norm + rb);
        }

        // rows are padded to 16 doubles; a CH > 16 segment may reach past the padded row end: clamp
        auto issue = [&](int c) {
            const uint32_t st = ring + (c % CMP_STAGES) * CMP_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int pr = i * 4 + sub;  // pair whose segment this lane helps to fetch
                const double *sa = reinterpret_cast<const double *>(
                    __shfl_sync(FULL, reinterpret_cast<unsigned long long>(pa), pr));
                const double *sb = reinterpret_cast<const double *>(
                    __shfl_sync(FULL, reinterpret_cast<unsigned long long>(pb), pr));
#pragma unroll
                for (int l = 0; l < PIECES / 8; l++) {     // 128-byte lines of the segment
                    const int e = c * CH + l * 16 + piece * 2;
                    if (e < p.ldr) {
                        const uint32_t col = (uint32_t)(l * 128 + ((piece ^ (pr & 7)) << 4));
                        cp_async16(st + (pr * 2 + 0) * ROWB + col, sa + e);
                        if (MODE != 4) cp_async16(st + (pr * 2 + 1) * ROWB + col, sb + e);
                    }
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
            const uint32_t st = ring + (c % CMP_STAGES) * CMP_STAGE_BYTES + lane * (2 * ROWB);
            const int e0 = c * CH;
            if (e0 + CH <= p.D) {
#pragma unroll
                for (int j = 0; j < PIECES; j++) {
                    const uint32_t col = (uint32_t)((j >> 3) * 128 + (((j & 7) ^ (lane & 7)) << 4));
                    const double2 a = lds128(st + col);
                    const double2 b = MODE == 4 ? a : lds128(st + ROWB + col);
                    acc.step<MODE>(a.x, b.x);
                    acc.step<MODE>(a.y, b.y);
                }
            } else {
                for (int j = 0; j < PIECES && e0 + 2 * j < p.D; j++) {
                    const uint32_t col = (uint32_t)((j >> 3) * 128 + (((j & 7) ^ (lane & 7)) << 4));
                    const double2 a = lds128(st + col);
                    const double2 b = MODE == 4 ? a : lds128(st + ROWB + col);
                    acc.step<MODE>(a.x, b.x);
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
                    o[0] = finish_cosine(acc.dot, na, nb);
                    o[1] = finish_euclid(acc.sum);
                    o[2] = acc.dot;
                } else {
                    o[0] = o[1] = o[2] = -1.0f;
                }
            } else {
                float r = -1.0f;  // vector_database.c:302-305 sentinel
                if (ok) {
                    if (MODE == 0) r = finish_cosine(acc.dot, na, nb);
                    if (MODE == 1) r = finish_euclid(acc.sum);
                    if (MODE == 2 || MODE == 4) r = acc.dot;
                }
                p.out[pid] = r;
            }
        }
    }
}

template <int MODE, int CH, int W>
static cudaError_t launch_compare_inst(const CompareArgs &a, int num_sms, cudaStream_t st) {
    constexpr size_t smem = (size_t)W * CMP_STAGES * 32 * 2 * CH * 8;   // 96 KB (two CTAs per SM) or 192 KB
    const u64 ngroups = (a.n + 31) / 32;
    u64 grid = (ngroups + W - 1) / W;
    const u64 maxgrid = (u64)num_sms * (smem <= 100 * 1024 ? 2 : 1);
    if (grid > maxgrid) grid = maxgrid;
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(compare_kernel<MODE, CH, W>, smem);
    if (e != cudaSuccess) return e;
    compare_kernel<MODE, CH, W><<<(unsigned)grid, W * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_compare_mode(const CompareArgs &a, int num_sms, int chunk, cudaStream_t st) {
    switch (chunk) {
        case 64: return launch_compare_inst<MODE, 64, 2>(a, num_sms, st);
        case 32: return launch_compare_inst<MODE, 32, 4>(a, num_sms, st);
        default: return launch_compare_inst<MODE, 16, 4>(a, num_sms, st);
    }
}

cudaError_t launch_compare(const CompareArgs &a, int num_sms, cudaStream_t st) {
    if (a.n == 0) return cudaSuccess;
    if (a.ldr % 16 != 0) return cudaErrorInvalidValue;
    static int chunk = 0;
    if (chunk == 0) {
        const char *v = getenv("SVDB_CMP_CHUNK");
        chunk = v ? atoi(v) : 32;   // 256-byte bursts: 1.03x the measured copy peak at D = 1536 (128-byte: 0.71x)
        if (chunk != 16 && chunk != 32 && chunk != 64) chunk = 32;
    }
    // short rows gain nothing from long bursts
    const int ch = a.D >= 4 * chunk ? chunk : 16;
    switch (a.mode) {
        case 0: return launch_compare_mode<0>(a, num_sms, ch, st);
        case 1: return launch_compare_mode<1>(a, num_sms, ch, st);
        case 2: return launch_compare_mode<2>(a, num_sms, ch, st);
        case 3: return launch_compare_mode<3>(a, num_sms, ch, st);
        case 4: return launch_compare_mode<4>(a, num_sms, ch, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svdb
