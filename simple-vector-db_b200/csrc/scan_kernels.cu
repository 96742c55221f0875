// scan_kernels.cu -- K1 (distance scan + on-chip top-k), finalize (merge + exact re-rank) and
// K7 (cross-shard merge) for sm_100a.
//
// Replaces the reference's nearest-neighbour loop, src/kdtree.c:131-162 (entry :171-178):
// the reference walks a pointer tree and, at the dimensions the benchmark uses, visits
// every node; here the append-only log of kd-points is a dense fp64 array in HBM that is
// streamed once per pass.
//
//   scan_wide_kernel   HBM-bound.  Each warp owns a ring of shared-memory stages that it
//                      fills itself with 1-D bulk async copies (cp.async.bulk -> UBLKCP,
//                      completion on an mbarrier), so the bytes in flight per SM are
//                      warps x stages x tile and cost no registers.  Distances are
//                      accumulated lane-parallel with FMA (an approximation of the
//                      reference's sequential sum with relative error <= eps); every warp
//                      keeps its 32 best (d~, seq) in registers, one per lane.
//   scan_ldg_kernel    same contract, rows streamed with ld.global.nc.v2.f64 (A/B variant).
//   scan_exact_kernel  one thread per entry, the reference's exact operation order.
//                      Primary path for thin rows (K <= 12), fallback for any K.
//   finalize_kernel    per query: merge the per-CTA lists, recompute the survivors in the
//                      reference's order (exact_sqdist), order by (d, seq), emit top-k and
//                      prove that no entry outside the candidate set can belong to it.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "scan_common.cuh"
#include "tail.cuh"

namespace svdb {

// =====================================================================================
// Wide rows, TMA bulk-copy ring.  TR rows per tile, NQ queries share the pass.
// LPR = lanes that share one row.  32: a warp walks a row 64 coordinates at a time (rows of any length).
// 16 / 8 / 4 (rows of <= 32 / 16 / 8 doubles; TR = 32, NQ = 1): 32 / LPR rows are processed side by side, one
// 128-bit load per lane and step, so that short rows do not leave most of the warp idle (K = 17: 9 of 32 lanes).
// =====================================================================================
template <int TR, int NQ, int LPR = 32>
__global__ void __launch_bounds__(512, 1) scan_wide_kernel(const __grid_constant__ ScanArgs p, int nstages, int smem_bytes) {
    static_assert(LPR == 32 || (TR == 32 && NQ == 1), "packed rows: 32-row tiles, one query");
    extern __shared__ __align__(128) unsigned char smem[];
    if (p.tail.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.tail.dbg[6] = global_timer_ns();
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = p.stride;
    const uint32_t row_bytes = (uint32_t)kpad * 8u;
    const uint32_t tile_bytes = TR * row_bytes;

    unsigned char *stage_base = smem;
    double *qs = reinterpret_cast<double *>(smem + (size_t)W * nstages * tile_bytes);
    Cand *mrg = reinterpret_cast<Cand *>(reinterpret_cast<unsigned char *>(qs) + (size_t)NQ * row_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(mrg) + (size_t)W * 32 * sizeof(Cand));
    const uint32_t qbar = smem_u32(bars + W * nstages);

    if (threadIdx.x == 0) {
        for (int i = 0; i <= W * nstages; i++) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
    }
    __syncthreads();

    // query tile -> shared memory through the TMA engine
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(qbar, NQ * row_bytes);
        for (int qi = 0; qi < NQ; qi++)
            bulk_g2s(smem_u32(qs) + qi * row_bytes, p.q + (size_t)qi * p.ldq, row_bytes, qbar);
    }

    const u64 ntiles_all = (p.n + TR - 1) / TR;
    // tile assignment: round-robin over all warps of the grid, or one contiguous slab per CTA
    // (then round-robin over the CTA's warps inside the slab)
    u64 gw, GW, ntiles;
    if (p.assign == 1) {
        const u64 per = (ntiles_all + gridDim.x - 1) / gridDim.x;
        const u64 lo = (u64)blockIdx.x * per;
        const u64 hi = lo + per < ntiles_all ? lo + per : ntiles_all;
        gw = lo + warp;
        GW = W;
        ntiles = lo < ntiles_all ? hi : 0;
    } else {
        gw = (u64)blockIdx.x * W + warp;
        GW = (u64)gridDim.x * W;
        ntiles = ntiles_all;
    }
    const uint32_t my_stage = smem_u32(stage_base) + (uint32_t)warp * nstages * tile_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * nstages);

    auto issue = [&](u64 t, int s) {
        const u64 row0 = t * TR;
        const u64 left = p.n - row0;
        const uint32_t rows = left < (u64)TR ? (uint32_t)left : (uint32_t)TR;
        const uint32_t bytes = rows * row_bytes;
        mbar_arrive_expect_tx(my_bar + 8 * s, bytes);
        bulk_g2s(my_stage + s * tile_bytes, p.pts + row0 * (u64)kpad, bytes, my_bar + 8 * s);
    };
    if (lane == 0) {
        for (int s = 0; s < nstages; s++) {
            const u64 t = gw + (u64)s * GW;
            if (t < ntiles) issue(t, s);
        }
    }

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();

    mbar_wait(qbar, 0);

    int s = 0;
    uint32_t phase = 0;
    const uint32_t qa0 = smem_u32(qs) + lane * 16;
    // packed rows: lane = (row within the step) * LPR + (pair of coordinates); the query pair stays in registers
    constexpr int PR = 32 / LPR;                       // rows side by side
    const int pj = lane % LPR, pg = lane / LPR;
    const bool pact = pj * 2 < kpad;
    double2 pq = make_double2(0.0, 0.0);
    if (LPR < 32 && pact) pq = lds128(smem_u32(qs) + pj * 16);
    for (u64 t = gw; t < ntiles; t += GW) {
        mbar_wait(my_bar + 8 * s, phase);
        double cd[NQ];
        int my_row;                                    // row of the tile whose key this lane ends up with
        bool my_own;
        if constexpr (LPR < 32) {
            // step i covers rows i*PR .. i*PR+PR-1; afterwards lane (g, j) holds LPR partial sums, one per step,
            // and the halving reduction over j leaves exactly one finished row in every lane: row j*PR + g
            double v[LPR];
            const uint32_t sa = my_stage + s * tile_bytes + (uint32_t)pg * row_bytes + (uint32_t)pj * 16;
#pragma unroll
            for (int i = 0; i < LPR; i++) {
                const double2 x = pact ? lds128(sa + (uint32_t)(i * PR) * row_bytes) : pq;
                const double a = x.x - pq.x;
                const double b = x.y - pq.y;
                v[i] = fma(b, b, a * a);
            }
#pragma unroll
            for (int st2 = 0, m = LPR / 2; m >= 1; m >>= 1, st2++) {
                const int cnt = LPR >> (st2 + 1);
                const bool up = (lane & m) != 0;
#pragma unroll
                for (int i = 0; i < cnt; i++) {
                    const double send = up ? v[i] : v[i + cnt];
                    const double keep = up ? v[i + cnt] : v[i];
                    v[i] = keep + shfl_xor_f64(send, m);
                }
            }
            cd[0] = v[0];
            my_row = pj * PR + pg;
            my_own = true;
        } else {
            // two independent accumulation chains per (row, query) while registers allow it
            constexpr int NA = (TR * NQ <= 4) ? 2 : 1;
            double acc[TR][NQ][NA];
#pragma unroll
            for (int r = 0; r < TR; r++)
#pragma unroll
                for (int qi = 0; qi < NQ; qi++)
#pragma unroll
                    for (int a = 0; a < NA; a++) acc[r][qi][a] = 0.0;

            uint32_t sa = my_stage + s * tile_bytes + lane * 16;
            uint32_t qa = qa0;
            // tiles of 16 / 32 rows are only picked for rows of <= 64 doubles: a single trip
#pragma unroll(TR >= 16 ? 1 : 4)
            for (int off = lane * 2; off < kpad; off += 64, sa += 512, qa += 512) {
                double2 qv[NQ];
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) qv[qi] = lds128(qa + qi * row_bytes);
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    const double2 x = lds128(sa + r * row_bytes);
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const double a = x.x - qv[qi].x;
                        const double b = x.y - qv[qi].y;
                        acc[r][qi][0] = fma(a, a, acc[r][qi][0]);
                        acc[r][qi][NA - 1] = fma(b, b, acc[r][qi][NA - 1]);
                    }
                }
            }
            // all-reduce over the lanes (reduce_rows: the combination tree is identical for every row, so
            // equal rows always get equal keys wherever they sit in the log)
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                double v[TR];
#pragma unroll
                for (int r = 0; r < TR; r++) v[r] = NA == 2 ? acc[r][qi][0] + acc[r][qi][NA - 1] : acc[r][qi][0];
                reduce_rows<TR>(v, lane);
                cd[qi] = v[0];
            }
            my_row = RowLane<TR>::row(lane);
            my_own = RowLane<TR>::owner(lane);
        }
        __syncwarp();
        // the stage is consumed: refill it before the (rare) list maintenance
        const u64 tn = t + (u64)nstages * GW;
        if (lane == 0 && tn < ntiles) issue(tn, s);
        if (++s == nstages) {
            s = 0;
            phase ^= 1;
        }
        const u64 row = t * TR + my_row;
        const bool has = my_own && row < p.n;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) wl[qi].offer(has, cd[qi], row, lane, p.cap);
    }

    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
    if (threadIdx.x == 0)                                // the tail reuses this memory (and brings its own barrier)
        for (int i = 0; i <= W * nstages; i++) mbar_inval(smem_u32(bars + i));
    scan_tail(p.tail, smem, smem_bytes);
}

// =====================================================================================
// K11: the same scan over the split-bf16 SHADOW of the log (umma_filter.cu: a hi plane and a lo plane of [n][Kp] bf16
// each, 4 bytes per coordinate together) -- HALF the bytes of the fp64 rows, and the scan is HBM-bound.
// Per coordinate x^ = hi + lo (exact in fp32, |x^ - x| <= (2^-16 + 2^-24)|x|), diff = x^ - fl32(q), key = sum diff^2
// accumulated in fp32 (FFMA, lane-parallel, butterfly).  With e_i = x^_i - x_i + q_i - q^_i:
//   |key - d| <= 2 sqrt(d) |e| + |e|^2 + (K/32 + 12) 2^-24 d  <=  (eta + ...) d + (1 + 1/eta) |e|^2,   eta = 2^-13,
// i.e. the (eps, eabs) form finalize_kernel already proves completeness for (shadow_eps / shadow_eabs_coef).
// Same ring of per-warp shared-memory stages fed by 1-D bulk async copies as scan_wide_kernel; TR <= 32 rows per tile,
// lane r keeps the key of row r of the tile, one offer per tile and query.
// =====================================================================================
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

template <int NQ>
__global__ void __launch_bounds__(512, 1) scan_shadow_kernel(const __grid_constant__ ShadowScanArgs p, int nstages, int TR, int smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Kp = p.Kp;
    const uint32_t plane_bytes = (uint32_t)Kp * 2u;   // one row of one plane
    const uint32_t row_bytes = 2u * plane_bytes;
    const uint32_t tile_bytes = (uint32_t)TR * row_bytes;
    const int trips = (Kp + 255) >> 8;                 // a warp covers 256 coordinates per trip, 8 per lane

    // queries in fp32, laid out so that lane l reads its 8 coordinates of a trip with two conflict-free 128-bit loads:
    // float4 index ((query * trips + trip) * 2 + half) * 32 + lane
    float *qs = reinterpret_cast<float *>(smem + (size_t)W * nstages * tile_bytes);
    Cand *mrg = reinterpret_cast<Cand *>(reinterpret_cast<unsigned char *>(qs) + (size_t)NQ * trips * 1024);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(mrg) + (size_t)W * 32 * sizeof(Cand));

    if (threadIdx.x == 0) {
        for (int i = 0; i < W * nstages; i++) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < NQ * trips * 256; i += blockDim.x) {
        const int qi = i / (trips * 256), c = i - qi * (trips * 256);
        const float v = c < p.K ? (float)p.q[(size_t)qi * p.ldq + c] : 0.f;
        qs[((qi * trips + (c >> 8)) * 2 + ((c & 7) >> 2)) * 128 + ((c & 255) >> 3) * 4 + (c & 3)] = v;
    }
    __syncthreads();

    const u64 ntiles = (p.n + TR - 1) / TR;
    const u64 gw = (u64)blockIdx.x * W + warp, GW = (u64)gridDim.x * W;
    const uint32_t my_stage = smem_u32(smem) + (uint32_t)warp * nstages * tile_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * nstages);
    const uint32_t qs_u = smem_u32(qs) + lane * 16;

    auto issue = [&](u64 t, int s) {
        const u64 row0 = t * TR;
        const u64 left = p.n - row0;
        const uint32_t rows = left < (u64)TR ? (uint32_t)left : (uint32_t)TR;
        const uint32_t bytes = rows * plane_bytes;      // per plane; the lo plane of the tile lands behind TR hi rows
        mbar_arrive_expect_tx(my_bar + 8 * s, 2 * bytes);
        bulk_g2s(my_stage + s * tile_bytes, reinterpret_cast<const unsigned char *>(p.xhi) + row0 * (u64)plane_bytes, bytes,
                 my_bar + 8 * s);
        bulk_g2s(my_stage + s * tile_bytes + (uint32_t)TR * plane_bytes,
                 reinterpret_cast<const unsigned char *>(p.xlo) + row0 * (u64)plane_bytes, bytes, my_bar + 8 * s);
    };
    if (lane == 0) {
        for (int s = 0; s < nstages; s++) {
            const u64 t = gw + (u64)s * GW;
            if (t < ntiles) issue(t, s);
        }
    }

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();

    int s = 0;
    uint32_t phase = 0;
    for (u64 t = gw; t < ntiles; t += GW) {
        mbar_wait(my_bar + 8 * s, phase);
        const u64 row0 = t * TR;
        const int rows = (int)(p.n - row0 < (u64)TR ? p.n - row0 : (u64)TR);
        float mykey[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) mykey[qi] = 0.f;
        for (int r = 0; r < rows; r++) {
            float acc[NQ];
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) acc[qi] = 0.f;
            const uint32_t base = my_stage + s * tile_bytes + (uint32_t)r * plane_bytes + lane * 16;
            for (int trip = 0, c0 = lane * 8; c0 < Kp; trip++, c0 += 256) {
                const uint4 h = lds_u128(base + trip * 512);
                const uint4 l = lds_u128(base + (uint32_t)TR * plane_bytes + trip * 512);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
                float x[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {          // a 32-bit word holds coordinates 2j (low half) and 2j+1 (high half)
                    x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
                    x[2 * j + 1] = __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
                }
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const uint32_t qa = qs_u + (uint32_t)((qi * trips + trip) * 2) * 512u;
                    float4 a, b;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(qa));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(qa + 512u));
                    const float q8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float d = x[j] - q8[j];
                        acc[qi] = fmaf(d, d, acc[qi]);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) acc[qi] += __shfl_xor_sync(FULL, acc[qi], m);
                if (lane == r) mykey[qi] = acc[qi];
            }
        }
        __syncwarp();
        // the stage is consumed: refill it before the (rare) list maintenance
        const u64 tn = t + (u64)nstages * GW;
        if (lane == 0 && tn < ntiles) issue(tn, s);
        if (++s == nstages) {
            s = 0;
            phase ^= 1;
        }
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) wl[qi].offer(lane < rows, (double)mykey[qi], row0 + lane, lane, p.cap);
    }

    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
    if (threadIdx.x == 0)
        for (int i = 0; i < W * nstages; i++) mbar_inval(smem_u32(bars + i));
    scan_tail(p.tail, smem, smem_bytes);
}

// =====================================================================================
// Wide rows, direct LDG.128 streaming (A/B variant of the same contract).
// =====================================================================================
template <int TR, int NQ>
__global__ void __launch_bounds__(256, 2) scan_ldg_kernel(ScanArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpad = p.stride;
    const uint32_t row_bytes = (uint32_t)kpad * 8u;
    double *qs = reinterpret_cast<double *>(smem);
    Cand *mrg = reinterpret_cast<Cand *>(smem + (size_t)NQ * row_bytes);

    for (int i = threadIdx.x; i < NQ * kpad; i += blockDim.x) qs[i] = p.q[(size_t)(i / kpad) * p.ldq + (i % kpad)];
    __syncthreads();

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();

    constexpr int U = (TR >= 4) ? 2 : 4;  // chunks loaded ahead of use
    const u64 ntiles = (p.n + TR - 1) / TR;
    const u64 gw = (u64)blockIdx.x * W + warp, GW = (u64)gridDim.x * W;
    const uint32_t qa0 = smem_u32(qs) + lane * 16;
    const int nchunks = (kpad + 63) >> 6;

    for (u64 t = gw; t < ntiles; t += GW) {
        const u64 row0 = t * TR;
        const u64 left = p.n - row0;
        const int rows = left < (u64)TR ? (int)left : TR;
        const double *base = p.pts + row0 * (u64)kpad + lane * 2;
        double acc[TR][NQ];
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) acc[r][qi] = 0.0;

        for (int c0 = 0; c0 < nchunks; c0 += U) {
            double2 x[U][TR];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int off = (c0 + u) * 64 + lane * 2;
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    const int rr = r < rows ? r : rows - 1;
                    x[u][r] = off < kpad ? ldg128_stream(base + (size_t)rr * kpad + (c0 + u) * 64) : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int off = (c0 + u) * 64 + lane * 2;
                if (off < kpad) {
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        const double2 qv = lds128(qa0 + qi * row_bytes + (c0 + u) * 512);
#pragma unroll
                        for (int r = 0; r < TR; r++) {
                            const double a = x[u][r].x - qv.x;
                            const double b = x[u][r].y - qv.y;
                            acc[r][qi] = fma(a, a, acc[r][qi]);
                            acc[r][qi] = fma(b, b, acc[r][qi]);
                        }
                    }
                }
            }
        }
        double cd[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            double v[TR];
#pragma unroll
            for (int r = 0; r < TR; r++) v[r] = acc[r][qi];
            reduce_rows<TR>(v, lane);
            cd[qi] = v[0];
        }
        const bool has = RowLane<TR>::owner(lane) && RowLane<TR>::row(lane) < rows;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) wl[qi].offer(has, cd[qi], row0 + RowLane<TR>::row(lane), lane, p.cap);
    }

    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
}

// =====================================================================================
// Exact scan: one thread per log entry, reference operation order (kdtree.c:134-137).
// =====================================================================================
template <int NQ>
__global__ void __launch_bounds__(512, 1) scan_exact_kernel(ScanArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = p.K;
    double *qs = reinterpret_cast<double *>(smem);
    Cand *mrg = reinterpret_cast<Cand *>(smem + (((size_t)NQ * K * 8 + 15) & ~(size_t)15));

    for (int i = threadIdx.x; i < NQ * K; i += blockDim.x) qs[i] = p.q[(size_t)(i / K) * p.ldq + (i % K)];
    __syncthreads();

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();

    // UNR rows per lane per step: UNR independent load streams in flight (the loop is latency-bound)
    constexpr int UNR = NQ >= 4 ? 1 : (NQ == 2 ? 2 : 4);
    const u64 gw = (u64)blockIdx.x * W + warp, GW = (u64)gridDim.x * W;
    for (u64 base = gw * (32 * UNR); base < p.n; base += GW * (32 * UNR)) {
        double d[UNR][NQ];
        const double *r[UNR];
        bool has[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const u64 row = base + u * 32 + lane;
            has[u] = row < p.n;
            r[u] = p.pts + (has[u] ? row : 0) * (u64)p.stride;
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) d[u][qi] = 0.0;
        }
        for (int i = 0; i < K; i++) {
            double x[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) x[u] = __ldg(r[u] + i);
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const double t = __dsub_rn(x[u], qs[qi * K + i]);
                    d[u][qi] = __dadd_rn(d[u][qi], __dmul_rn(t, t));
                }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++)
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) wl[qi].offer(has[u], d[u][qi], base + u * 32 + lane, lane, p.cap);
    }

    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
}

// =====================================================================================
// finalize as a launch of its own: one CTA of 8 warps per query (tail.cuh: finalize_query).  The single-query scans
// run the same function in their last CTA instead (scan_tail).
// =====================================================================================
constexpr int FIN_WARPS = 8;
constexpr size_t FIN_SMEM = fin_head_bytes(FIN_WARPS) + (size_t)32 * 257 * 8;
// a single-query scan's short lists (selection path of finalize_query): all of them at once, 296 x 15 x 16 bytes
constexpr size_t FIN_SMEM_ALL_LISTS = fin_head_bytes(FIN_WARPS) + (size_t)96 * 1024;

__global__ void __launch_bounds__(FIN_WARPS * 32) finalize_kernel(const __grid_constant__ FinalArgs p, int fsm_bytes) {
    extern __shared__ __align__(128) unsigned char fsm[];
    if (threadIdx.x == 0) fin_bar_init(fsm, FIN_WARPS);
    __syncthreads();
    uint32_t phase = 0;
    finalize_query(p, blockIdx.x, fsm, fsm_bytes, phase);
}

// =====================================================================================
// K7: merge of gathered per-shard results. in: [nshards][nq][k]; one warp per query.
// =====================================================================================
__global__ void __launch_bounds__(32) merge_candidates_kernel(const svdb_candidate *in, int nshards, int nq, int k,
                                                              svdb_candidate *out) {
    const int qi = blockIdx.x, lane = threadIdx.x;
    WarpList wl;
    wl.reset();
    u64 flags = 0;
    const int total = nshards * k;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        double d = CUDART_INF;
        u64 s = SEQ_NONE;
        if (i < total) {
            const svdb_candidate c = in[((size_t)(i / k) * nq + qi) * k + (i % k)];
            d = c.dist;
            s = c.seq;
            flags |= c.flags & ~SVDB_CAND_TIE;
        }
        wl.offer(s != SEQ_NONE, d, s, lane);
    }
    // SVDB_CAND_TIE of the merged answer: >= 2 entries at the merged minimum, or one that was flagged by its shard
    double dmin;
    u64 smin;
    wl.key_at(0, dmin, smin);
    int at_min = 0;
    if (smin != SEQ_NONE) {
        for (int base = 0; base < total; base += 32) {
            const int i = base + lane;
            if (i < total) {
                const svdb_candidate c = in[((size_t)(i / k) * nq + qi) * k + (i % k)];
                if (c.seq != SEQ_NONE && c.dist == dmin) at_min += (c.flags & SVDB_CAND_TIE) ? 2 : 1;
            }
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        flags |= __shfl_xor_sync(FULL, flags, m);
        at_min += __shfl_xor_sync(FULL, at_min, m);
    }
    if (at_min >= 2) flags |= SVDB_CAND_TIE;
    if (lane < k) {
        svdb_candidate c;
        c.dist = wl.d;
        c.seq = wl.seq;
        c.index = (u64)SVDB_NONE;
        c.flags = flags;
        if (wl.seq != SEQ_NONE) {
            for (int i = 0; i < total; i++) {
                const svdb_candidate *src = &in[((size_t)(i / k) * nq + qi) * k + (i % k)];
                if (src->seq == wl.seq) {
                    c.index = src->index;
                    break;
                }
            }
        }
        out[(size_t)qi * k + lane] = c;
    }
}

// ---- small utility kernels -------------------------------------------------------------
__global__ void extract_prefix_kernel(const double *src, int ld_src, double *dst, int ld_dst, int K, u64 n) {
    const u64 total = n * (u64)ld_dst;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const u64 r = i / ld_dst;
        const int c = (int)(i % ld_dst);
        dst[i] = c < K ? src[r * ld_src + c] : 0.0;
    }
}
__global__ void iota_kernel(u64 *dst, u64 base, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] = base + i;
}
__global__ void pad_queries_kernel(const double *src, int ldq, double *dst, int ldp, int K, int nq) {
    const int total = nq * ldp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / ldp, c = i % ldp;
        dst[i] = c < K ? src[(size_t)r * ldq + c] : 0.0;
    }
}

// =====================================================================================
// host-side launchers
// =====================================================================================
static int pick_tile_rows(const ScanTuning &t, int row_bytes, int nq) {
    int tr = t.tile_rows;
    if (tr <= 0) {
        // 5-8 KB tiles measured best (profiles/r01_ksweep_*: 3 KB tiles 0.88-0.90 x peak, 4 KB 1.04-1.05, 5-8 KB
        // 1.06-1.09, 10 KB 1.03): the smallest power of two rows that reaches 5 KB
        tr = 1;
        while (tr < 32 && tr * row_bytes < 5120) tr *= 2;
    }
    // registers: TR x NQ accumulators per lane
    const int budget = nq == 1 ? 32 : 16;
    while (tr > 1 && tr * nq > budget) tr >>= 1;
    int p2 = 1;
    while (p2 * 2 <= tr && p2 < 32) p2 *= 2;
    return p2;
}

int scan_num_lists(const ScanTuning &t, bool /*wide*/) { return t.num_sms * (t.ctas_per_sm > 0 ? t.ctas_per_sm : 1); }

template <int TR, int NQ, int LPR = 32>
static cudaError_t launch_wide_inst(const ScanTuning &t, const ScanArgs &a, cudaStream_t st) {
    const int row_bytes = a.stride * 8;
    const int grid = scan_num_lists(t, true);
    int W = t.warps < 1 ? 1 : (t.warps > 16 ? 16 : t.warps);
    if constexpr (TR > 8 || LPR != 32) {
        if (t.variant == 1) return cudaErrorInvalidValue;     // the LDG variant keeps whole tiles in registers: TR <= 8
    } else if (t.variant == 1) {
        if (W > 8) W = 8;
        const size_t smem = (size_t)NQ * row_bytes + (size_t)W * 32 * sizeof(Cand);
        if (smem > (size_t)MAX_SMEM) return cudaErrorInvalidValue;
        static SmemOptIn optin;
        cudaError_t e = optin.ensure(scan_ldg_kernel<TR, NQ>, smem);
        if (e != cudaSuccess) return e;
        scan_ldg_kernel<TR, NQ><<<grid, W * 32, smem, st>>>(a);
        return cudaGetLastError();
    }
    int NS = t.stages < 2 ? 2 : t.stages;
    // row lengths that do not divide 4 KB leave ~3 KB tiles (K = 48, 96: 0.88-0.90 x peak): keep >= 8 KB per warp in flight
    while (NS < 4 && (size_t)NS * TR * row_bytes < 8192) NS++;
    const int cps = t.ctas_per_sm > 0 ? t.ctas_per_sm : 1;
    auto need = [&](int w, int ns) {
        return (size_t)w * ns * TR * row_bytes + (size_t)NQ * row_bytes + (size_t)w * 32 * sizeof(Cand) + (size_t)(w * ns + 1) * 8;
    };
    size_t budget = (size_t)MAX_SMEM / cps - (cps > 1 ? 1024 : 0);
    const int W0 = W, NS0 = NS;
    while (need(W, NS) > budget && NS > 2) NS--;
    while (need(W, NS) > budget && W > 2) W--;
    if (need(W, NS) > budget) {
        // too big to co-reside cps CTAs per SM (long rows x many queries): keep the grid (the list
        // layout depends on it) and let the CTAs run in waves with the whole SM's shared memory each
        budget = (size_t)MAX_SMEM;
        W = W0;
        NS = NS0;
        while (need(W, NS) > budget && NS > 2) NS--;
        while (need(W, NS) > budget && W > 1) W--;
        if (need(W, NS) > budget) return cudaErrorInvalidValue;
    }
    // the fused tail (finalize in the last CTA) reuses the ring's shared memory
    const size_t tail_need = a.tail.ticket ? fin_head_bytes(W) + std::max<size_t>(FIN_MIN_TBUF, (size_t)grid * a.cap * sizeof(Cand)) : 0;
    const size_t smem = std::max(need(W, NS), std::min(tail_need, budget));
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_wide_kernel<TR, NQ, LPR>, smem);
    if (e != cudaSuccess) return e;
    scan_wide_kernel<TR, NQ, LPR><<<grid, W * 32, smem, st>>>(a, NS, (int)smem);
    return cudaGetLastError();
}

template <int NQ>
static cudaError_t launch_wide_nq(int tr, const ScanTuning &t, const ScanArgs &a, cudaStream_t st) {
    switch (tr) {
        case 32:
            if constexpr (NQ == 1) return launch_wide_inst<32, NQ>(t, a, st);
        case 16:
            if constexpr (NQ == 1) return launch_wide_inst<16, NQ>(t, a, st);
        case 8:
            if constexpr (NQ <= 2) return launch_wide_inst<8, NQ>(t, a, st);
        case 4:
            if constexpr (NQ <= 4) return launch_wide_inst<4, NQ>(t, a, st);
        case 2:
            return launch_wide_inst<2, NQ>(t, a, st);
        default:
            return launch_wide_inst<1, NQ>(t, a, st);
    }
}

cudaError_t launch_scan_wide(const ScanTuning &t, const ScanArgs &a, cudaStream_t st) {
    if (a.stride & 1) return cudaErrorInvalidValue;
    if (a.nq == 1 && a.stride <= 32 && t.variant == 0 && t.tile_rows <= 0) {
        // short rows, one query: several rows side by side in every warp step, 32-row tiles
        if (a.stride <= 8) return launch_wide_inst<32, 1, 4>(t, a, st);
        if (a.stride <= 16) return launch_wide_inst<32, 1, 8>(t, a, st);
        return launch_wide_inst<32, 1, 16>(t, a, st);
    }
    int tr = pick_tile_rows(t, a.stride * 8, a.nq);
    if (t.variant == 1 && tr > 8) tr = 8;
    switch (a.nq) {
        case 1: return launch_wide_nq<1>(tr, t, a, st);
        case 2: return launch_wide_nq<2>(tr, t, a, st);
        case 4: return launch_wide_nq<4>(tr, t, a, st);
        case 8: return launch_wide_nq<8>(tr, t, a, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int NQ>
static cudaError_t launch_shadow_inst(const ScanTuning &t, const ShadowScanArgs &a, cudaStream_t st) {
    const size_t row_bytes = (size_t)a.Kp * 4;
    int TR = (int)(6144 / row_bytes);                  // ~6 KB per bulk copy, like K1's tiles
    TR = TR < 1 ? 1 : (TR > 32 ? 32 : TR);
    const int trips = (a.Kp + 255) / 256;
    const int grid = scan_num_lists(t, true);
    int W = t.warps < 1 ? 1 : (t.warps > 16 ? 16 : t.warps);
    int NS = t.stages < 2 ? 2 : t.stages;
    while (NS < 4 && (size_t)NS * TR * row_bytes < 8192) NS++;
    const int cps = t.ctas_per_sm > 0 ? t.ctas_per_sm : 1;
    auto need = [&](int w, int ns) {
        return (size_t)w * ns * TR * row_bytes + (size_t)NQ * trips * 1024 + (size_t)w * 32 * sizeof(Cand) + (size_t)w * ns * 8;
    };
    size_t budget = (size_t)MAX_SMEM / cps - (cps > 1 ? 1024 : 0);
    const int W0 = W, NS0 = NS;
    while (need(W, NS) > budget && NS > 2) NS--;
    while (need(W, NS) > budget && W > 2) W--;
    if (need(W, NS) > budget) {
        budget = (size_t)MAX_SMEM;
        W = W0;
        NS = NS0;
        while (need(W, NS) > budget && NS > 2) NS--;
        while (need(W, NS) > budget && W > 1) W--;
        if (need(W, NS) > budget) return cudaErrorInvalidValue;
    }
    const size_t smem = std::max(need(W, NS), a.tail.ticket ? fin_head_bytes(W) + FIN_MIN_TBUF : (size_t)0);
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_shadow_kernel<NQ>, smem);
    if (e != cudaSuccess) return e;
    scan_shadow_kernel<NQ><<<grid, W * 32, smem, st>>>(a, NS, TR, (int)smem);
    return cudaGetLastError();
}

cudaError_t launch_scan_shadow(const ScanTuning &t, const ShadowScanArgs &a, cudaStream_t st) {
    if (a.Kp % 64 || a.Kp < a.K || a.n >= (1ull << 32) || !a.xhi || !a.xlo) return cudaErrorInvalidValue;
    switch (a.nq) {
        case 1: return launch_shadow_inst<1>(t, a, st);
        case 2: return launch_shadow_inst<2>(t, a, st);
        case 4: return launch_shadow_inst<4>(t, a, st);
        case 8: return launch_shadow_inst<8>(t, a, st);
        default: return cudaErrorInvalidValue;
    }
}
// key error of K11: |key - d| <= shadow_eps(K) * d + shadow_eabs_coef() * (max|x|^2 + |q|^2)
double shadow_eps(int K) { return ldexp(1.0, -13) + ((double)K / 32.0 + 12.0) * ldexp(1.0, -24); }
double shadow_eabs_coef() {
    const double delta = ldexp(1.0, -16) + ldexp(1.0, -24);
    return (1.0 + 8192.0) * 2.0 * delta * delta * 1.01;
}

template <int NQ>
static cudaError_t launch_exact_inst(const ScanTuning &t, const ScanArgs &a, cudaStream_t st) {
    // latency-bound gather-like loop: fill the SM with warps; small logs do not need them all
    const int W = a.n >= (u64)t.num_sms * 32 * 16 ? 16 : 8;
    const int grid = scan_num_lists(t, false);
    const size_t smem = (((size_t)NQ * a.K * 8 + 15) & ~(size_t)15) + (size_t)W * 32 * sizeof(Cand);
    if (smem > (size_t)MAX_SMEM) return cudaErrorInvalidValue;
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_exact_kernel<NQ>, smem);
    if (e != cudaSuccess) return e;
    scan_exact_kernel<NQ><<<grid, W * 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scan_exact(const ScanTuning &t, const ScanArgs &a, cudaStream_t st) {
    switch (a.nq) {
        case 1: return launch_exact_inst<1>(t, a, st);
        case 2: return launch_exact_inst<2>(t, a, st);
        case 4: return launch_exact_inst<4>(t, a, st);
        case 8: return launch_exact_inst<8>(t, a, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_finalize(const FinalArgs &a, cudaStream_t st) {
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(finalize_kernel, FIN_SMEM_ALL_LISTS);
    if (e != cudaSuccess) return e;
    const size_t lists_bytes = (size_t)a.nlists * a.cap * sizeof(Cand);
    const bool approx = a.sq_mode || a.eps >= 0.0;
    const size_t smem = approx && a.cap < 16 && fin_head_bytes(FIN_WARPS) + lists_bytes > FIN_SMEM &&
                                fin_head_bytes(FIN_WARPS) + lists_bytes <= FIN_SMEM_ALL_LISTS
                            ? FIN_SMEM_ALL_LISTS
                            : FIN_SMEM;
    finalize_kernel<<<a.nq, FIN_WARPS * 32, smem, st>>>(a, (int)smem);
    return cudaGetLastError();
}

cudaError_t launch_merge_candidates(const svdb_candidate *in, int nshards, int nq, int k, svdb_candidate *out,
                                    cudaStream_t st) {
    merge_candidates_kernel<<<nq, 32, 0, st>>>(in, nshards, nq, k, out);
    return cudaGetLastError();
}

cudaError_t launch_extract_prefix(const double *src, int ld_src, double *dst, int ld_dst, int K, u64 n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const u64 total = n * (u64)ld_dst;
    const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    extract_prefix_kernel<<<grid, 256, 0, st>>>(src, ld_src, dst, ld_dst, K, n);
    return cudaGetLastError();
}

cudaError_t launch_iota(u64 *dst, u64 base, u64 n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    iota_kernel<<<grid, 256, 0, st>>>(dst, base, n);
    return cudaGetLastError();
}

cudaError_t launch_pad_queries(const double *src, int ldq, double *dst, int ldp, int K, int nq, cudaStream_t st) {
    if (nq == 0) return cudaSuccess;
    const int total = nq * ldp;
    pad_queries_kernel<<<(total + 255) / 256, 256, 0, st>>>(src, ldq, dst, ldp, K, nq);
    return cudaGetLastError();
}

}  // namespace svdb
