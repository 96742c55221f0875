// plane_scan.cu -- K12: the single-query distance scan over ONE low-precision plane of the log, sm_100a.
//
// Replaces the reference's nearest-neighbour loop, src/kdtree.c:131-162 (entry :171-178), for wide kd-points.  K1
// (scan_kernels.cu) streams the fp64 log at the HBM roofline, so a faster answer needs FEWER BYTES: this kernel reads only
// the bf16 HI plane of the split-bf16 shadow that K10 keeps anyway (umma_filter.cu: x^ = bf16(fl32(x)), [n][Kp] bf16,
// 2 bytes per coordinate = a quarter of the fp64 rows), forms key = |x^ - fl32(q)|^2 in fp32 and leaves exactness to the
// re-rank: finalize_query recomputes the survivors from the fp64 rows in the reference's operation order and PROVES that
// no row outside the candidate lists can belong to the top-k, from the bound in its natural (square-root) form
//     | sqrt(key) - sqrt(d) |  <=  E + gamma (sqrt(d) + E),
//     E = max_r |x_r - x^_r|_2 (measured in fp64 when the plane is written, plane_err) + |q - fl32(q)|_2,
//     gamma = (Kp/32 + 12) 2^-24 (fp32 rounding of differences, squares and the lane-parallel sum; all terms >= 0)
// -- a triangle inequality, no Cauchy-Schwarz slack on the cross term.  When the proof fails (near-ties closer than the
// plane resolves) the query is flagged SVDB_CAND_UNSAFE and re-answered from the fp64 rows (K1); answers are identical
// to the reference either way.
//
// HBM-bound: algorithmic bytes per launch = n * Kp * 2.  Same per-warp ring of shared-memory stages fed by 1-D bulk
// async copies (UBLKCP) as K1; the query lives in REGISTERS (fp32, 8 coordinates per lane and 256-coordinate trip), so
// the only shared-memory traffic is one LDS.128 per 8 coordinates of the plane.  Short rows (Kp = 64, 128) are packed
// 4 / 2 to a warp step.  The launch's last CTA runs finalize (and the cross-shard exchange) itself: one launch per query.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "scan_common.cuh"
#include "tail.cuh"

namespace svdb {

__device__ __forceinline__ uint4 pl_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// eight bf16 (four 32-bit words, even coordinate in the low half) -> fp32, exact
__device__ __forceinline__ void bf16x8_to_f32(const uint4 &h, float (&x)[8]) {
    const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        x[2 * j] = __uint_as_float(w[j] << 16);
        x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
}

// TRIPS >= 1: Kp <= 256 * TRIPS, queries in registers.  LPR = 32: a warp walks one row per step, TR rows per tile.
// LPR = 16 / 8 (TRIPS = 1, NQ = 1, TR = 32, Kp = 128 / 64): 2 / 4 rows side by side.
template <int NQ, int TRIPS, int LPR, int TR>
__global__ void __launch_bounds__(256, 2) scan_plane_kernel(const __grid_constant__ PlaneScanArgs p, int nstages, int smem_bytes) {
    static_assert(LPR == 32 || (TRIPS == 1 && NQ == 1 && TR == 32), "packed rows: one trip, one query, 32-row tiles");
    extern __shared__ __align__(128) unsigned char smem[];
    if (p.tail.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.tail.dbg[6] = global_timer_ns();
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Kp = p.Kp;
    const uint32_t row_bytes = (uint32_t)Kp * 2u;
    const uint32_t tile_bytes = (uint32_t)TR * row_bytes;
    Cand *mrg = reinterpret_cast<Cand *>(smem + (size_t)W * nstages * tile_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(mrg) + (size_t)W * 32 * sizeof(Cand));
    uint32_t *tile_of = reinterpret_cast<uint32_t *>(bars + W * nstages) + warp * 4;   // which tile sits in each of this warp's stages

    if (threadIdx.x == 0) {
        for (int i = 0; i < W * nstages; i++) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
    }
    __syncthreads();

    // Load balance (option "scan.dynamic_tiles", off by default).  With the static round-robin the CTAs of a 1.25M-row
    // shard finish 11 us apart (4 % of the scan: HBM channel luck), and the step ends with the slowest.  Handing tiles out
    // through a device counter (next to the tail's ticket, re-armed by the last CTA) closes the spread to 3 us -- but one
    // address takes ~1 G atomics/s: with every tile dynamic the scan ran 33 % SLOWER, with the last eighth dynamic still
    // 2.6 % slower than static (profiles/r02_K12_dynamic_tiles_ab.txt).  Kept as an A/B: p.dyn_eighths of a warp's
    // share are dynamic.  A warp asks for its next tile BEFORE it waits for the current one.
    const u64 ntiles = (p.n + TR - 1) / TR;
    const u64 gw = (u64)blockIdx.x * W + warp, GW = (u64)gridDim.x * W;
    unsigned *next_tile = (p.tail.ticket && p.dyn_eighths > 0) ? p.tail.ticket + 2 : nullptr;
    const u64 n_static = next_tile ? (ntiles / GW) * (u64)(8 - p.dyn_eighths) / 8 : ~0ull;   // static tiles per warp
    const u64 t_static = next_tile ? n_static * GW : 0;                       // tiles [0, t_static) are static
    u64 my_i = 0;                                                              // tiles this warp has asked for so far
    constexpr uint32_t NO_TILE = 0xffffffffu;
    const uint32_t my_stage = smem_u32(smem) + (uint32_t)warp * nstages * tile_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * nstages);
    auto issue = [&](u64 t, int s) {
        const u64 row0 = t * TR;
        const u64 left = p.n - row0;
        const uint32_t rows = left < (u64)TR ? (uint32_t)left : (uint32_t)TR;
        const uint32_t bytes = rows * row_bytes;
        mbar_arrive_expect_tx(my_bar + 8 * s, bytes);
        bulk_g2s(my_stage + s * tile_bytes, reinterpret_cast<const unsigned char *>(p.xhi) + row0 * (u64)row_bytes, bytes,
                 my_bar + 8 * s);
    };
    auto next = [&]() -> u64 {                            // lane 0 only
        if (my_i < n_static) return gw + (my_i++) * GW;
        return t_static + (u64)atomicAdd(next_tile, 1u);
    };
    if (lane == 0) {
        for (int s = 0; s < nstages; s++) {
            const u64 t = next();
            if (t < ntiles) issue(t, s);
            tile_of[s] = t < ntiles ? (uint32_t)t : NO_TILE;
        }
    }
    __syncwarp();

    // the query, fp32, in registers: lane (j = lane % LPR) owns coordinates trip * 256 + j * 8 .. + 7; zeros beyond K
    // match the zero padding of the plane
    const int pj = lane % LPR, pg = lane / LPR;
    float qr[NQ][TRIPS][8];
    bool act[TRIPS];
    unsigned long long qhash = 0;                       // of the raw bits this lane read (re-checked after pdl_wait)
#pragma unroll
    for (int t = 0; t < TRIPS; t++) {
        const int c0 = t * 256 + pj * 8;
        act[t] = c0 < Kp;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double qc = c0 + j < p.K ? ld_cv_f64(p.q + (size_t)qi * p.ldq + c0 + j) : 0.0;
                qhash = mix64(qhash, (unsigned long long)__double_as_longlong(qc));
                qr[qi][t][j] = __double2float_rn(qc);
            }
    }

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        const uint32_t t32 = tile_of[s];
        if (t32 == NO_TILE) break;                     // tiles are handed out in order: nothing follows an empty stage
        const u64 t = t32;
        // the tile that will refill this stage: asked for now, needed after the arithmetic
        u64 tn = ~0ull;
        if (lane == 0) tn = next();
        mbar_wait(my_bar + 8 * s, phase);
        float key[NQ];
        int my_row;
        bool my_own;
        if constexpr (LPR < 32) {
            constexpr int PR = 32 / LPR;               // rows side by side
            float v[LPR];
            const uint32_t sa = my_stage + s * tile_bytes + (uint32_t)pg * row_bytes + (uint32_t)pj * 16;
#pragma unroll
            for (int i = 0; i < LPR; i++) {
                float x[8];
                bf16x8_to_f32(pl_lds128(sa + (uint32_t)(i * PR) * row_bytes), x);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float d0 = x[j] - qr[0][0][j], d1 = x[j + 1] - qr[0][0][j + 1];
                    a0 = fmaf(d0, d0, a0);
                    a1 = fmaf(d1, d1, a1);
                }
                v[i] = a0 + a1;
            }
            reduce_packed<LPR>(v, lane);
            key[0] = v[0];
            my_row = pj * PR + pg;
            my_own = true;
        } else {
            float acc[TR][NQ];
#pragma unroll
            for (int r = 0; r < TR; r++)
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) acc[r][qi] = 0.f;
            const uint32_t sa = my_stage + s * tile_bytes + lane * 16;
            constexpr int RC = TR < 4 ? TR : 4;        // rows whose loads are in flight together
#pragma unroll
            for (int tr = 0; tr < TRIPS; tr++) {
#pragma unroll
                for (int r0 = 0; r0 < TR; r0 += RC) {
                    uint4 h[RC];
#pragma unroll
                    for (int r = 0; r < RC; r++)
                        h[r] = act[tr] ? pl_lds128(sa + (r0 + r) * row_bytes + tr * 512) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int r = 0; r < RC; r++) {
                        float x[8];
                        bf16x8_to_f32(h[r], x);
#pragma unroll
                        for (int qi = 0; qi < NQ; qi++) {
                            float a0 = acc[r0 + r][qi], a1 = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                const float d0 = x[j] - qr[qi][tr][j], d1 = x[j + 1] - qr[qi][tr][j + 1];
                                a0 = fmaf(d0, d0, a0);
                                a1 = fmaf(d1, d1, a1);
                            }
                            acc[r0 + r][qi] = a0 + a1;
                        }
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                float v[TR];
#pragma unroll
                for (int r = 0; r < TR; r++) v[r] = acc[r][qi];
                reduce_rows<TR>(v, lane);
                key[qi] = v[0];
            }
            my_row = RowLane<TR>::row(lane);
            my_own = RowLane<TR>::owner(lane);
        }
        __syncwarp();
        // the stage is consumed: refill it before the (rare) list maintenance
        if (lane == 0) {
            if (tn < ntiles) issue(tn, s);
            tile_of[s] = tn < ntiles ? (uint32_t)tn : NO_TILE;
        }
        __syncwarp();
        if (++s == nstages) {
            s = 0;
            phase ^= 1;
        }
        const u64 row = t * TR + my_row;
        const bool has = my_own && row < p.n;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) wl[qi].offer(has, (double)key[qi], row, lane, p.cap);
    }

    if (p.pdl) {                                        // see scan_plane8_kernel
        pdl_launch_dependents();
        pdl_wait();
        unsigned long long h2 = 0;
#pragma unroll
        for (int t = 0; t < TRIPS; t++)
#pragma unroll
            for (int qi = 0; qi < NQ; qi++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int c = t * 256 + pj * 8 + j;
                    h2 = mix64(h2, (unsigned long long)__double_as_longlong(c < p.K ? ld_cv_f64(p.q + (size_t)qi * p.ldq + c) : 0.0));
                }
        if (__any_sync(FULL, h2 != qhash) && lane == 0) atomicOr(p.tail.ticket + 3, 1u);
    }
    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
    if (threadIdx.x == 0)                                // the tail reuses this memory (and brings its own barrier)
        for (int i = 0; i < W * nstages; i++) mbar_inval(smem_u32(bars + i));
    scan_tail(p.tail, smem, smem_bytes);
}

// ---- host ----------------------------------------------------------------------------------------------------------
double plane_gamma(int Kp) { return ((double)Kp / 32.0 + 12.0) * ldexp(1.0, -24); }

template <int NQ, int TRIPS, int LPR, int TR>
static cudaError_t launch_plane_inst(const ScanTuning &t, const PlaneScanArgs &a, cudaStream_t st) {
    const size_t row_bytes = (size_t)a.Kp * 2;
    const int grid = a.grid > 0 ? a.grid : scan_num_lists(t, true);
    int W = t.warps < 1 ? 1 : (t.warps > 8 ? 8 : t.warps);
    int NS = t.stages < 2 ? 2 : t.stages;
    while (NS < 4 && (size_t)NS * TR * row_bytes < 8192) NS++;
    const int cps = t.ctas_per_sm > 0 ? t.ctas_per_sm : 1;
    auto need = [&](int w, int ns) { return (size_t)w * ns * TR * row_bytes + (size_t)w * 32 * sizeof(Cand) + (size_t)w * ns * 8 + (size_t)w * 16; };
    const size_t budget = (size_t)MAX_SMEM / cps - (cps > 1 ? 1024 : 0);
    while (need(W, NS) > budget && NS > 2) NS--;
    while (need(W, NS) > budget && W > 1) W--;
    if (need(W, NS) > budget) return cudaErrorInvalidValue;
    // the tail wants every CTA's list in shared memory at once (selection path): nlists x cap x 16 bytes per query
    const size_t tail_need = a.tail.ticket ? fin_head_bytes(W) + std::max<size_t>(FIN_MIN_TBUF, (size_t)grid * a.cap * sizeof(Cand)) : 0;
    size_t smem = std::max(need(W, NS), std::min(tail_need, budget));
    if (a.grid > 0) smem = std::max(smem, std::min(budget, (size_t)MAX_SMEM / (cps + 1) + 1024));    // see launch_plane8_inst
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_plane_kernel<NQ, TRIPS, LPR, TR>, smem);
    if (e != cudaSuccess) return e;
    if (a.pdl && a.tail.ticket) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)(W * 32));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, scan_plane_kernel<NQ, TRIPS, LPR, TR>, a, NS, (int)smem);
    }
    scan_plane_kernel<NQ, TRIPS, LPR, TR><<<grid, W * 32, smem, st>>>(a, NS, (int)smem);
    return cudaGetLastError();
}

bool plane_scan_supports(int Kp, int nq) { return Kp >= 64 && Kp % 64 == 0 && Kp <= 1024 && (nq == 1 || (nq == 2 && Kp <= 512)); }

// Tiles of 4-8 KB (K1's measurements: 5-8 KB bulk copies stream best).
cudaError_t launch_scan_plane(const ScanTuning &t, const PlaneScanArgs &a, cudaStream_t st) {
    if (!plane_scan_supports(a.Kp, a.nq) || a.Kp < a.K || !a.xhi || a.n == 0) return cudaErrorInvalidValue;
    const int Kp = a.Kp;
    if (a.nq == 1) {
        if (Kp == 64) return launch_plane_inst<1, 1, 8, 32>(t, a, st);
        if (Kp == 128) return launch_plane_inst<1, 1, 16, 32>(t, a, st);
        if (Kp <= 256) return launch_plane_inst<1, 1, 32, 16>(t, a, st);
        if (Kp <= 512) return launch_plane_inst<1, 2, 32, 8>(t, a, st);
        if (Kp <= 768) return launch_plane_inst<1, 3, 32, 4>(t, a, st);
        return launch_plane_inst<1, 4, 32, 4>(t, a, st);
    }
    if (Kp <= 256) return launch_plane_inst<2, 1, 32, 16>(t, a, st);
    return launch_plane_inst<2, 2, 32, 8>(t, a, st);
}

// =====================================================================================================================
// K13: the same scan over a ONE-BYTE plane.  x^_i = lo + step * u_i with u_i = round((x_i - lo) / step) in 0..255 on ONE
// store-wide grid (lo, step from the range of the rows present when the plane is first built; rows appended later are
// clamped).  The query goes onto a grid 256 times finer, Q_i = round(256 (q_i - lo) / step) in 0..65535 = 256 a_i + b_i, and
//     sum_i (256 u_i - Q_i)^2 = 65536 sum u_i^2 - 512 (256 sum u_i a_i + sum u_i b_i) + sum Q_i^2
// is formed EXACTLY from three byte dot products (dp4a, 4 coordinates per instruction, the query's digits in registers):
// key = (step / 256)^2 * that = |x^ - q^|^2 with one rounding.  Nothing about the arithmetic is approximate; the only error
// is the quantisation itself, measured when the plane is written (max_r |x_r - x^_r|_2) and, for the query, formed by
// finalize from the same expression: the sqrt-form bound of K12 with gamma = 2^-50.  A quarter of K12's bytes again
// (an eighth of the fp64 rows): algorithmic bytes per launch = n * Kp.  HBM-bound; one launch per query (fused tail).
// Data a uniform grid resolves badly (heavy tails, a few huge coordinates) shows up as a large measured error: the
// proof fails, the query is re-answered from the fp64 rows, and the engine stops using the plane (engine.cu).
// =====================================================================================================================
// LPR = 32: a warp walks one row per step (8 bytes per lane and 256-coordinate trip).  LPR = 8 / 4 (TRIPS = 1, Kp = 128 / 64):
// 16 bytes per lane, 4 / 8 rows side by side, 32 rows per round, TR / 32 rounds per tile (tiles of 8 KB whatever the row length).
// NQ = 2 (LPR = 32 only): two queries share the pass -- the bytes and the u.u dot products are read and formed once, each
// query adds its own two dp4a per four coordinates (5 instead of 6 per query pair and word).
template <int NQ, int TRIPS, int TR, int LPR>
__global__ void __launch_bounds__(256, 2) scan_plane8_kernel(const __grid_constant__ Plane8ScanArgs p, int nstages, int smem_bytes) {
    static_assert(LPR == 32 || (TRIPS == 1 && TR % 32 == 0 && NQ == 1), "packed rows: one trip, whole rounds of 32 rows, one query");
    extern __shared__ __align__(128) unsigned char smem[];
    if (p.tail.dbg && threadIdx.x == 0) {
        if (blockIdx.x == 0) p.tail.dbg[6] = global_timer_ns();
        p.tail.dbg[32 + gridDim.x + blockIdx.x] = global_timer_ns();         // behind the CTAs' finish stamps (tail.cuh)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.tail.dbg[32 + 2 * gridDim.x + blockIdx.x] = smid;
    }
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Kp = p.Kp;
    const uint32_t row_bytes = (uint32_t)Kp;
    const uint32_t tile_bytes = (uint32_t)TR * row_bytes;
    Cand *mrg = reinterpret_cast<Cand *>(smem + (size_t)W * nstages * tile_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(mrg) + (size_t)W * 32 * sizeof(Cand));

    if (threadIdx.x == 0) {
        for (int i = 0; i < W * nstages; i++) mbar_init(smem_u32(bars + i), 1);
        mbar_fence_init();
    }
    __syncthreads();

    const u64 ntiles = (p.n + TR - 1) / TR;
    const u64 gw = (u64)blockIdx.x * W + warp, GW = (u64)gridDim.x * W;
    const uint32_t my_stage = smem_u32(smem) + (uint32_t)warp * nstages * tile_bytes;
    const uint32_t my_bar = smem_u32(bars + warp * nstages);
    auto issue = [&](u64 t, int s) {
        const u64 row0 = t * TR;
        const u64 left = p.n - row0;
        const uint32_t rows = left < (u64)TR ? (uint32_t)left : (uint32_t)TR;
        const uint32_t bytes = rows * row_bytes;
        mbar_arrive_expect_tx(my_bar + 8 * s, bytes);
        bulk_g2s(my_stage + s * tile_bytes, p.x8 + row0 * (u64)row_bytes, bytes, my_bar + 8 * s);
    };
    if (lane == 0) {
        for (int s = 0; s < nstages; s++) {
            const u64 t = gw + (u64)s * GW;
            if (t < ntiles) issue(t, s);
        }
    }

    // the query's digits, packed like the plane's bytes: lane (j = lane % LPR) owns coordinates trip * 256 + j * 8 .. + 7
    const double lo = p.par->lo, step = p.par->step;
    const int pj = lane % LPR, pg = lane / LPR;
    constexpr int CPL = LPR < 32 ? 16 : 8;              // coordinates (= bytes of the plane) per lane and trip
    uint32_t qa[NQ][TRIPS][CPL / 4], qb[NQ][TRIPS][CPL / 4];
    bool act[TRIPS];
    long long qq[NQ];                                   // sum Q_i^2 over the lane's coordinates
    unsigned long long qhash = 0;                       // of the raw bits this lane read (re-checked after pdl_wait)
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) {
        qq[qi] = 0;
#pragma unroll
        for (int t = 0; t < TRIPS; t++) {
            const int c0 = t * 256 + pj * CPL;
            act[t] = c0 < Kp;
#pragma unroll
            for (int h = 0; h < CPL / 4; h++) {
                uint32_t wa = 0, wb = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = c0 + h * 4 + j;
                    const double qc = c < p.K ? ld_cv_f64(p.q + (size_t)qi * p.ldq + c) : 0.0;
                    qhash = mix64(qhash, (unsigned long long)__double_as_longlong(qc));
                    const unsigned Q = c < p.K ? p8_quant_q(qc, lo, step) : 0u;
                    wa |= (Q >> 8) << (8 * j);
                    wb |= (Q & 255u) << (8 * j);
                    qq[qi] += (long long)Q * Q;
                }
                qa[qi][t][h] = wa;
                qb[qi][t][h] = wb;
            }
        }
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) qq[qi] += __shfl_xor_sync(FULL, qq[qi], m);     // every group of LPR lanes holds the whole query
    }
    const double c2 = (step / 256.0) * (step / 256.0);

    WarpList wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) wl[qi].reset();
    int s = 0;
    uint32_t phase = 0;
    for (u64 t = gw; t < ntiles; t += GW) {
        mbar_wait(my_bar + 8 * s, phase);
        const u64 tn = t + (u64)nstages * GW;
        if constexpr (LPR < 32) {
            constexpr int PR = 32 / LPR;               // rows side by side
            long long key[TR / 32];
            const uint32_t sa = my_stage + s * tile_bytes + (uint32_t)pg * row_bytes + (uint32_t)pj * 16;
#pragma unroll
            for (int h = 0; h < TR / 32; h++) {
                long long v[LPR];
#pragma unroll
                for (int i = 0; i < LPR; i++) {
                    const uint4 w4 = pl_lds128(sa + (uint32_t)(h * 32 + i * PR) * row_bytes);
                    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                    uint32_t s2 = 0, A = 0, Bq = 0;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        s2 = __dp4a(w[u], w[u], s2);
                        A = __dp4a(w[u], qa[0][0][u], A);
                        Bq = __dp4a(w[u], qb[0][0][u], Bq);
                    }
                    v[i] = 65536ll * (long long)s2 - 131072ll * (long long)A - 512ll * (long long)Bq;
                }
                reduce_packed<LPR>(v, lane);
                key[h] = v[0];
            }
            __syncwarp();
            if (lane == 0 && tn < ntiles) issue(tn, s);
            if (++s == nstages) {
                s = 0;
                phase ^= 1;
            }
#pragma unroll
            for (int h = 0; h < TR / 32; h++) {
                const u64 row = t * TR + (u64)(h * 32 + pj * PR + pg);
                wl[0].offer(row < p.n, (double)(key[h] + qq[0]) * c2, row, lane, p.cap);
            }
        } else {
            long long v[NQ][TR];
            const uint32_t sa = my_stage + s * tile_bytes + lane * 8;
#pragma unroll
            for (int r = 0; r < TR; r++) {
                uint32_t s2 = 0, A[NQ], B[NQ];
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) A[qi] = 0, B[qi] = 0;
#pragma unroll
                for (int tr = 0; tr < TRIPS; tr++) {
                    uint32_t w0 = 0, w1 = 0;
                    if (act[tr]) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(sa + r * row_bytes + tr * 256));
                    s2 = __dp4a(w0, w0, s2);
                    s2 = __dp4a(w1, w1, s2);
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) {
                        A[qi] = __dp4a(w0, qa[qi][tr][0], A[qi]);
                        A[qi] = __dp4a(w1, qa[qi][tr][1], A[qi]);
                        B[qi] = __dp4a(w0, qb[qi][tr][0], B[qi]);
                        B[qi] = __dp4a(w1, qb[qi][tr][1], B[qi]);
                    }
                }
#pragma unroll
                for (int qi = 0; qi < NQ; qi++)
                    v[qi][r] = 65536ll * (long long)s2 - 131072ll * (long long)A[qi] - 512ll * (long long)B[qi];
            }
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) reduce_rows<TR>(v[qi], lane);
            __syncwarp();
            // the stage is consumed: refill it before the (rare) list maintenance
            if (lane == 0 && tn < ntiles) issue(tn, s);
            if (++s == nstages) {
                s = 0;
                phase ^= 1;
            }
            const u64 row = t * TR + RowLane<TR>::row(lane);
            const bool has = RowLane<TR>::owner(lane) && row < p.n;
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) wl[qi].offer(has, (double)(v[qi][0] + qq[qi]) * c2, row, lane, p.cap);
        }
    }

    if (p.pdl) {
        // The rows are read; what follows writes.  Let the next launch on the stream start its own (read-only) scan now,
        // and touch nothing mutable -- lists, ticket, answers, the exchange -- before the launch in front of this one has
        // completed.  This scan may itself have started before that launch's completion: the one input a caller can
        // change between two calls is the query, so it is read again (past the caches) and compared.
        pdl_launch_dependents();
        pdl_wait();
        unsigned long long h2 = 0;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++)
#pragma unroll
            for (int t = 0; t < TRIPS; t++)
#pragma unroll
                for (int j = 0; j < CPL; j++) {
                    const int c = t * 256 + pj * CPL + j;
                    h2 = mix64(h2, (unsigned long long)__double_as_longlong(c < p.K ? ld_cv_f64(p.q + (size_t)qi * p.ldq + c) : 0.0));
                }
        if (__any_sync(FULL, h2 != qhash) && lane == 0) atomicOr(p.tail.ticket + 3, 1u);
    }
    const int nlists = gridDim.x;
#pragma unroll
    for (int qi = 0; qi < NQ; qi++)
        cta_merge_emit(wl[qi], mrg, W, warp, lane, p.cap, p.lists + ((size_t)qi * nlists + blockIdx.x) * p.cap);
    if (threadIdx.x == 0)                                // the tail reuses this memory (and brings its own barrier)
        for (int i = 0; i < W * nstages; i++) mbar_inval(smem_u32(bars + i));
    scan_tail(p.tail, smem, smem_bytes);
}

// ---- the plane's grid and its rows ----
__device__ __forceinline__ u64 p8_ord(double d) {
    const u64 b = (u64)__double_as_longlong(d);
    return b ^ ((u64)((long long)b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double p8_unord(u64 k) {
    return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k));
}
__global__ void plane8_range_init_kernel(Plane8Par *par) {
    par->min_ord = ~0ull;
    par->max_ord = 0ull;
}
__global__ void __launch_bounds__(256) plane8_range_kernel(const double *__restrict__ src, int ld, int K, u64 first, u64 n, Plane8Par *par) {
    u64 mn = ~0ull, mx = 0ull;
    const u64 total = n * (u64)K;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const double v = src[(first + i / K) * (u64)ld + i % K];
        if (v - v == 0.0) {                              // finite
            const u64 o = p8_ord(v);
            mn = o < mn ? o : mn;
            mx = o > mx ? o : mx;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const u64 a = __shfl_xor_sync(FULL, mn, m), b = __shfl_xor_sync(FULL, mx, m);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(&par->min_ord, mn);
        atomicMax(&par->max_ord, mx);
    }
}
__global__ void plane8_grid_kernel(Plane8Par *par) {
    double lo = 0.0, hi = 0.0;
    if (par->min_ord <= par->max_ord) {
        lo = p8_unord(par->min_ord);
        hi = p8_unord(par->max_ord);
    }
    par->lo = lo;
    par->step = hi > lo ? (hi - lo) / 255.0 : 1.0;
}
// one warp per row, four coordinates (one 32-bit word of the plane) per lane and step
__global__ void __launch_bounds__(256) plane8_build_kernel(const double *__restrict__ src, int ld, int K, int Kp, u64 first, u64 n,
                                                           const Plane8Par *__restrict__ par, unsigned char *__restrict__ dst,
                                                           unsigned long long *__restrict__ err_bits) {
    const int lane = threadIdx.x & 31;
    const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = ((u64)gridDim.x * blockDim.x) >> 5;
    const double lo = par->lo, step = par->step;
    double emax = 0.0;
    for (u64 i = gw; i < n; i += GW) {
        const u64 r = first + i;
        const double *row = src + r * (u64)ld;
        double e2 = 0.0;
        for (int c = lane * 4; c < Kp; c += 128) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (c + j < K) {
                    const double v = row[c + j];
                    const unsigned u = p8_quant_x(v, lo, step);
                    w |= u << (8 * j);
                    const double d = v - (lo + step * (double)u);
                    e2 = fma(d, d, e2);
                }
            }
            *reinterpret_cast<uint32_t *>(dst + r * (u64)Kp + c) = w;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) e2 += __shfl_xor_sync(FULL, e2, m);
        if (e2 == e2 && e2 < CUDART_INF) emax = fmax(emax, e2);    // rows with a non-finite coordinate never win (kdtree.c:139)
    }
    if (err_bits != nullptr && lane == 0 && emax > 0.0)
        atomicMax(err_bits, (unsigned long long)__double_as_longlong(sqrt(emax) * (1.0 + 1e-12)));
}

cudaError_t launch_plane8_build(const double *src, int ld, int K, int Kp, u64 first, u64 n, Plane8Par *par, bool choose_grid,
                                unsigned char *dst, unsigned long long *err_bits, int num_sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (choose_grid) {
        plane8_range_init_kernel<<<1, 1, 0, st>>>(par);
        u64 g = (n * (u64)K + 255) / 256;
        if (g > (u64)num_sms * 16) g = (u64)num_sms * 16;
        plane8_range_kernel<<<(unsigned)g, 256, 0, st>>>(src, ld, K, first, n, par);
        plane8_grid_kernel<<<1, 1, 0, st>>>(par);
    }
    u64 grid = (n + 7) / 8;
    if (grid > (u64)num_sms * 16) grid = (u64)num_sms * 16;
    plane8_build_kernel<<<(unsigned)grid, 256, 0, st>>>(src, ld, K, Kp, first, n, par, dst, err_bits);
    return cudaGetLastError();
}

bool plane8_scan_supports(int Kp) { return Kp >= 64 && Kp % 64 == 0 && Kp <= 1024; }

template <int NQ, int TRIPS, int TR, int LPR>
static cudaError_t launch_plane8_inst(const ScanTuning &t, const Plane8ScanArgs &a, cudaStream_t st) {
    const size_t row_bytes = (size_t)a.Kp;
    const int grid = a.grid > 0 ? a.grid : scan_num_lists(t, true);
    int W = t.warps < 1 ? 1 : (t.warps > 8 ? 8 : t.warps);
    int NS = t.stages < 2 ? 2 : t.stages;
    while (NS < 4 && (size_t)NS * TR * row_bytes < 8192) NS++;
    const int cps = t.ctas_per_sm > 0 ? t.ctas_per_sm : 1;
    auto need = [&](int w, int ns) { return (size_t)w * ns * TR * row_bytes + (size_t)w * 32 * sizeof(Cand) + (size_t)w * ns * 8; };
    const size_t budget = (size_t)MAX_SMEM / cps - (cps > 1 ? 1024 : 0);
    while (need(W, NS) > budget && NS > 2) NS--;
    while (need(W, NS) > budget && W > 1) W--;
    if (need(W, NS) > budget) return cudaErrorInvalidValue;
    // the tail wants every CTA's list in shared memory at once (selection path): 296 lists x cap x 16 bytes
    const size_t tail_need = a.tail.ticket ? fin_head_bytes(W) + std::max<size_t>(FIN_MIN_TBUF, (size_t)grid * a.cap * sizeof(Cand)) : 0;
    size_t smem = std::max(need(W, NS), std::min(tail_need, budget));
    // A launch that may start under the tail of the one in front of it (a.grid = one CTA less than the machine holds) is
    // placed by the hardware as slots free up, depth-first: 3 of these CTAs fit an SM by registers and shared memory, and
    // a chain of such launches ended up with 0..3 CTAs per SM and a 35-80 us spread of finish times (profiles/
    // r02_overlap_probe_before_smem_pad.txt).  Asking for more than a third of the SM's shared memory keeps it at cps per SM.
    if (a.grid > 0) smem = std::max(smem, std::min(budget, (size_t)MAX_SMEM / (cps + 1) + 1024));
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_plane8_kernel<NQ, TRIPS, TR, LPR>, smem);
    if (e != cudaSuccess) return e;
    if (a.pdl && a.tail.ticket) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)(W * 32));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, scan_plane8_kernel<NQ, TRIPS, TR, LPR>, a, NS, (int)smem);
    }
    scan_plane8_kernel<NQ, TRIPS, TR, LPR><<<grid, W * 32, smem, st>>>(a, NS, (int)smem);
    return cudaGetLastError();
}

bool plane8_scan_supports_two(int Kp) { return Kp >= 192 && Kp % 64 == 0 && Kp <= 1024; }

cudaError_t launch_scan_plane8(const ScanTuning &t, const Plane8ScanArgs &a, cudaStream_t st) {
    if (!plane8_scan_supports(a.Kp) || a.Kp < a.K || !a.x8 || !a.par || a.n == 0) return cudaErrorInvalidValue;
    if (a.nq == 2) {
        if (!plane8_scan_supports_two(a.Kp)) return cudaErrorInvalidValue;
        if (a.Kp <= 256) return launch_plane8_inst<2, 1, 16, 32>(t, a, st);
        if (a.Kp <= 512) return launch_plane8_inst<2, 2, 8, 32>(t, a, st);
        if (a.Kp <= 768) return launch_plane8_inst<2, 3, 8, 32>(t, a, st);
        return launch_plane8_inst<2, 4, 4, 32>(t, a, st);
    }
    if (a.nq != 1) return cudaErrorInvalidValue;
    if (a.Kp == 64) return launch_plane8_inst<1, 1, 128, 4>(t, a, st);
    if (a.Kp == 128) return launch_plane8_inst<1, 1, 64, 8>(t, a, st);
    if (a.Kp == 192) return launch_plane8_inst<1, 1, 32, 32>(t, a, st);
    if (a.Kp <= 256) return launch_plane8_inst<1, 1, 16, 32>(t, a, st);
    if (a.Kp <= 512) return launch_plane8_inst<1, 2, 16, 32>(t, a, st);
    if (a.Kp <= 768) return launch_plane8_inst<1, 3, 8, 32>(t, a, st);
    return launch_plane8_inst<1, 4, 8, 32>(t, a, st);
}

}  // namespace svdb
