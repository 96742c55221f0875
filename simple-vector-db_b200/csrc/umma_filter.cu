// umma_filter.cu -- K10: batched-query distance keys on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as K2 (mma_kernels.cu): per-CTA lists of the best APPROXIMATE keys for every query of a batch,
// finished by finalize_kernel's reference-order re-rank (kdtree.c:134-137) and completeness proof.  K2 forms the
// keys on the FP64 tensor cores (DMMA, 37 TFLOP/s); tcgen05 has no FP64 kind, but the keys only have to be good
// enough to keep every row that could be in the top-k, so K10 forms them from a SPLIT-BF16 copy of the log:
//
//   x = xh + xl + r,  xh = bf16(x), xl = bf16(fl32(x) - xh)   (two planes of [n][Kp] bf16: hi, lo), |r| <= (2^-16 + 2^-24) |x|        (same for q)
//   <x, q> ~ <xh, qh> + <xh, ql> + <xl, qh>            three kind::f16 UMMAs per 16 coordinates, fp32 accumulators in TMEM
//   d~(r, q) = |x_r|^2 + |q|^2 - 2 <x_r, q>            |x|^2, |q|^2 from the fp64 rows (precomputed, rounded to fp32 once)
//
// Error: the dropped terms are bounded by 3.1 * 2^-16 sum|x_i q_i|.  The fp32 accumulation of the 3K exact bf16 products in
// K/16 steps of three chained instructions was MEASURED, not modelled (scripts/umma_accumulator_probe.py,
// profiles/r02_umma_accumulator_probe.jsonl, pinned by tests/test_gpu_umma.py): rows and queries made of +-powers of two
// (no lo plane, exact products) in truncation-adversarial patterns -- one product of magnitude 1 and K-1 same-sign
// products of 2^-30 .. 2^-16, big products at the head of every instruction, alternating signs, ramps -- lose at most
// 1.5 * 2^-24 sum|x_i q_i| per chained instruction (K = 64: nothing at all; the loss appears once the running sum dwarfs the
// addends, as for an adder that aligns the 16 products and the accumulator to the largest exponent, keeps a few guard
// bits and truncates: <= 1 ulp = 2 * 2^-24 of the running sum per instruction).  The budget is FOUR times that model,
// 8 * 2^-24 = 2^-21 per instruction, i.e. (3K/16) 2^-21 sum|x_i q_i| per key (5.3 x the measured worst case); the three
// fp32 roundings of the key add 2^-21 (|x|^2 + |q|^2).  With
// sum|x_i q_i| <= |x||q| <= (|x|^2 + |q|^2)/2 the key error is ABSOLUTE, E = coef (max|x|^2 + |q|^2) like K2's, and
// finalize_kernel widens its re-rank window and its proof by it.  Non-finite keys (overflow of bf16/fp32 on huge
// values) are kept as candidates, never dropped, so extreme data ends in the exact fallback instead of a wrong answer.
//
// Kernel: one persistent CTA per SM; CTA c serves query group c % ngroups (bn = 64/128/256 queries) over row tiles
// stream, stream + nstreams, ... (CTAs of one stream run together, so a row tile comes from HBM once and from L2 after).
//   warp 0   : TMA producer  -- four 2-D tiled loads per stage (row hi/lo planes 128 x 64, query hi/lo planes bn x 64,
//              SWIZZLE_128B) completing on the stage's mbarrier
//   warp 1   : MMA issuer    -- one lane issues 12 tcgen05.mma (M=128, N=bn, K=16) per stage into one of two TMEM
//              accumulator stages; tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 4-11: epilogue     -- tcgen05.ld 32 lanes x 32 columns, |x|^2 - 2 acc against the query's threshold (tau - |q|^2, in
//              shared memory); every lane (= row) collects the pass bits of its 32 keys, a survivor's own lane reserves its
//              slot in the (CTA, query) buffer in global memory with one shared-memory atomic and stores it there;
//              thresholds come from the whole query group: every CTA publishes the smallest key it has kept per query,
//              the cap-th smallest of those (keys of distinct rows) bounds the group's cap-th smallest key and is handed
//              round through one u32 per query in global memory (atomicMin / looked at every few tiles);
//              between two CTA barriers -- after every tile at first, then after every fourth -- a warp prunes the buffers
//              it owns that are filling up to the `cap` smallest (bitwise selection of the cap-th smallest key, uf_prune).
//              A buffer that overflows between two checks is emitted as unprovable (-> SVDB_CAND_UNSAFE), never truncated.
#include <cuda.h>
#include <cuda_bf16.h>
#include <float.h>

#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "smem_optin.h"
#include "umma_select.cuh"

namespace svdb {

constexpr int UF_M = 128;                       // rows per tile = TMEM lanes
constexpr int UF_NMAX = 256;                    // queries per CTA group, at most
constexpr int UF_KC = 64;                       // bf16 coordinates per stage row = one 128-byte swizzle row
constexpr int UF_STAGES = 2;                     // ring depth when the queries are streamed with the rows
constexpr int UF_MAX_STAGES = 6;                 // ... and at most when they are resident (rows only in the ring)
constexpr int UF_A_BYTES = UF_M * 128;          // one plane of a row tile
constexpr int UF_B_BYTES = UF_NMAX * 128;       // one plane of a query tile
constexpr int UF_STAGE_BYTES = 2 * UF_A_BYTES + 2 * UF_B_BYTES;     // 96 KB: row hi/lo + query hi/lo of one 64-coordinate chunk
constexpr int UF_RING_BYTES = UF_STAGES * UF_STAGE_BYTES;           // 192 KB of operand staging either way
constexpr int UF_THREADS = 384;                  // TMA warp, MMA warp, two idle, eight epilogue warps
constexpr int UF_TAIL = 256 + 4 * UF_NMAX * 4;                      // barriers + tmem pointer, cnt / tau / qn / smallest key kept
constexpr int UF_SMEM = UF_RING_BYTES + UF_TAIL + 1024;             // + slack to align the stages to 1024 bytes

// Short kd-points (config 2 / 5 rows: K = 128): a row tile is only one or two 64-coordinate chunks, so with the queries
// streamed next to the rows (96 KB per chunk, two stages) the ring is ONE tile deep and every tile waits out a TMA round
// trip -- measured 6 us per 128-row tile at 1M x 128 where the tensor work needs 0.4 and HBM 1.4.  When the query planes
// of the CTA's group fit next to at least two row stages they are loaded ONCE and stay resident: the ring then carries
// rows only (32 KB per chunk, up to six stages deep) and the query planes are no longer re-read from L2 for every tile.
__host__ __device__ inline int uf_resident_stages(int bn, int Kp) {
    const int qbytes = (Kp / UF_KC) * 2 * bn * 128;
    const int left = UF_RING_BYTES - qbytes;
    if (left < 2 * 2 * UF_A_BYTES) return 0;                        // not resident: stream the queries (UF_STAGES stages)
    const int st = left / (2 * UF_A_BYTES);
    return st > UF_MAX_STAGES ? UF_MAX_STAGES : st;
}

int umma_kpad(int K) { return (K + UF_KC - 1) / UF_KC * UF_KC; }
int umma_group_size(size_t nq) { return nq <= 64 ? 64 : (nq <= 128 ? 128 : 256); }
int umma_resident_stages(int bn, int Kp) { return uf_resident_stages(bn, Kp); }
size_t umma_buf_bytes(int ngroups, int nstreams, int bn) { return (size_t)ngroups * nstreams * bn * UF_BUF * 8; }
double umma_eabs_coef(int K) { return 3.2 * ldexp(1.0, -16) + (3.0 * K / 16.0) * ldexp(1.0, -21) + 8.0 * ldexp(1.0, -20); }

// ---- PTX wrappers (tcgen05 / TMA); the mbarrier ones live in common.cuh ----
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives lane (base lane + t), register i = column i
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile stored as rows of 128 bytes (64 bf16), SWIZZLE_128B:
// 8-row groups 1024 bytes apart (SBO), version 1 (sm_100), layout type 2.  Stepping 16 coordinates along K inside
// the 128-byte row advances the start address by 32 bytes (the swizzle is applied to the address bits by the hardware).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// Instruction descriptor: D fp32 (bits 4-5 = 1), A and B bf16 (bits 7-9 / 10-12 = 1), both K-major (bits 15, 16 = 0),
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UF_M >> 4) << 24);
}

__global__ void __launch_bounds__(UF_THREADS, 1)
umma_filter_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                   const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql, UmmaArgs p) {
    extern __shared__ unsigned char uf_smem_raw[];
    const uint32_t raw = smem_u32(uf_smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;                   // SWIZZLE_128B tiles want 1024-byte alignment
    unsigned char *sgen = uf_smem_raw + (sbase - raw);
    unsigned char *tail = sgen + UF_RING_BYTES;
    const uint32_t tail_u = sbase + UF_RING_BYTES;
    // barriers: full[S] (TMA -> MMA), empty[S] (MMA -> TMA), tfull[2] (MMA -> epilogue), tempty[2] (epilogue -> MMA), qfull
    const uint32_t bar_full = tail_u, bar_empty = tail_u + 64, bar_tfull = tail_u + 128, bar_tempty = tail_u + 144, bar_q = tail_u + 160;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 176);
    unsigned *cnt_s = reinterpret_cast<unsigned *>(tail + 256);
    float *thr_s = reinterpret_cast<float *>(tail + 256 + UF_NMAX * 4);   // per query: tau - |q|^2, tau = cap-th smallest key so far
    float *qn_s = reinterpret_cast<float *>(tail + 256 + 2 * UF_NMAX * 4);
    uint32_t *lmin_s = reinterpret_cast<uint32_t *>(tail + 256 + 3 * UF_NMAX * 4);   // per query: smallest key this CTA has kept (uf_ord)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bn = p.bn;
    const int group = blockIdx.x % p.ngroups, stream = blockIdx.x / p.ngroups;
    const int q0 = group * bn;
    const int nchunks = p.Kp / UF_KC;
    const u64 ntiles = (p.n + UF_M - 1) / UF_M;
    // option umma.debug_keys: CTA 0 also leaves cycle counts of its three roles in 16 floats behind the key dump (128 x UF_NMAX)
    const bool prof = p.dbg_keys != nullptr && blockIdx.x == 0;
    long long clk[6] = {0, 0, 0, 0, 0, 0};
    // resident queries (p.qres stages of rows behind nchunks x [hi | lo] query tiles) or queries streamed with the rows
    const int nst = p.qres > 0 ? p.qres : UF_STAGES;
    const uint32_t qchunk_bytes = 2u * (uint32_t)bn * 128u;
    const uint32_t ring0 = p.qres > 0 ? sbase + (uint32_t)nchunks * qchunk_bytes : sbase;
    const uint32_t stage_bytes = p.qres > 0 ? 2u * UF_A_BYTES : (uint32_t)UF_STAGE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 8);            // one arrival per epilogue warp
        }
        mbar_init(bar_q, 1);
        mbar_fence_init();
    }
    if (warp == 1) {                                     // the whole warp allocates all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < UF_NMAX; j += UF_THREADS) {
        cnt_s[j] = 0;
        lmin_s[j] = 0xffffffffu;
        thr_s[j] = CUDART_INF_F;
        qn_s[j] = j < bn ? (float)p.qnorm[q0 + j] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (p.qres > 0) {                            // the group's query planes, once
                mbar_arrive_expect_tx(bar_q, (uint32_t)nchunks * qchunk_bytes);
                for (int kc = 0; kc < nchunks; kc++) {
                    tma_load_2d(sbase + kc * qchunk_bytes, &map_qh, kc * UF_KC, q0, bar_q);
                    tma_load_2d(sbase + kc * qchunk_bytes + (uint32_t)bn * 128u, &map_ql, kc * UF_KC, q0, bar_q);
                }
            }
            const uint32_t stage_tx = p.qres > 0 ? 2u * UF_A_BYTES : 2 * UF_A_BYTES + 2 * (uint32_t)bn * 128u;
            int stage = 0;
            uint32_t phase = 0;
            const long long cstart = prof ? clock64() : 0;
            for (u64 tile = stream; tile < ntiles; tile += p.nstreams) {
                const int row0 = (int)(tile * UF_M);
                for (int kc = 0; kc < nchunks; kc++) {
                    const long long c0 = prof ? clock64() : 0;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (prof) clk[0] += clock64() - c0;
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t st = ring0 + stage * stage_bytes;
                    mbar_arrive_expect_tx(full, stage_tx);
                    tma_load_2d(st, &map_xh, kc * UF_KC, row0, full);                                  // row tile, hi plane
                    tma_load_2d(st + UF_A_BYTES, &map_xl, kc * UF_KC, row0, full);                     //           lo plane
                    if (p.qres == 0) {
                        tma_load_2d(st + 2 * UF_A_BYTES, &map_qh, kc * UF_KC, q0, full);               // queries, hi plane
                        tma_load_2d(st + 2 * UF_A_BYTES + UF_B_BYTES, &map_ql, kc * UF_KC, q0, full);  //          lo plane
                    }
                    if (++stage == nst) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (prof) {
                p.dbg_keys[128 * UF_NMAX + 0] = (float)clk[0] * 1e-3f;                    // producer: waiting for a free smem stage
                p.dbg_keys[128 * UF_NMAX + 1] = (float)(clock64() - cstart) * 1e-3f;      // producer: whole loop
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(bn);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            if (p.qres > 0) {
                mbar_wait(bar_q, 0);
                tc_fence_after();
            }
            const long long cstart = prof ? clock64() : 0;
            for (u64 tile = stream; tile < ntiles; tile += p.nstreams) {
                long long c0 = prof ? clock64() : 0;
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);      // the epilogue has drained this accumulator stage
                if (prof) clk[0] += clock64() - c0;
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)as * UF_NMAX;
                for (int kc = 0; kc < nchunks; kc++) {
                    c0 = prof ? clock64() : 0;
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (prof) clk[1] += clock64() - c0;
                    tc_fence_after();
                    const uint32_t st = ring0 + stage * stage_bytes;
                    const uint32_t qb = p.qres > 0 ? sbase + kc * qchunk_bytes : st + 2 * UF_A_BYTES;
                    const uint32_t qlo_off = p.qres > 0 ? (uint32_t)bn * 128u : (uint32_t)UF_B_BYTES;
#pragma unroll
                    for (int j = 0; j < UF_KC / 16; j++) {
                        const uint64_t xh = umma_desc(st + j * 32), xl = umma_desc(st + UF_A_BYTES + j * 32);
                        const uint64_t qh = umma_desc(qb + j * 32);
                        const uint64_t ql = umma_desc(qb + qlo_off + j * 32);
                        tc_mma(acc, xh, ql, idesc, (kc | j) != 0);      // the small products first
                        tc_mma(acc, xl, qh, idesc, 1);
                        tc_mma(acc, xh, qh, idesc, 1);
                    }
                    tc_commit(bar_empty + 8 * stage);            // smem stage free once these MMAs have read it
                    if (++stage == nst) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(bar_tfull + 8 * as);                   // accumulator complete
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
            if (prof) {
                p.dbg_keys[128 * UF_NMAX + 2] = (float)clk[0] * 1e-3f;                    // MMA issuer: waiting for the epilogue (tempty)
                p.dbg_keys[128 * UF_NMAX + 3] = (float)clk[1] * 1e-3f;                    // MMA issuer: waiting for loads (full)
                p.dbg_keys[128 * UF_NMAX + 4] = (float)(clock64() - cstart) * 1e-3f;      // MMA issuer: whole loop
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: keys, thresholds, candidate buffers =====================
        // Eight warps: warp % 4 is the TMEM lane quarter (32 rows) a warp may read, (warp - 4) / 4 the half of the
        // group's query columns it looks at.
        const int wi = warp - 4, ew = warp & 3, half = wi >> 2;
        const unsigned below = (1u << lane) - 1u;
        UfEntry *bufs = reinterpret_cast<UfEntry *>(p.bufs) + (size_t)blockIdx.x * bn * UF_BUF;
        const int c_lo = half * (bn / 64), c_hi = c_lo + bn / 64;  // 32-column chunks of this warp
        int as = 0, it = 0;
        uint32_t aphase = 0;
        // |x|^2 of this lane's row of the NEXT tile and the group's shared thresholds are fetched before the wait for the
        // accumulator, not after it: two global round trips per tile came off the epilogue's critical path
        // (an ncu capture at 1M x 128 showed the F2F behind the norm load and the threshold load as the top long-scoreboard
        // stalls, and the eight warps wait for each other twice per tile; profiles/r02_K10_role_cycles_before.jsonl)
        const int jmine = wi + 8 * lane;                  // bn / 8 <= 32 queries per warp: lane t looks after query wi + 8 t
        const bool mineok = lane < bn / 8;
        // Thresholds from the whole group (round 2): every CTA of a query group publishes, per query, the smallest key it has
        // kept (gmin).  Those are keys of nstreams DISTINCT rows, so the cap-th smallest of them bounds the group's cap-th
        // smallest key from above -- and sits near the quantile 1 / (rows one CTA has seen), where a CTA's own cap-th smallest
        // key (what gtau carried before) sits near cap / (rows one CTA has seen): ~cap times fewer survivors to append,
        // sort and prune.  CTA `stream` looks after the queries stream, stream + nstreams, ... of its group, one per warp.
        const bool use_gmin = p.gmin != nullptr && p.nstreams >= p.cap;
        uint32_t pub = 0xffffffffu;                        // what this lane last published for query jmine
        bool ovf_mine = false;                            // query jmine's buffer overflowed between two checks: unprovable
        const int jsel = stream + p.nstreams * wi;        // the query whose group threshold this warp refreshes
        double xn_next = 0.0;
        {
            const u64 r0 = (u64)stream * UF_M + ew * 32 + lane;
            if (stream < ntiles && r0 < p.n) xn_next = __ldg(p.xnorm + r0);
        }
        for (u64 tile = stream; tile < ntiles; tile += p.nstreams) {
            // The bookkeeping between the two CTA barriers below (prune check, publishing minima, threshold refresh) runs after
            // every tile while the thresholds are still loose, then after every fourth: the barriers were 2.1-4.4 of the
            // 5.6-9.4 kcycles a tile cost at kd_dim 128 (profiles/r02_K10_role_cycles_ab.jsonl).  A buffer that nevertheless
            // fills up between two checks (rows arriving in order of decreasing distance, say) loses nothing silently: its
            // query is emitted as unprovable (ovf_mine below) and re-answered by the caller's next rung.
            const bool sect = p.sparse_checks == 0 || it < 8 || (it & 3) == 3;
            const bool refresh = sect;
            // the group threshold (a 32-step search, ~2.7 kcycles) tightens with the logarithm of the rows seen: computed after
            // tiles 0..7, then whenever the tile count reaches a power of two, and every 64th tile
            const bool select = sect && (p.sparse_checks == 0 ? (it < 16 || (it & 3) == 0) : (it < 8 || ((it + 1) & it) == 0 || (it & 63) == 63));
            uint32_t g = 0xffffffffu;
            if (refresh && mineok) g = __ldcg(p.gtau + q0 + jmine);
            const u64 row = tile * UF_M + ew * 32 + lane;
            const bool ok = row < p.n;
            const float xn = ok ? (float)xn_next : 0.f;
            {
                const u64 rn = row + (u64)p.nstreams * UF_M;
                xn_next = (tile + p.nstreams < ntiles && rn < p.n) ? __ldg(p.xnorm + rn) : 0.0;
            }
            long long c0 = prof ? clock64() : 0;
            mbar_wait(bar_tfull + 8 * as, aphase);
            if (prof) clk[0] += clock64() - c0, c0 = clock64();
            tc_fence_after();
            const bool dbg = p.dbg_keys != nullptr && blockIdx.x == 0 && tile == (u64)stream;
            for (int c = c_lo; c < c_hi; c++) {
                uint32_t v[32];
                long long c1 = prof ? clock64() : 0;
                tc_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * UF_NMAX + c * 32), v);
                float th[32];                                     // thresholds of these 32 queries (constant during the tile)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 t4 = reinterpret_cast<const float4 *>(thr_s + c * 32)[i];
                    th[4 * i] = t4.x, th[4 * i + 1] = t4.y, th[4 * i + 2] = t4.z, th[4 * i + 3] = t4.w;
                }
                tc_wait_ld();
                if (prof) clk[3] += clock64() - c1, c1 = clock64();
                // key < tau  <=>  |x|^2 - 2 acc < tau - |q|^2 =: thr (kept conservative, see uf_thr); NaN passes.
                // Every lane (= row) collects the pass bits of its own 32 keys: no vote per key.  (The first version took one
                // ballot per key and transposed the survivors to the query's lane with shuffles: 5 instructions per key on
                // the fast path and ~600 cycles per chunk for the appends, which are NOT rare -- the threshold only tightens
                // when a buffer has taken 128 new entries, so ~60 keys per tile survive at any store size;
                // profiles/r02_K10_role_cycles_before.jsonl.)
                unsigned pm = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const bool pass = !(fmaf(-2.f, __uint_as_float(v[i]), xn) >= th[i]);
                    pm |= pass ? (1u << i) : 0u;
                }
                if (!ok) pm = 0;
                if (prof) clk[4] += clock64() - c1, c1 = clock64();
                if (__any_sync(FULL, pm != 0)) {
                    // a survivor's own lane reserves its slot (one shared-memory atomic per survivor) and stores it
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        if ((pm >> i) & 1u) {
                            const int j = c * 32 + i;
                            float key = fmaf(-2.f, __uint_as_float(v[i]), xn + qn_s[j]);
                            if (!(fabsf(key) <= FLT_MAX)) key = -FLT_MAX;       // NaN / inf: keep it, never drop it
                            const unsigned pos = atomicAdd(&cnt_s[j], 1u);
                            if (pos < (unsigned)UF_BUF) bufs[(size_t)j * UF_BUF + pos] = UfEntry{key, (uint32_t)row};
                            atomicMin(lmin_s + j, uf_ord(key));
                        }
                    }
                    __syncwarp();
                }
                if (prof) clk[5] += clock64() - c1;
                if (dbg) {
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        p.dbg_keys[(size_t)(ew * 32 + lane) * bn + c * 32 + i] = fmaf(-2.f, __uint_as_float(v[i]), xn + qn_s[c * 32 + i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (prof) clk[1] += clock64() - c0, c0 = clock64();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);      // this warp is done with the accumulator stage
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
            // ---- prune the buffers that could overflow before the next check (warp wi owns queries wi, wi+8, ...) ----
            if (sect) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            {
                // g: the smallest cap-th key any CTA of this query group had published when this tile began: an upper bound
                // of the group's cap-th smallest key, so keys above it are nobody's candidates (the CTAs converge as one)
                const unsigned cmine = mineok ? cnt_s[jmine] : 0u;
                if (cmine > (unsigned)UF_BUF) ovf_mine = true;           // entries were lost
                unsigned m = __ballot_sync(FULL, cmine > (unsigned)(p.sparse_checks && it >= 8 ? UF_BUF / 4 : UF_BUF - UF_M));
                while (m) {
                    const int j = wi + 8 * (__ffs(m) - 1);
                    m &= m - 1;
                    const unsigned cnt = min(cnt_s[j], (unsigned)UF_BUF);
                    float tau;
                    bool dropped;
                    const unsigned kept = uf_prune(bufs + (size_t)j * UF_BUF, cnt, p.cap, lane, tau, dropped);
                    __syncwarp();
                    if (lane == 0) {
                        cnt_s[j] = kept;
                        if (dropped) {
                            thr_s[j] = fminf(thr_s[j], uf_thr(tau, qn_s[j]));   // from now on only keys below the cap-th smallest
                            atomicMin(p.gtau + q0 + j, uf_ord(tau));
                        }
                    }
                }
                __syncwarp();
                if (use_gmin) {
                    if (mineok) {
                        const uint32_t lm = lmin_s[jmine];
                        if (lm < pub) {
                            p.gmin[(size_t)(q0 + jmine) * p.nstreams + stream] = lm;
                            pub = lm;
                        }
                    }
                    if (select && jsel < bn) {
                        // cap-th smallest of the group's published minima of query jsel: 32-step search on the ordered bits
                        uint32_t val[5];
                        int have = 0;
#pragma unroll
                        for (int t = 0; t < 5; t++) {
                            const int s2 = lane + 32 * t;
                            val[t] = s2 < p.nstreams ? __ldcg(p.gmin + (size_t)(q0 + jsel) * p.nstreams + s2) : 0xffffffffu;
                            have += val[t] != 0xffffffffu;
                        }
                        have = __reduce_add_sync(FULL, have);
                        if (have >= p.cap) {
                            uint32_t x = 0;
                            for (int bit = 31; bit >= 0; bit--) {
                                const uint32_t cand = x | (1u << bit);
                                int c2 = 0;
#pragma unroll
                                for (int t = 0; t < 5; t++) c2 += val[t] < cand;
                                if (__reduce_add_sync(FULL, c2) < p.cap) x = cand;
                            }
                            // everybody (this CTA included: thr_s[jsel] belongs to another warp) picks it up with the next look at gtau
                            if (lane == 0) atomicMin(p.gtau + q0 + jsel, x);
                        }
                    }
                    __syncwarp();
                }
                if (g != 0xffffffffu) thr_s[jmine] = fminf(thr_s[jmine], uf_thr(uf_unord(g), qn_s[jmine]));
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (prof) clk[2] += clock64() - c0;
            it++;
        }
        if (prof && wi == 0 && lane == 0) {
            p.dbg_keys[128 * UF_NMAX + 5] = (float)clk[0] * 1e-3f;                        // epilogue warp 4: waiting for the accumulator
            p.dbg_keys[128 * UF_NMAX + 6] = (float)clk[1] * 1e-3f;                        //                 keys, thresholds, appends
            p.dbg_keys[128 * UF_NMAX + 7] = (float)clk[2] * 1e-3f;                         //                 barriers + pruning
            p.dbg_keys[128 * UF_NMAX + 8] = (float)it;
            p.dbg_keys[128 * UF_NMAX + 9] = (float)clk[3] * 1e-3f;                         //   of the keys part: tcgen05.ld + thresholds from smem
            p.dbg_keys[128 * UF_NMAX + 10] = (float)clk[4] * 1e-3f;                         //                     compare, 32 keys per lane
            p.dbg_keys[128 * UF_NMAX + 11] = (float)clk[5] * 1e-3f;                         //                     appending survivors
        }
        // ---- emit: the cap smallest keys of every query of the group, ascending, for finalize_kernel ----
        if (mineok && cnt_s[jmine] > (unsigned)UF_BUF) ovf_mine = true;
        for (int j = wi; j < bn; j += 8) {
            UfEntry *b = bufs + (size_t)j * UF_BUF;
            float tau;
            bool dropped;
            const unsigned kept = uf_prune(b, min(cnt_s[j], (unsigned)UF_BUF), p.cap, lane, tau, dropped);
            __syncwarp();
            // a buffer that lost entries cannot vouch for anything: all its keys become -FLT_MAX, which makes the list's bound
            // -FLT_MAX and the completeness proof of this query fail in finalize (SVDB_CAND_UNSAFE -> re-answered)
            const bool ovf = __shfl_sync(FULL, ovf_mine, (j - wi) >> 3);
            u64 v = ~0ull;                                        // (key, row) in one word; empty slots sort last
            if ((unsigned)lane < kept) {
                const UfEntry e = b[lane];
                v = ((u64)uf_ord(ovf ? -FLT_MAX : e.key) << 32) | e.row;
            }
            v = uf_sort32(v, lane);
            if (lane < p.cap) {
                Cand c{CUDART_INF, SEQ_NONE};
                if (v != ~0ull) c = Cand{(double)uf_unord((uint32_t)(v >> 32)), v & 0xffffffffull};
                p.lists[((size_t)(q0 + j) * p.nstreams + stream) * p.cap + lane] = c;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- fp64 rows -> the two bf16 planes of the shadow, [n][Kp] each (zeros beyond K): hi = bf16(fl32(x)),
// lo = bf16(fl32(x) - hi); either destination may be NULL (the lo plane is only built once K10 / K11 ask for it).
// One warp per row.  err_bits (may be NULL): running maximum over the rows of |x - hi|_2, the
// error of the HI plane alone, formed in fp64 and rounded up -- what K12's completeness proof needs (plane_scan.cu).
// Rows with a non-finite coordinate are left out of it: their distance is never finite, so they can never win
// (kdtree.c:139) and no bound on them is needed. ----
__global__ void __launch_bounds__(256) split_bf16_kernel(const double *__restrict__ src, int ld, int K, int Kp, u64 first, u64 n,
                                                         __nv_bfloat16 *__restrict__ dst_hi, __nv_bfloat16 *__restrict__ dst_lo,
                                                         unsigned long long *__restrict__ err_bits) {
    const int lane = threadIdx.x & 31;
    const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = ((u64)gridDim.x * blockDim.x) >> 5;
    double emax = 0.0;
    for (u64 i = gw; i < n; i += GW) {
        const u64 r = first + i;
        const double *row = src + r * (u64)ld;
        double e2 = 0.0;
        for (int c = lane * 2; c < Kp; c += 64) {
            const double v0 = c < K ? row[c] : 0.0, v1 = c + 1 < K ? row[c + 1] : 0.0;
            const float f0 = __double2float_rn(v0), f1 = __double2float_rn(v1);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(f0 - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(f1 - __bfloat162float(h1));
            if (dst_hi) *reinterpret_cast<__nv_bfloat162 *>(dst_hi + r * (u64)Kp + c) = __nv_bfloat162(h0, h1);
            if (dst_lo) *reinterpret_cast<__nv_bfloat162 *>(dst_lo + r * (u64)Kp + c) = __nv_bfloat162(l0, l1);
            const double d0 = v0 - (double)__bfloat162float(h0), d1 = v1 - (double)__bfloat162float(h1);
            e2 = fma(d0, d0, e2);
            e2 = fma(d1, d1, e2);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) e2 += __shfl_xor_sync(FULL, e2, m);
        if (e2 == e2 && e2 < CUDART_INF) emax = fmax(emax, e2);
    }
    if (err_bits != nullptr && lane == 0 && emax > 0.0) {
        const double e = sqrt(emax) * (1.0 + 1e-12);           // non-negative doubles order like their bit patterns
        atomicMax(err_bits, (unsigned long long)__double_as_longlong(e));
    }
}

cudaError_t launch_split_bf16(const double *src, int ld, int K, int Kp, u64 first, u64 n, uint16_t *dst_hi, uint16_t *dst_lo,
                              unsigned long long *err_bits, int num_sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    u64 grid = (n + 7) / 8;
    if (grid > (u64)num_sms * 16) grid = (u64)num_sms * 16;
    split_bf16_kernel<<<(unsigned)grid, 256, 0, st>>>(src, ld, K, Kp, first, n, reinterpret_cast<__nv_bfloat16 *>(dst_hi),
                                                      reinterpret_cast<__nv_bfloat16 *>(dst_lo), err_bits);
    return cudaGetLastError();
}

// ---- host: tensor maps (driver entry point resolved at run time, like arena.cu) and the launch ----
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled(std::string &why) {
    static EncodeTiledFn fn = nullptr;
    static std::string err;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr)
            err = std::string("cannot resolve driver entry point cuTensorMapEncodeTiled: ") + cudaGetErrorString(e);
        else
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    if (!fn) why = err;
    return fn;
}
// one plane: [rows][Kp] bf16, box = 64 coordinates x box_rows rows, 128-byte swizzle, zeros out of bounds
bool make_map(CUtensorMap *m, const void *base, u64 rows, int Kp, int box_rows, std::string &why) {
    EncodeTiledFn fn = encode_tiled(why);
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)UF_KC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
        return false;
    }
    return true;
}
}  // namespace

cudaError_t launch_umma_filter(const UmmaArgs &a, cudaStream_t st, std::string *why) {
    std::string err;
    if (a.qres != 0 && a.qres != uf_resident_stages(a.bn, a.Kp)) {
        if (why) *why = "umma filter: bad resident-query setting";
        return cudaErrorInvalidValue;
    }
    if ((a.bn != 64 && a.bn != 128 && a.bn != 256) || a.ngroups < 1 || a.nstreams < 1 || a.Kp % UF_KC || a.Kp < a.K || a.cap < 1 ||
        a.cap > 32 || a.n == 0 || a.n >= (1ull << 31)) {
        if (why) *why = "umma filter: bad launch shape";
        return cudaErrorInvalidValue;
    }
    CUtensorMap mxh, mxl, mqh, mql;
    const u64 nqp = (u64)a.ngroups * a.bn;
    if (!make_map(&mxh, a.xhi, a.n, a.Kp, UF_M, err) || !make_map(&mxl, a.xlo, a.n, a.Kp, UF_M, err) ||
        !make_map(&mqh, a.qhi, nqp, a.Kp, a.bn, err) || !make_map(&mql, a.qlo, nqp, a.Kp, a.bn, err)) {
        if (why) *why = err;
        return cudaErrorNotSupported;
    }
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(umma_filter_kernel, UF_SMEM);
    if (e != cudaSuccess) return e;
    umma_filter_kernel<<<a.ngroups * a.nstreams, UF_THREADS, UF_SMEM, st>>>(mxh, mxl, mqh, mql, a);
    return cudaGetLastError();
}

}  // namespace svdb
