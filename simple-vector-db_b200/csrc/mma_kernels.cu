// mma_kernels.cu -- K2: batched-query distance scan on the FP64 tensor cores (DMMA), sm_100a.
//
// Same contract as K1 (scan_kernels.cu) -- per-CTA lists of the best approximate keys for each
// query, finished by finalize_kernel's reference-order re-rank -- for batches large enough that
// one pass over the log is compute-bound rather than HBM-bound.
// Replaces src/kdtree.c:134-137 evaluated for Q queries at once.
//
//   d~(r, q) = |x_r|^2 + |q|^2 - 2 <x_r, q>
// The cross term is a [128 rows] x [64 queries] x K GEMM tile per CTA on
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no FP64 kind), 8 warps of 32x32.
// Row and query chunks of 32 coordinates are staged with 16-byte cp.async into a 3-stage ring,
// XOR-swizzled per 128-byte line so the m8n8k4 fragment loads hit the 2-wavefront minimum.
// |x_r|^2 is precomputed at insert (rownorm_kernel), |q|^2 when the batch is padded.
// After the K loop the 128x64 keys go through shared memory to the warp that owns the query:
// every warp keeps the register-resident top-32 lists (WarpList) of 8 of the 64 queries.
//
// The GEMM form cancels, so its error is ABSOLUTE: |d~ - d_ref| <= E = c (K+8) u (max|x|^2 + |q|^2);
// finalize_kernel widens its candidate window and its completeness proof by E (FinalArgs::eabs*).
#include "common.cuh"
#include "kernels.h"
#include "smem_optin.h"

namespace svdb {

constexpr int MM_ROWS = 128, MM_KC = 32, MM_STAGES = 3, MM_THREADS = 256;
constexpr int MM_X_BYTES = MM_ROWS * MM_KC * 8;        // 32 KB
constexpr int MM_DT_LD = MM_ROWS + 2;                  // keys tile [NQT][130] doubles

// Shape of one CTA for a query group of NQT queries (64 / 32 / 16): 8 warps, WR x WC of them,
// each owning MI x NI accumulator fragments of 8 rows x 8 queries.
template <int NQT>
struct MmaShape {
    static constexpr int WC = NQT == 64 ? 2 : 1;
    static constexpr int WR = 8 / WC;
    static constexpr int MI = MM_ROWS / (8 * WR);      // 4 (NQT=64) or 2
    static constexpr int NI = NQT / (8 * WC);          // 4, 4, 2
    static constexpr int QPW = NQT / 8;                // queries whose lists one warp owns
    static constexpr int Q_BYTES = NQT * MM_KC * 8;
    static constexpr int STAGE_BYTES = MM_X_BYTES + Q_BYTES;
    static constexpr int DT_BYTES = NQT * MM_DT_LD * 8;
    static constexpr int SMEM = MM_STAGES * STAGE_BYTES + DT_BYTES + NQT * 8;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// A [rows][32 doubles] tile is two 128-byte lines per row; the eight 16-byte chunks of a line are
// XOR-permuted with a per-row mask chosen so that the fragment loads below -- 128-bit loads in
// which a quarter-warp touches rows r, r+1 and four consecutive chunks each -- are conflict-free:
// the masks of rows r and r^1 differ in bit 2.
__device__ __forceinline__ int row_mask(int r) { return ((r & 1) << 2) | ((r >> 1) & 3); }
__device__ __forceinline__ uint32_t chunk_off(int r, int ch /* 16-byte chunk 0..15 of the 256-byte row */) {
    return (uint32_t)(r * 256 + (ch >> 3) * 128 + (((ch & 7) ^ row_mask(r)) << 4));
}

template <int NQT>
__global__ void __launch_bounds__(MM_THREADS, 1) scan_mma_kernel(MmaArgs p) {
    using S = MmaShape<NQT>;
    constexpr int MM_Q = NQT, MM_STAGE_BYTES = S::STAGE_BYTES, MM_DT_BYTES = S::DT_BYTES;
    constexpr int MI = S::MI, NI = S::NI, QPW = S::QPW;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = blockIdx.x % p.ngroups, stream = blockIdx.x / p.ngroups;
    const uint32_t sbase = smem_u32(smem);
    double *dt = reinterpret_cast<double *>(smem + MM_STAGES * MM_STAGE_BYTES);
    double *qn_s = reinterpret_cast<double *>(smem + MM_STAGES * MM_STAGE_BYTES + MM_DT_BYTES);

    const int q0 = group * MM_Q;                       // first query of this CTA's group
    if (tid < MM_Q) qn_s[tid] = p.qnorm[q0 + tid];
    const int nchunks = (p.K + MM_KC - 1) / MM_KC;
    const u64 ntiles = (p.n + MM_ROWS - 1) / MM_ROWS;
    const u64 my_tiles = ntiles > (u64)stream ? (ntiles - stream + p.nstreams - 1) / p.nstreams : 0;
    const u64 total = my_tiles * nchunks;              // (tile, chunk) steps of this CTA, one continuous pipeline

    WarpList wl[QPW];                                  // this warp owns queries warp*QPW .. of the group
#pragma unroll
    for (int j = 0; j < QPW; j++) wl[j].reset();

    const int wr = warp / S::WC, wc = warp % S::WC;    // warp tile: rows wr*MI*8.., queries wc*NI*8..
    const int g = lane >> 2, t4 = lane & 3;

    // producer cursor: which (tile, chunk) the next cp.async batch fetches, into which stage
    u64 p_tile = stream;
    int p_kc = 0, p_stage = 0;
    auto issue_next = [&]() {
        const uint32_t st = sbase + p_stage * MM_STAGE_BYTES;
        const int c0 = p_kc * MM_KC;
        const u64 row0 = p_tile * MM_ROWS;
#pragma unroll
        for (int i = 0; i < 8; i++) {                  // rows: 128 x 16 chunks of 16 bytes
            const int idx = tid + i * MM_THREADS;
            const int r = idx >> 4, ch = idx & 15;
            const int col = c0 + ch * 2;
            const u64 row = row0 + r;
            const bool ok = row < p.n && col < p.stride;
            const double *src = ok ? p.pts + row * (u64)p.stride + col : p.pts;
            cp_async16_zfill(st + chunk_off(r, ch), src, ok ? 16u : 0u);
        }
#pragma unroll
        for (int i = 0; i < MM_Q * 16 / MM_THREADS; i++) {   // queries: NQT x 16 chunks
            const int idx = tid + i * MM_THREADS;
            const int r = idx >> 4, ch = idx & 15;
            const int col = c0 + ch * 2;
            const bool ok = col < p.ldq;
            const double *src = ok ? p.q + (size_t)(q0 + r) * p.ldq + col : p.q;
            cp_async16_zfill(st + MM_X_BYTES + chunk_off(r, ch), src, ok ? 16u : 0u);
        }
        if (++p_kc == nchunks) {
            p_kc = 0;
            p_tile += p.nstreams;
        }
        if (++p_stage == MM_STAGES) p_stage = 0;
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < MM_STAGES - 1; s++) {
        if ((u64)s < total) issue_next();
        cp_commit();
    }

    u64 c_tile = stream;                               // consumer cursor
    int c_kc = 0, c_stage = 0;
    for (u64 it = 0; it < total; it++) {
        cp_wait<MM_STAGES - 2>();
        __syncthreads();                               // step `it` landed; the stage consumed last step is free again
        if (it + MM_STAGES - 1 < total) issue_next();  // next tile's first chunks are prefetched under this tile's last MMAs
        cp_commit();
        const uint32_t xs = sbase + c_stage * MM_STAGE_BYTES;
        const uint32_t qs = xs + MM_X_BYTES;
        // Eight coordinates per step: lane t4 fetches coordinates {2*t4, 2*t4+1} of the step with ONE
        // 128-bit load per fragment row and feeds .x to one m8n8k4 and .y to the next (the k slots of an
        // MMA may hold any 4 coordinates as long as A and B agree).
#pragma unroll
        for (int j = 0; j < MM_KC / 8; j++) {
            double2 a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = lds128(xs + chunk_off(wr * (MI * 8) + mi * 8 + g, j * 4 + t4));
#pragma unroll
            for (int ni = 0; ni < NI; ni++) b[ni] = lds128(qs + chunk_off(wc * (NI * 8) + ni * 8 + g, j * 4 + t4));
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x);
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y);
        }
        if (++c_stage == MM_STAGES) c_stage = 0;
        if (++c_kc < nchunks) continue;

        // ---- tile finished: keys -> shared memory, transposed to [query][row] ----
        const u64 row0 = c_tile * MM_ROWS;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int r = wr * (MI * 8) + mi * 8 + g;
            const double xn = row0 + r < p.n ? __ldg(p.xnorm + row0 + r) : 0.0;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int qc = wc * (NI * 8) + ni * 8 + t4 * 2 + h;
                    dt[qc * MM_DT_LD + r] = fma(-2.0, acc[mi][ni][h], xn + qn_s[qc]);
                    acc[mi][ni][h] = 0.0;
                }
            }
        }
        __syncthreads();
        // selection: each warp feeds the lists of its 8 queries; the loop-top barrier of the next step
        // orders these reads before the next tile's keys overwrite dt
#pragma unroll
        for (int j = 0; j < QPW; j++) {
            const double *col = dt + (warp * QPW + j) * MM_DT_LD;
#pragma unroll
            for (int it2 = 0; it2 < MM_ROWS / 32; it2++) {
                const int r = it2 * 32 + lane;
                wl[j].offer(row0 + r < p.n, col[r], row0 + r, lane, p.cap);
            }
        }
        c_kc = 0;
        c_tile += p.nstreams;
    }
    cp_wait<0>();

#pragma unroll
    for (int j = 0; j < QPW; j++) {
        const int q = q0 + warp * QPW + j;
        if (q < p.nq && lane < p.cap)
            p.lists[((size_t)q * p.nstreams + stream) * p.cap + lane] = Cand{wl[j].d, wl[j].seq};
    }
}

template <int NQT>
static cudaError_t launch_scan_mma_inst(const MmaArgs &a, cudaStream_t st) {
    static SmemOptIn optin;
    cudaError_t e = optin.ensure(scan_mma_kernel<NQT>, MmaShape<NQT>::SMEM);
    if (e != cudaSuccess) return e;
    scan_mma_kernel<NQT><<<a.ngroups * a.nstreams, MM_THREADS, MmaShape<NQT>::SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scan_mma(const MmaArgs &a, cudaStream_t st) {
    if (a.ngroups < 1 || a.nstreams < 1 || (a.stride & 1) || (a.ldq & 1)) return cudaErrorInvalidValue;
    switch (a.group) {
        case 64: return launch_scan_mma_inst<64>(a, st);
        case 32: return launch_scan_mma_inst<32>(a, st);
        case 16: return launch_scan_mma_inst<16>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

// queries per CTA group for a batch of nq: small batches get small groups (less padding)
int mma_group_size(size_t nq) { return nq <= 16 ? 16 : (nq <= 32 ? 32 : 64); }

// ---- |x_r|^2 at insert: one warp per row, any summation order (these feed approximate keys only) ----
__global__ void __launch_bounds__(256) rownorm_kernel(const double *__restrict__ pts, int stride, int K, u64 first, u64 n,
                                                      double *__restrict__ out, unsigned long long *max_bits) {
    const int lane = threadIdx.x & 31;
    const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = ((u64)gridDim.x * blockDim.x) >> 5;
    double mx = 0.0;
    for (u64 r = gw; r < n; r += GW) {
        const double *row = pts + (first + r) * (u64)stride;
        double s = 0.0;
        for (int i = lane; i < K; i += 32) {
            const double x = row[i];
            s = fma(x, x, s);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += shfl_xor_f64(s, m);
        if (lane == 0) out[first + r] = s;
        if (s == s && s > mx) mx = s;                  // NaN rows never win anyway; keep the bound finite
    }
    if (lane == 0 && mx > 0.0 && mx < CUDART_INF) atomicMax(max_bits, (unsigned long long)__double_as_longlong(mx));
}

cudaError_t launch_rownorm(const double *pts, int stride, int K, u64 first, u64 n, double *out, unsigned long long *max_bits,
                           int num_sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const u64 warps = n;
    u64 grid = (warps + 7) / 8;
    if (grid > (u64)num_sms * 8) grid = (u64)num_sms * 8;
    rownorm_kernel<<<(unsigned)grid, 256, 0, st>>>(pts, stride, K, first, n, out, max_bits);
    return cudaGetLastError();
}

// Pad a batch of queries to [nq_pad][ldp] (zeros beyond K and beyond nq) and compute |q|^2.
__global__ void __launch_bounds__(128) prep_queries_kernel(const double *__restrict__ src, int ldq, int K, int nq,
                                                           double *__restrict__ dst, int ldp, double *__restrict__ qnorm) {
    const int q = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < ldp; c += blockDim.x) {
        const double v = (q < nq && c < K) ? src[(size_t)q * ldq + c] : 0.0;
        dst[(size_t)q * ldp + c] = v;
        s = fma(v, v, s);
    }
    __shared__ double part[4];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += shfl_xor_f64(s, m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) qnorm[q] = part[0] + part[1] + part[2] + part[3];
}

cudaError_t launch_prep_queries(const double *src, int ldq, int K, int nq, int nq_pad, double *dst, int ldp, double *qnorm,
                                cudaStream_t st) {
    if (nq_pad == 0) return cudaSuccess;
    prep_queries_kernel<<<nq_pad, 128, 0, st>>>(src, ldq, K, nq, dst, ldp, qnorm);
    return cudaGetLastError();
}

}  // namespace svdb
