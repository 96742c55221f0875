// scan_common.cuh -- pieces shared by the distance scans (scan_kernels.cu, plane_scan.cu): the CTA-level merge of the
// per-warp candidate lists and the row-halving warp reduction.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "smem_optin.h"

namespace svdb {

constexpr int MAX_SMEM = 232448;  // 227 KB opt-in limit per CTA on sm_100

// ---- CTA-level merge of the per-warp lists, one query -------------------------------
__device__ __forceinline__ void cta_merge_emit(WarpList &mine, Cand *mrg, int W, int warp, int lane, int cap,
                                               Cand *out) {
    mrg[warp * 32 + lane] = Cand{mine.d, mine.seq};
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < W; w++) {
            const Cand c = mrg[w * 32 + lane];
            mine.offer(c.seq != SEQ_NONE, c.d, c.seq, lane, cap);
        }
        if (lane < cap) out[lane] = Cand{mine.d, mine.seq};
    }
    __syncthreads();
}

// ---- reduce TR per-lane partial sums over the warp with TR-1 + (5 - log2 TR) shuffles ------------
// A plain butterfly costs 5 shuffles (10 SHFL.32) per row, which is what bounds short rows (measured:
// 0.23-0.65 x HBM peak for kd_dim 17..48).  Here the first log2(TR) steps HALVE the set instead: a lane
// passes the rows it gives up to its partner and adds what it receives to the rows it keeps.  Every row
// is still combined by the same tree over the lane indices (xor 16, 8, 4, 2, 1; addition commutes), so a
// key is the same bits whatever TR is and wherever the row sits in a tile.
// On return the total of row r is in v[0] of the lanes with (lane >> (5 - log2 TR)) == r.
__device__ __forceinline__ double shfl_xor_t(double v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ float shfl_xor_t(float v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ long long shfl_xor_t(long long v, int m) { return __shfl_xor_sync(FULL, v, m); }
template <int TR, typename T>
__device__ __forceinline__ void reduce_rows(T (&v)[TR], int lane) {
    constexpr int L = TR == 32 ? 5 : TR == 16 ? 4 : TR == 8 ? 3 : TR == 4 ? 2 : TR == 2 ? 1 : 0;
#pragma unroll
    for (int s = 0; s < L; s++) {
        const int m = 16 >> s;
        const int cnt = TR >> (s + 1);
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < cnt; i++) {
            const T send = up ? v[i] : v[i + cnt];
            const T keep = up ? v[i + cnt] : v[i];
            v[i] = keep + shfl_xor_t(send, m);
        }
    }
#pragma unroll
    for (int m = 16 >> L; m >= 1; m >>= 1) v[0] += shfl_xor_t(v[0], m);
}
// The same halving over groups of LPR lanes that hold LPR partial sums each (packed short rows: lane = group * LPR + j):
// afterwards v[0] of lane (g, j) is the total of the j-th of the group's LPR values.
template <int LPR, typename T>
__device__ __forceinline__ void reduce_packed(T (&v)[LPR], int lane) {
#pragma unroll
    for (int st = 0, m = LPR / 2; m >= 1; m >>= 1, st++) {
        const int cnt = LPR >> (st + 1);
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < cnt; i++) {
            const T send = up ? v[i] : v[i + cnt];
            const T keep = up ? v[i + cnt] : v[i];
            v[i] = keep + shfl_xor_t(send, m);
        }
    }
}
template <int TR>
struct RowLane {
    static constexpr int T = TR == 32 ? 5 : TR == 16 ? 4 : TR == 8 ? 3 : TR == 4 ? 2 : TR == 2 ? 1 : 0;
    static constexpr int SH = 5 - T;
    __device__ static __forceinline__ int row(int lane) { return lane >> SH; }
    __device__ static __forceinline__ bool owner(int lane) { return (lane & ((1 << SH) - 1)) == 0; }
};

}  // namespace svdb
