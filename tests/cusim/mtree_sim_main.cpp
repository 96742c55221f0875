// Runs the K8/K9 kernel source of csrc/median_tree.cu under the CPU emulation in cuda_runtime.h (this directory)
// and checks: the build is a permutation with valid median splits at every node; the traversal returns the
// smallest (reference-order distance, seq) of tree + tail and flags exactly the queries whose minimum is
// shared by entries with different coordinates.   usage: mtree_sim <n> <K> <tail> <nq> <dist> <seed> [levels per block: 3 | 1]
#include <stdio.h>

#include <random>
#include <utility>

#include "median_tree_sim.inc"

using svdb::u64;

static double sqdist(const double *p, const double *q, int K) {   // kdtree.c:134-137
    double d = 0.0;
    for (int i = 0; i < K; i++) {
        const double t = p[i] - q[i];
        d = d + t * t;
    }
    return d;
}

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    const u64 n = strtoull(argv[1], 0, 10);
    const int K = atoi(argv[2]);
    const u64 tail = strtoull(argv[3], 0, 10);
    const int nq = atoi(argv[4]);
    const int blk = argc > 7 ? atoi(argv[7]) : 3;
    const int dist = atoi(argv[5]);       // 0 uniform, 1 coarse grid (ties, duplicates), 2 sorted, 3 uniform with non-finite rows
    std::mt19937_64 rng(strtoull(argv[6], 0, 10));
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const u64 nt = n + tail;
    const int stride = K;
    std::vector<double> pts(std::max<u64>(nt, 1) * stride);
    auto gen = [&](u64 i, int c) -> double {
        if (dist == 1) return (double)(int)(rng() % 7) * 0.5 - 1.5;
        if (dist == 2) return (double)i + 0.25 * c;
        return U(rng);
    };
    for (u64 i = 0; i < nt; i++)
        for (int c = 0; c < K; c++) pts[i * stride + c] = gen(i, c);
    if (dist == 3)
        for (u64 i = 0; i < nt; i += 7) pts[i * stride + (i % K)] = (i % 3 == 0) ? NAN : ((i % 3 == 1) ? INFINITY : -INFINITY);
    std::vector<double> Q((size_t)nq * K);
    for (int i = 0; i < nq; i++)
        for (int c = 0; c < K; c++) Q[(size_t)i * K + c] = dist == 2 ? (double)(rng() % (nt ? nt : 1)) + 0.3
                                              : gen(0, c) + ((dist == 1 && (i & 1)) ? 0.25 : 0.0);   // between grid points: distinct ties
    std::vector<u64> log_index(std::max<u64>(nt, 1));
    for (u64 i = 0; i < nt; i++) log_index[i] = 1000 + i;

    const int L = svdb::mtree_levels(n);
    std::vector<double> split(svdb::mtree_split_count(n, blk), NAN), mpts(std::max<u64>(n, 1) * K, NAN);
    std::vector<uint32_t> mseq(std::max<u64>(n, 1), 0xdeadbeefu);
    int levels = -1, launches = 0;
    if (svdb::launch_mtree_build(pts.data(), stride, K, n, blk, split.data(), mpts.data(), mseq.data(), 4, nullptr, &levels, &launches)) return 3;
    if (levels != L) { printf("levels %d != %d\n", levels, L); return 1; }
    // permutation + payload
    std::vector<char> seen(n, 0);
    for (u64 p = 0; p < n; p++) {
        if (mseq[p] >= n || seen[mseq[p]]) { printf("mseq is not a permutation at %llu\n", (unsigned long long)p); return 1; }
        seen[mseq[p]] = 1;
        if (memcmp(&mpts[p * K], &pts[(u64)mseq[p] * stride], K * 8)) { printf("mpts mismatch at %llu\n", (unsigned long long)p); return 1; }
    }
    // every node: left <= split <= right on its axis (NaN coordinates sort with +inf)
    for (int l = 0; l < L; l++)
        for (u64 j = 0; j < (1ull << l); j++) {
            const u64 lo = svdb::mt_bound(j, n, l), hi = svdb::mt_bound(j + 1, n, l), mid = svdb::mt_bound(2 * j + 1, n, l + 1);
            const double s = split[svdb::mt_slot(l, j, blk)];
            if (!(lo < mid && mid < hi)) { printf("empty side at level %d seg %llu\n", l, (unsigned long long)j); return 1; }
            for (u64 p = lo; p < hi; p++) {
                double x = mpts[p * K + l % K];
                if (x != x) x = INFINITY;
                if (p < mid ? !(x <= s) : !(x >= s)) {
                    printf("split violated: level %d seg %llu pos %llu x %g s %g\n", l, (unsigned long long)j, (unsigned long long)p, x, s);
                    return 1;
                }
            }
        }
    for (u64 j = 0; j < (1ull << L); j++)
        if (svdb::mt_bound(j + 1, n, L) - svdb::mt_bound(j, n, L) > 32) { printf("leaf too large\n"); return 1; }

    svdb::MtreeView t;
    t.split = split.data(); t.mpts = mpts.data(); t.mseq = mseq.data(); t.n_built = n; t.levels = L; t.block_levels = blk;
    int flagged = 0;
    for (int lanes : {32, 16, 8}) {
        std::vector<svdb_candidate> out(nq);
        std::vector<unsigned> marks(nq, 77u);
        memset(out.data(), 0xEE, out.size() * sizeof(svdb_candidate));
        if (svdb::launch_mtree_nearest(t, pts.data(), stride, K, nt, Q.data(), K, nq, 1, log_index.data(), 5000, 1, lanes, marks.data(), out.data(), nullptr)) return 3;
        for (int i = 0; i < nq; i++) {
            const double *q = &Q[(size_t)i * K];
            double bd = INFINITY;
            u64 bs = ~0ull;
            for (u64 e = 0; e < nt; e++) {
                const double d = sqdist(&pts[e * stride], q, K);
                if (d < bd) { bd = d; bs = e; }
            }
            bool tie = false;
            if (bs != ~0ull)
                for (u64 e = 0; e < nt; e++)
                    if (sqdist(&pts[e * stride], q, K) == bd)
                        for (int c = 0; c < K; c++) tie |= pts[e * stride + c] != pts[bs * stride + c];
            const svdb_candidate &c = out[i];
            const bool ok = bs == ~0ull ? (c.seq == ~0ull && c.index == (u64)SVDB_NONE && c.dist == INFINITY && c.flags == 0)
                                        : (c.seq == bs + 5000 && c.index == 1000 + bs && memcmp(&c.dist, &bd, 8) == 0 &&
                                           c.flags == (tie ? SVDB_CAND_TIE : 0ull));
            if (!ok || marks[i] != (tie ? 1u : 0u)) {
                printf("lanes %d query %d: got seq %llu dist %.17g flags %llu, want seq %llu dist %.17g tie %d\n", lanes, i,
                       (unsigned long long)c.seq, c.dist, (unsigned long long)c.flags, (unsigned long long)(bs + 5000), bd, (int)tie);
                return 1;
            }
            flagged += tie && lanes == 32;
        }
    }
    // k > 1: the k smallest (distance, seq), flagged like k = 1
    for (int k : {2, 5, 24}) {
        std::vector<svdb_candidate> out((size_t)nq * k);
        std::vector<unsigned> marks(nq, 77u);
        memset(out.data(), 0xEE, out.size() * sizeof(svdb_candidate));
        if (svdb::launch_mtree_nearest(t, pts.data(), stride, K, nt, Q.data(), K, nq, k, log_index.data(), 5000, 1, 32, marks.data(), out.data(), nullptr)) return 3;
        for (int i = 0; i < nq; i++) {
            const double *q = &Q[(size_t)i * K];
            std::vector<std::pair<double, u64>> all;
            for (u64 e = 0; e < nt; e++) {
                const double d = sqdist(&pts[e * stride], q, K);
                if (d < INFINITY) all.push_back({d, e});
            }
            std::sort(all.begin(), all.end());
            bool tie = false;
            for (size_t a = 1; a < all.size() && all[a].first == all[0].first; a++)
                for (int c = 0; c < K; c++) tie |= pts[all[a].second * stride + c] != pts[all[0].second * stride + c];
            for (int r = 0; r < k; r++) {
                const svdb_candidate &c = out[(size_t)i * k + r];
                const bool ok = (size_t)r < all.size()
                                    ? (c.seq == all[r].second + 5000 && c.index == 1000 + all[r].second &&
                                       memcmp(&c.dist, &all[r].first, 8) == 0 && c.flags == (tie ? SVDB_CAND_TIE : 0ull))
                                    : (c.seq == ~0ull && c.index == (u64)SVDB_NONE && c.dist == INFINITY && c.flags == 0);
                if (!ok || marks[i] != (tie ? 1u : 0u)) {
                    printf("k %d query %d rank %d: got seq %llu dist %.17g flags %llu mark %u (tie %d)\n", k, i, r,
                           (unsigned long long)c.seq, c.dist, (unsigned long long)c.flags, marks[i], (int)tie);
                    return 1;
                }
            }
        }
    }
    printf("OK n=%llu K=%d tail=%llu levels=%d launches=%d flagged=%d/%d\n", (unsigned long long)n, K, (unsigned long long)tail, L,
           launches, flagged, nq);
    return 0;
}
