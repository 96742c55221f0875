// CPU emulation build of K10's selection source (csrc/umma_select.cuh): one warp prunes random candidate buffers
// (uf_prune), then sorts the survivors (uf_sort32) exactly like umma_filter_kernel's emit pass; the result is checked
// against std::sort.  Test infrastructure (tests/test_umma_select_sim.py).
//   umma_select_sim <seed> <nbuf> <cap> <mode>     mode 0: random keys, 1: few distinct keys (ties), 2: negatives and -FLT_MAX
#include <float.h>
#include <stdio.h>

#include <random>

#include "cusim_common.h"
#include "umma_select.cuh"

using namespace svdb;

struct Out {
    unsigned kept;
    int dropped;
    float tau;
    unsigned long long sorted[32];
    UfEntry front[32];
};

__global__ void select_kernel(UfEntry *bufs, const unsigned *cnts, int nbuf, int cap, Out *out) {
    const int lane = threadIdx.x & 31;
    for (int b = 0; b < nbuf; b++) {
        UfEntry *buf = bufs + (size_t)b * UF_BUF;
        float tau = 0.f;
        bool dropped = false;
        const unsigned kept = uf_prune(buf, cnts[b], cap, lane, tau, dropped);
        __syncwarp(FULL);
        u64 v = ~0ull;
        if ((unsigned)lane < kept) {
            const UfEntry e = buf[lane];
            v = ((u64)uf_ord(e.key) << 32) | e.row;
            out[b].front[lane] = e;
        }
        v = uf_sort32(v, lane);
        out[b].sorted[lane] = v;
        if (lane == 0) {
            out[b].kept = kept;
            out[b].dropped = dropped;
            out[b].tau = tau;
        }
    }
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const unsigned seed = (unsigned)atoi(argv[1]);
    const int nbuf = atoi(argv[2]), cap = atoi(argv[3]), mode = atoi(argv[4]);
    std::mt19937 rng(seed);
    std::vector<UfEntry> bufs((size_t)nbuf * UF_BUF), orig;
    std::vector<unsigned> cnts(nbuf);
    for (int b = 0; b < nbuf; b++) {
        cnts[b] = b == 0 ? 0 : (b == 1 ? UF_BUF : (b == 2 ? (unsigned)cap : (b == 3 ? (unsigned)cap + 1 : rng() % (UF_BUF + 1))));
        for (int i = 0; i < UF_BUF; i++) {
            float key;
            if (mode == 1) key = (float)(rng() % 5);
            else if (mode == 2) key = (rng() % 7 == 0) ? -FLT_MAX : ((float)(rng() % 2000) - 1000.f) * 0.37f;
            else key = std::uniform_real_distribution<float>(0.f, 200.f)(rng);
            bufs[(size_t)b * UF_BUF + i] = UfEntry{key, (uint32_t)rng()};
        }
    }
    orig = bufs;
    std::vector<Out> out(nbuf);
    cusim::launch(1, 32, select_kernel, bufs.data(), (const unsigned *)cnts.data(), nbuf, cap, out.data());
    // thr must never exclude a key below tau
    for (float tau : {1.f, 133.7f, -5.f, 1e-20f, 3e30f})
        for (float qn : {0.f, 256.f, 1e-10f, 1e25f}) {
            const float thr = uf_thr(tau, qn);
            if (!((double)thr >= (double)tau - (double)qn)) { printf("FAIL thr %g %g\n", tau, qn); return 1; }
        }
    for (int b = 0; b < nbuf; b++) {
        const unsigned cnt = cnts[b];
        std::vector<float> keys;
        for (unsigned i = 0; i < cnt; i++) keys.push_back(orig[(size_t)b * UF_BUF + i].key);
        std::sort(keys.begin(), keys.end());
        const unsigned want_kept = std::min<unsigned>(cnt, (unsigned)cap);
        if (out[b].kept != want_kept) { printf("FAIL buf %d kept %u want %u\n", b, out[b].kept, want_kept); return 1; }
        if ((out[b].dropped != 0) != (cnt > (unsigned)cap)) { printf("FAIL buf %d dropped flag\n", b); return 1; }
        if (out[b].dropped && out[b].tau != keys[cap - 1]) { printf("FAIL buf %d tau %g want %g\n", b, out[b].tau, keys[cap - 1]); return 1; }
        // the survivors are exactly the `kept` smallest keys, ascending, each an entry of the original buffer
        for (unsigned i = 0; i < 32; i++) {
            const unsigned long long v = out[b].sorted[i];
            if (i >= want_kept) {
                if (v != ~0ull) { printf("FAIL buf %d slot %u not empty\n", b, i); return 1; }
                continue;
            }
            const float key = uf_unord((uint32_t)(v >> 32));
            if (key != keys[i]) { printf("FAIL buf %d slot %u key %g want %g\n", b, i, key, keys[i]); return 1; }
            if (i > 0 && out[b].sorted[i - 1] > v) { printf("FAIL buf %d order\n", b); return 1; }
            bool found = false;
            for (unsigned j = 0; j < cnt && !found; j++)
                found = orig[(size_t)b * UF_BUF + j].key == key && orig[(size_t)b * UF_BUF + j].row == (uint32_t)v;
            if (!found) { printf("FAIL buf %d slot %u is not an entry of the buffer\n", b, i); return 1; }
        }
        // no entry kept twice
        for (unsigned i = 1; i < want_kept; i++)
            if (out[b].sorted[i] == out[b].sorted[i - 1]) {
                unsigned dup = 0;
                for (unsigned j = 0; j < cnt; j++)
                    dup += (((u64)uf_ord(orig[(size_t)b * UF_BUF + j].key) << 32) | orig[(size_t)b * UF_BUF + j].row) == out[b].sorted[i];
                if (dup < 2) { printf("FAIL buf %d duplicate survivor\n", b); return 1; }
            }
    }
    printf("OK %d buffers\n", nbuf);
    return 0;
}
