// cuda_runtime.h -- NOT the CUDA header.  A small CPU emulation of the CUDA execution model used by
// tests/test_mtree_sim.py to run the K8/K9 kernel SOURCE (median_tree.cu, launch syntax rewritten by the
// test) on the host: every CTA runs as blockDim.x std::threads, __syncthreads is a barrier, the warp
// collectives (__shfl_up_sync, __reduce_*_sync, __ballot_sync) rendezvous the lanes named by their mask.
// Test infrastructure only; nothing in the product includes it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define SVDB_CUSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static          /* CTAs run one after the other, so one static copy per kernel is "per CTA" */
#define __align__(x)

typedef int cudaError_t;
typedef void *cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorInvalidValue = 1;
inline cudaError_t cudaMalloc(void **p, size_t bytes) { *p = malloc(bytes); memset(*p, 0xA5, bytes); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }

namespace cusim {
struct Dim { unsigned x = 1, y = 1, z = 1; };
struct Barrier {                       // __syncthreads with threads that may have left the kernel
    std::mutex m;
    std::condition_variable cv;
    int expected = 0, waiting = 0;
    unsigned gen = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        if (++waiting == expected) { waiting = 0; gen++; cv.notify_all(); return; }
        const unsigned g = gen;
        cv.wait(lk, [&] { return gen != g; });
    }
    void drop() {
        std::unique_lock<std::mutex> lk(m);
        expected--;
        if (expected > 0 && waiting == expected) { waiting = 0; gen++; cv.notify_all(); }
    }
};
struct Warp {
    std::mutex m;
    std::condition_variable cv;
    struct Group { int arrived = 0; unsigned gen = 0; uint64_t slot[32], snap[32]; };
    std::map<unsigned, Group> groups;
    // every lane named by mask calls this with its value; returns all 32 values as they were at the rendezvous
    void gather(unsigned mask, int lane, uint64_t v, uint64_t out[32]) {
        std::unique_lock<std::mutex> lk(m);
        Group &g = groups[mask];
        g.slot[lane] = v;
        if (++g.arrived == __builtin_popcount(mask)) {
            memcpy(g.snap, g.slot, sizeof g.snap);
            g.arrived = 0;
            g.gen++;
            cv.notify_all();
        } else {
            const unsigned gen = g.gen;
            cv.wait(lk, [&] { return g.gen != gen; });
        }
        memcpy(out, g.snap, sizeof g.snap);
    }
};
struct Tls { Dim tid, bid, bdim, gdim; Barrier *bar = nullptr; Warp *warp = nullptr; int lane = 0; };
inline thread_local Tls tls;

template <typename... P, typename... A>
void launch(unsigned grid, unsigned block, void (*kernel)(P...), A... args) {
    for (unsigned b = 0; b < grid; b++) {
        Barrier bar;
        bar.expected = (int)block;
        std::vector<Warp> warps((block + 31) / 32);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t, b] {
                tls.tid.x = t; tls.bid.x = b; tls.bdim.x = block; tls.gdim.x = grid;
                tls.bar = &bar; tls.warp = &warps[t / 32]; tls.lane = (int)(t % 32);
                kernel(static_cast<P>(args)...);
                bar.drop();
            });
        for (auto &x : th) x.join();
    }
}
}  // namespace cusim

#define threadIdx (cusim::tls.tid)
#define blockIdx (cusim::tls.bid)
#define blockDim (cusim::tls.bdim)
#define gridDim (cusim::tls.gdim)

inline void __syncthreads() { cusim::tls.bar->wait(); }
inline void __syncwarp(unsigned mask) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, 0, s);
}
template <typename T> inline uint64_t cusim_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof v); return b; }
template <typename T> inline T cusim_from(uint64_t b) { T v; memcpy(&v, &b, sizeof v); return v; }
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, int o) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, cusim_bits(v), s);
    return cusim::tls.lane >= o ? cusim_from<T>(s[cusim::tls.lane - o]) : v;
}
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, cusim_bits(v), s);
    return cusim_from<T>(s[src & 31]);
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int x) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, cusim_bits(v), s);
    return cusim_from<T>(s[(cusim::tls.lane ^ x) & 31]);
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, v, s);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) r += (unsigned)s[i];
    return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, v, s);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) r = std::min(r, (unsigned)s[i]);
    return r;
}
inline unsigned __ballot_sync(unsigned mask, bool p) {
    uint64_t s[32];
    cusim::tls.warp->gather(mask, cusim::tls.lane, p ? 1 : 0, s);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i & 1) && s[i]) r |= 1u << i;
    return r;
}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline long long __double_as_longlong(double x) { long long v; memcpy(&v, &x, 8); return v; }
inline double __longlong_as_double(long long v) { double x; memcpy(&x, &v, 8); return x; }
inline int __double2hiint(double x) { return (int)((unsigned long long)__double_as_longlong(x) >> 32); }
inline int __double2loint(double x) { return (int)((unsigned long long)__double_as_longlong(x) & 0xffffffffull); }
inline double __hiloint2double(int hi, int lo) {
    return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo));
}
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
using std::max;
using std::min;
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
