// smem_optin.h -- per-(kernel, device) opt-in to more than 48 KB of dynamic shared memory (host side).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

namespace svdb {

// Opt-in to more than 48 KB of dynamic shared memory is a property of (kernel, DEVICE): one instance per kernel
// instantiation remembers what each device ordinal has been configured for, so that engines on different GPUs of one
// process never skip it (and engines holding different mutexes never race on it).
struct SmemOptIn {
    static constexpr int MAX_DEV = 64;
    std::atomic<size_t> have[MAX_DEV];
    std::mutex mu;
    SmemOptIn() {
        for (auto &h : have) h.store(0);
    }
    template <typename Kernel>
    cudaError_t ensure(Kernel kernel, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= MAX_DEV) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (have[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
        std::lock_guard<std::mutex> g(mu);             // the attribute only ever grows
        if (have[dev].load(std::memory_order_relaxed) >= bytes) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) have[dev].store(bytes, std::memory_order_release);
        return e;
    }
};

}  // namespace svdb
