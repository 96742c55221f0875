// arena.h -- growable device buffers for the HBM-resident store.
//
// A DeviceBuffer reserves a large virtual address range once and maps physical HBM into
// it as the store grows (CUDA virtual memory management: cuMemAddressReserve / cuMemCreate /
// cuMemMap).  Growing never moves data, so a 60+ GB shard can keep appending without ever
// needing twice its size, and device pointers handed to kernels stay valid.
// The driver entry points are resolved at run time through cudaGetDriverEntryPoint, so
// the library has no link-time dependency on libcuda and still loads on a box without a
// driver (where every call then fails loudly).
// SVDB_ARENA=malloc switches to plain cudaMalloc + copy-on-grow (diagnostic aid only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace svdb {

class DeviceBuffer {
public:
    DeviceBuffer() = default;
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;

    // Reserve address space for up to max_bytes on `device`. No physical memory yet.
    bool init(int device, size_t max_bytes, std::string &err);
    // Make at least `bytes` usable (mapped + read/write). Existing contents are preserved.
    // `st` is only used by the malloc fallback mode (copy of the old contents).
    bool ensure(size_t bytes, cudaStream_t st, std::string &err);
    void release();

    void *ptr() const { return reinterpret_cast<void *>(base_); }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(base_); }
    size_t mapped() const { return mapped_; }
    size_t reserved() const { return reserved_; }

private:
    int device_ = 0;
    bool vmm_ = true;
    uintptr_t base_ = 0;
    size_t reserved_ = 0, mapped_ = 0, gran_ = 0;
    std::vector<unsigned long long> handles_;
    std::vector<size_t> handle_sizes_;
};

}  // namespace svdb
