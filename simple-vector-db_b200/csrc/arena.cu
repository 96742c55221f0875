// arena.cu -- see arena.h
#include "arena.h"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace svdb {

namespace {

struct DriverApi {
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    bool ok = false;
    std::string why;
};

template <typename F>
bool resolve(const char *name, F &fn, std::string &why) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        why = std::string("cannot resolve driver entry point ") + name + ": " + cudaGetErrorString(e);
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

DriverApi &driver() {
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        DriverApi &a = api;
        a.ok = resolve("cuMemGetAllocationGranularity", a.MemGetAllocationGranularity, a.why) &&
               resolve("cuMemAddressReserve", a.MemAddressReserve, a.why) &&
               resolve("cuMemAddressFree", a.MemAddressFree, a.why) && resolve("cuMemCreate", a.MemCreate, a.why) &&
               resolve("cuMemRelease", a.MemRelease, a.why) && resolve("cuMemMap", a.MemMap, a.why) &&
               resolve("cuMemUnmap", a.MemUnmap, a.why) && resolve("cuMemSetAccess", a.MemSetAccess, a.why) &&
               resolve("cuGetErrorString", a.GetErrorString, a.why);
    });
    return api;
}

std::string cu_err(const char *what, CUresult r) {
    const char *s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return std::string(what) + " failed: " + (s ? s : "unknown CUresult") + " (" + std::to_string((int)r) + ")";
}

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

}  // namespace

bool DeviceBuffer::init(int device, size_t max_bytes, std::string &err) {
    release();
    device_ = device;
    const char *mode = getenv("SVDB_ARENA");
    vmm_ = !(mode && strcmp(mode, "malloc") == 0);
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {
        err = std::string("cudaSetDevice failed: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    if (!vmm_) {
        reserved_ = max_bytes;
        return true;
    }
    DriverApi &d = driver();
    if (!d.ok) {
        err = d.why;
        return false;
    }
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    CUresult r = d.MemGetAllocationGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran_ == 0) {
        err = cu_err("cuMemGetAllocationGranularity", r);
        return false;
    }
    reserved_ = round_up(max_bytes ? max_bytes : gran_, gran_);
    CUdeviceptr p = 0;
    r = d.MemAddressReserve(&p, reserved_, 0, 0, 0);
    if (r != CUDA_SUCCESS) {
        err = cu_err("cuMemAddressReserve", r);
        reserved_ = 0;
        return false;
    }
    base_ = (uintptr_t)p;
    return true;
}

bool DeviceBuffer::ensure(size_t bytes, cudaStream_t st, std::string &err) {
    if (bytes <= mapped_) return true;
    if (bytes > reserved_) {
        err = "store exceeds the reserved address range (" + std::to_string(reserved_) + " bytes)";
        return false;
    }
    cudaSetDevice(device_);
    if (!vmm_) {
        size_t target = mapped_ + mapped_ / 2;
        if (target < bytes) target = bytes;
        if (target > reserved_) target = reserved_;
        void *np = nullptr;
        if (cudaMalloc(&np, target) != cudaSuccess) {
            target = bytes;
            if (cudaMalloc(&np, target) != cudaSuccess) {
                err = std::string("cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError());
                return false;
            }
        }
        if (mapped_) {
            cudaMemcpyAsync(np, (void *)base_, mapped_, cudaMemcpyDeviceToDevice, st);
            cudaStreamSynchronize(st);
            cudaFree((void *)base_);
        }
        base_ = (uintptr_t)np;
        mapped_ = target;
        return true;
    }
    DriverApi &d = driver();
    // geometric growth, bounded: at most +50% or +4 GiB beyond what was asked
    size_t extra = mapped_ / 2;
    const size_t lo = 2 * gran_, hi = (size_t)4 << 30;
    if (extra < lo) extra = lo;
    if (extra > hi) extra = hi;
    size_t target = round_up(bytes > mapped_ + extra ? bytes : mapped_ + extra, gran_);
    if (target > reserved_) target = reserved_;
    const size_t exact = round_up(bytes, gran_);

    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device_;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;

    const size_t piece_max = round_up((size_t)8 << 30, gran_);
    while (mapped_ < target) {
        size_t piece = target - mapped_;
        if (piece > piece_max) piece = piece_max;
        CUmemGenericAllocationHandle h;
        CUresult r = d.MemCreate(&h, piece, &prop, 0);
        if (r != CUDA_SUCCESS) {
            if (mapped_ >= exact) break;     // the optional head-room did not fit: fine
            if (target > exact) {            // retry without head-room
                target = exact;
                continue;
            }
            err = cu_err("cuMemCreate", r) + " while growing to " + std::to_string(bytes) + " bytes";
            return false;
        }
        r = d.MemMap((CUdeviceptr)(base_ + mapped_), piece, 0, h, 0);
        if (r != CUDA_SUCCESS) {
            d.MemRelease(h);
            err = cu_err("cuMemMap", r);
            return false;
        }
        r = d.MemSetAccess((CUdeviceptr)(base_ + mapped_), piece, &acc, 1);
        if (r != CUDA_SUCCESS) {
            d.MemUnmap((CUdeviceptr)(base_ + mapped_), piece);
            d.MemRelease(h);
            err = cu_err("cuMemSetAccess", r);
            return false;
        }
        handles_.push_back((unsigned long long)h);
        handle_sizes_.push_back(piece);
        mapped_ += piece;
    }
    return mapped_ >= bytes;
}

void DeviceBuffer::release() {
    if (!base_ && !reserved_) return;
    cudaSetDevice(device_);
    if (!vmm_) {
        if (base_) cudaFree((void *)base_);
    } else if (driver().ok) {
        DriverApi &d = driver();
        size_t off = 0;
        for (size_t i = 0; i < handles_.size(); i++) {
            d.MemUnmap((CUdeviceptr)(base_ + off), handle_sizes_[i]);
            d.MemRelease((CUmemGenericAllocationHandle)handles_[i]);
            off += handle_sizes_[i];
        }
        if (base_) d.MemAddressFree((CUdeviceptr)base_, reserved_);
    }
    handles_.clear();
    handle_sizes_.clear();
    base_ = 0;
    reserved_ = mapped_ = 0;
}

}  // namespace svdb
