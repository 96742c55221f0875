// stands in for csrc/common.cuh in the CPU emulation build: only what median_tree.cu uses from it
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
namespace svdb {
typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;
constexpr u64 SEQ_NONE = ~0ull;
}
#include "warplist.cuh"
