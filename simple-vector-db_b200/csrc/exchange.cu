// exchange.cu -- the cross-shard candidate exchange over NVLink peer memory, fused with K7.
//
// One process per GPU.  Every rank owns a small gather buffer in its own HBM and maps all the
// peers' buffers (CUDA IPC).  After finalize, a rank STORES its k x 32-byte answers straight
// into every peer's buffer (NVLink/NVSwitch writes), fences, and publishes an epoch flag there;
// the merge kernel on each rank spins on its LOCAL flags until every shard of this epoch has
// landed and merges (K7) in the same launch.  No NCCL call, no host round trip: the exchange is
// two tiny launches behind the scan on the same stream.  (The payload is 32..8192 bytes per
// rank, so this is purely about latency.)
//
// Buffer layout (per rank, cudaMalloc'ed so that it can be exported):
//   [ flags: 32 x u64 ]  flags[src] = last epoch whose data from `src` is complete;
//                        flags[31]  = this rank's own epoch counter, advanced ON THE DEVICE by the push
//                        kernel, so that the exchange can sit inside a replayed CUDA graph
//   [ parity 0: world x max_rec candidates ][ parity 1: ... ]   double-buffered by epoch parity
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"
#include "engine.h"
#include "kernels.h"
#include "tail.cuh"

namespace svdb {

// local results (nrec candidates) -> slot `rank` of every peer's gather buffer, then the flags (tail.cuh: xch_push)
__global__ void __launch_bounds__(256) exchange_push_kernel(const svdb_candidate *__restrict__ local, int nrec, PeerPtrs peers,
                                                            int rank, int world, size_t max_rec) {
    __shared__ u64 s_epoch;
    xch_push(local, nrec, peers, rank, world, max_rec, &s_epoch);
}

// wait until every shard's data of `epoch` has landed locally, then K7 (one warp per query)
__global__ void __launch_bounds__(32) exchange_merge_kernel(const unsigned char *mine, int world, size_t max_rec,
                                                            int nq, int k, svdb_candidate *out) {
    const int qi = blockIdx.x, lane = threadIdx.x;
    const u64 epoch = __ldcg(reinterpret_cast<const u64 *>(mine) + XCH_EPOCH_SLOT);
    xch_wait(mine, world, epoch, lane);
    const svdb_candidate *in = reinterpret_cast<const svdb_candidate *>(
        mine + XCH_FLAG_BYTES + (size_t)(epoch & 1) * world * max_rec * sizeof(svdb_candidate));
    merge_gathered(in, world, max_rec, qi, k, lane, out);
}

// generic all-gather on the same buffers: wait for every rank's `nrec` 32-byte records of this epoch and copy
// them out in rank order (out: [world][nrec] records)
__global__ void __launch_bounds__(256) exchange_wait_copy_kernel(const unsigned char *mine, int world, size_t max_rec, int nrec,
                                                                 u64 *__restrict__ out) {
    const u64 *flags = reinterpret_cast<const u64 *>(mine);
    const u64 epoch = __ldcg(flags + XCH_EPOCH_SLOT);
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(flags + threadIdx.x) < epoch) {
            if (clock64() - t0 > 20000000000ll) __trap();
        }
    }
    __syncthreads();
    const u64 *in = reinterpret_cast<const u64 *>(mine + XCH_FLAG_BYTES + (size_t)(epoch & 1) * world * max_rec * sizeof(svdb_candidate));
    const int words = nrec * 4;
    for (int r = 0; r < world; r++)
        for (int i = threadIdx.x; i < words; i += blockDim.x) out[(size_t)r * words + i] = __ldcg(in + (size_t)r * max_rec * 4 + i);
}

}  // namespace svdb

using namespace svdb;

struct svdb_exchange {
    int device = 0, rank = 0, world = 1;
    size_t max_rec = 0, bytes = 0;
    unsigned char *mine = nullptr;
    PeerPtrs peers{};
    bool opened[XCH_MAX_WORLD] = {};
    // staging of the host-level all-gather (tie resolution): one chunk of max_rec records per rank
    unsigned char *d_send = nullptr, *d_recv = nullptr, *h_send = nullptr, *h_recv = nullptr;
};

namespace svdb {
// enqueue "store to every peer + wait for every peer + merge" on st (collective; see svdb_exchange_merge)
cudaError_t exchange_enqueue(svdb_exchange *x, cudaStream_t st, const svdb_candidate *d_local, size_t nq, size_t k,
                             svdb_candidate *out) {
    exchange_push_kernel<<<1, 256, 0, st>>>(d_local, (int)(nq * k), x->peers, x->rank, x->world, x->max_rec);
    exchange_merge_kernel<<<(unsigned)nq, 32, 0, st>>>(x->mine, x->world, x->max_rec, (int)nq, (int)k, out);
    return cudaGetLastError();
}
bool exchange_fits(const svdb_exchange *x, size_t nq, size_t k) { return x && nq * k <= x->max_rec; }
int exchange_rank(const svdb_exchange *x) { return x->rank; }
// what a scan's fused tail needs to exchange and merge by itself (kernels.h: TailArgs)
void exchange_fill_tail(const svdb_exchange *x, TailArgs &t, svdb_candidate *xout) {
    t.world = x->world;
    t.rank = x->rank;
    t.peers = x->peers;
    t.max_rec = x->max_rec;
    t.xout = xout;
}
int exchange_world(const svdb_exchange *x) { return x->world; }

// Host buffers in, host buffers out: recv = world blocks of `bytes`, in rank order.  Collective; synchronizes st.
int exchange_allgather_host(svdb_exchange *x, cudaStream_t st, const void *send, void *recv, size_t bytes) {
    if (!x || !send || !recv) return SVDB_ERR_ARG;
    if (bytes == 0) return SVDB_OK;
    cudaError_t ce = cudaSetDevice(x->device);
    const size_t chunk = x->max_rec * sizeof(svdb_candidate);
    if (ce == cudaSuccess && !x->d_send) {
        ce = cudaMalloc(&x->d_send, chunk);
        if (ce == cudaSuccess) ce = cudaMalloc(&x->d_recv, chunk * x->world);
        if (ce == cudaSuccess) ce = cudaMallocHost(&x->h_send, chunk);
        if (ce == cudaSuccess) ce = cudaMallocHost(&x->h_recv, chunk * x->world);
    }
    for (size_t off = 0; off < bytes && ce == cudaSuccess; off += chunk) {
        const size_t len = std::min(chunk, bytes - off);
        const int nrec = (int)((len + sizeof(svdb_candidate) - 1) / sizeof(svdb_candidate));
        memset(x->h_send, 0, (size_t)nrec * sizeof(svdb_candidate));
        memcpy(x->h_send, static_cast<const unsigned char *>(send) + off, len);
        ce = cudaMemcpyAsync(x->d_send, x->h_send, (size_t)nrec * sizeof(svdb_candidate), cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) break;
        exchange_push_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const svdb_candidate *>(x->d_send), nrec, x->peers, x->rank,
                                                x->world, x->max_rec);
        exchange_wait_copy_kernel<<<1, 256, 0, st>>>(x->mine, x->world, x->max_rec, nrec, reinterpret_cast<u64 *>(x->d_recv));
        ce = cudaGetLastError();
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(x->h_recv, x->d_recv, (size_t)x->world * nrec * sizeof(svdb_candidate), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) break;
        for (int r = 0; r < x->world; r++)
            memcpy(static_cast<unsigned char *>(recv) + (size_t)r * bytes + off,
                   x->h_recv + (size_t)r * nrec * sizeof(svdb_candidate), len);
    }
    if (ce != cudaSuccess) {
        set_last_error(std::string("exchange all-gather: ") + cudaGetErrorString(ce));
        cudaGetLastError();
        return SVDB_ERR_CUDA;
    }
    return SVDB_OK;
}
}  // namespace svdb

extern "C" {

int svdb_exchange_create(int device, int rank, int world, size_t max_records, svdb_exchange **out, unsigned char handle_out[64]) {
    if (!out || !handle_out || world < 1 || world > XCH_MAX_WORLD || rank < 0 || rank >= world || max_records == 0) {
        set_last_error("bad argument to svdb_exchange_create");
        return SVDB_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    *out = nullptr;
    svdb_exchange *x = new (std::nothrow) svdb_exchange();
    if (!x) return SVDB_ERR_OOM;
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->max_rec = max_records;
    x->bytes = XCH_FLAG_BYTES + 2 * (size_t)world * max_records * sizeof(svdb_candidate);
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaMalloc(&x->mine, x->bytes);
    if (ce == cudaSuccess) ce = cudaMemset(x->mine, 0, x->bytes);
    cudaIpcMemHandle_t h;
    if (ce == cudaSuccess) ce = cudaIpcGetMemHandle(&h, x->mine);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        set_last_error(std::string("svdb_exchange_create: ") + cudaGetErrorString(ce));
        cudaGetLastError();
        if (x->mine) cudaFree(x->mine);
        delete x;
        return SVDB_ERR_CUDA;
    }
    memcpy(handle_out, &h, 64);
    x->peers.p[rank] = x->mine;
    *out = x;
    return SVDB_OK;
}

int svdb_exchange_connect(svdb_exchange *x, const unsigned char *all_handles) {
    if (!x || !all_handles) return SVDB_ERR_ARG;
    cudaSetDevice(x->device);
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) {
            set_last_error(std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(ce));
            cudaGetLastError();
            return SVDB_ERR_CUDA;
        }
        x->peers.p[r] = static_cast<unsigned char *>(p);
        x->opened[r] = true;
    }
    return SVDB_OK;
}

void svdb_exchange_destroy(svdb_exchange *x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; r++)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->peers.p[r]);
    if (x->mine) cudaFree(x->mine);
    if (x->d_send) cudaFree(x->d_send);
    if (x->d_recv) cudaFree(x->d_recv);
    if (x->h_send) cudaFreeHost(x->h_send);
    if (x->h_recv) cudaFreeHost(x->h_recv);
    delete x;
}

// d_local: this shard's nq x k candidates (e.g. from svdb_nearest_batch_device) -> d_out: merged nq x k.
// Collective: every rank must call it the same number of times. Asynchronous on `stream`.
int svdb_exchange_merge(svdb_exchange *x, void *stream, const svdb_candidate *d_local, size_t nq, size_t k, svdb_candidate *d_out) {
    if (!x || !d_local || !d_out || k < 1 || k > SVDB_MAX_K || nq * k > x->max_rec) {
        set_last_error("bad argument to svdb_exchange_merge (nq*k exceeds the exchange capacity?)");
        return SVDB_ERR_ARG;
    }
    if (nq == 0) return SVDB_OK;
    cudaSetDevice(x->device);
    cudaError_t ce = exchange_enqueue(x, (cudaStream_t)stream, d_local, nq, k, d_out);
    if (ce != cudaSuccess) {
        set_last_error(std::string("svdb_exchange_merge: ") + cudaGetErrorString(ce));
        return SVDB_ERR_CUDA;
    }
    return SVDB_OK;
}

}  // extern "C"
