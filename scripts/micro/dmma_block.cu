// dmma_block.cu -- does the DMMA rate depend on the register blocking (operand reuse pattern)?
// Each warp keeps an MI x NI block of m8n8k4 accumulators and cycles through MI a-fragments and NI
// b-fragments held in registers (no memory traffic at all).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MI, int NI, bool NI_OUTER>
__global__ void __launch_bounds__(256, 1) blk_kernel(double *out, int iters) {
    double c[MI][NI][2], a[MI], b[NI];
    for (int i = 0; i < MI; i++) { a[i] = threadIdx.x * 1e-3 + i; for (int j = 0; j < NI; j++) c[i][j][0] = c[i][j][1] = 0.0; }
    for (int j = 0; j < NI; j++) b[j] = 1.0 + threadIdx.x * 1e-6 + j;
    for (int it = 0; it < iters; it++) {
        if (NI_OUTER) {
#pragma unroll
            for (int j = 0; j < NI; j++)
#pragma unroll
                for (int i = 0; i < MI; i++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        } else {
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        // perturb the fragments a little so that nothing can be hoisted (cheap: MI + NI DADDs per MI*NI DMMAs)
#pragma unroll
        for (int i = 0; i < MI; i++) a[i] += 1e-9;
#pragma unroll
        for (int j = 0; j < NI; j++) b[j] += 1e-9;
    }
    double s = 0;
    for (int i = 0; i < MI; i++) for (int j = 0; j < NI; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MI, int NI, bool NO>
void run(const char *name, double *out, int sms) {
    const int iters = 4000, grid = sms, threads = 256;
    blk_kernel<MI, NI, NO><<<grid, threads>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    blk_kernel<MI, NI, NO><<<grid, threads>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * 256 * MI * NI * (double)iters * (threads / 32) * grid;
    printf("%-28s 8 warps/SM: %8.3f ms %7.2f TFLOP/s\n", name, ms, fl / ms / 1e9);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 256);
    run<4, 4, false>("4x4 mi-outer", out, sms);
    run<4, 4, true>("4x4 ni-outer", out, sms);
    run<2, 8, false>("2x8 mi-outer", out, sms);
    run<8, 2, false>("8x2 mi-outer", out, sms);
    run<2, 4, false>("2x4 mi-outer", out, sms);
    run<4, 2, false>("4x2 mi-outer", out, sms);
    run<1, 8, false>("1x8", out, sms);
    run<8, 1, false>("8x1", out, sms);
    run<2, 2, false>("2x2", out, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
