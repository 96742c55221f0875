// fp64_peak.cu -- what the FP64 pipes of this GPU sustain: scalar DFMA vs mma.sync f64 (DMMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; run: ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma884_kernel(double *out, int iters) {
    double c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: A 16x16 (8 regs/lane), B 16x8 (4 regs), C 16x8 (4 regs)
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void dmma16816_kernel(double *out, int iters) {
    double c[4][4], a[8], b[4];
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c[j][k] = 0.0;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-3 + k;
    for (int k = 0; k < 4; k++) b[k] = 1.0 + threadIdx.x * 1e-6 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma16816(c[j], a, b);
    }
    double s = 0;
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) s += c[j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        const int threads = warps * 32, grid = sms * 2;
        float ms = time_ms([&] { dfma_kernel<<<grid, threads>>>(out, iters); });
        double fl = 2.0 * 8 * iters * (double)threads * grid;
        printf("DFMA      warps/CTA=%2d (2 CTA/SM): %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
        ms = time_ms([&] { dmma884_kernel<<<grid, threads>>>(out, iters); });
        fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * grid;
        printf("DMMA 884  warps/CTA=%2d (2 CTA/SM): %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
        ms = time_ms([&] { dmma16816_kernel<<<grid, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)warps * grid;
        printf("DMMA16816 warps/CTA=%2d (2 CTA/SM): %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms / 1e9);
    }
    // sustained: ~1 s of back-to-back DMMA / DFMA (the power cap, not the pipe, may set the ceiling)
    {
        const int threads = 256, grid = sms * 2, reps = 60;
        for (int which = 0; which < 2; which++) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaDeviceSynchronize();
            cudaEventRecord(a);
            for (int r = 0; r < reps; r++) {
                if (which == 0) dmma884_kernel<<<grid, threads>>>(out, iters * 4);
                else dfma_kernel<<<grid, threads>>>(out, iters * 32);
            }
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double fl = which == 0 ? 2.0 * 8 * 8 * 4 * 8 * (iters * 4.0) * 8 * grid * reps
                                         : 2.0 * 8 * (iters * 32.0) * threads * grid * reps;
            printf("%s sustained over %.0f ms: %7.2f TFLOP/s\n", which == 0 ? "DMMA 884" : "DFMA    ", ms, fl / ms / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
