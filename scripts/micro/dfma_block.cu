// dfma_block.cu -- what does a register-blocked scalar-DFMA inner loop reach when its operands come from
// shared memory in the pattern a K2 rewrite would use?  (DMMA.8x8x4 peaks at the same 37 TFLOP/s as DFMA on
// B200 and K2 gets 0.72 of it; this asks whether plain DFMA with 8 rows x 4 queries per thread would do better.)
//
// CTA = 256 threads = 16 row groups x 16 query groups over a [128 rows][KC] x [64 queries][KC] tile pair per
// "stage"; thread (rg, qg) owns rows rg*8..+7 and queries qg, qg+16, qg+32, qg+48.  Per pair of coordinates:
// 8 + 4 LDS.128 and 64 DFMA.  Operand tiles are XOR-swizzled per 16-byte chunk like the real kernel.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWS = 128, NQ = 64, KC = 32, STAGES = 3;
constexpr int X_BYTES = ROWS * KC * 8, Q_BYTES = NQ * KC * 8, STAGE_BYTES = X_BYTES + Q_BYTES;

__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int RT, int QT>   // rows and queries per thread (RT*QT accumulators)
__global__ void __launch_bounds__(256, 1) dfma_kernel(double *out, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    double *f = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 8; i += blockDim.x) f[i] = 1.0 + 1e-6 * (i & 1023);
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr int RG = ROWS / RT, QG = NQ / QT;      // 16 x 16 for 8 x 4
    static_assert(RG * QG == 256, "one thread per (row group, query group)");
    const int rg = threadIdx.x / QG, qg = threadIdx.x % QG;
    double acc[RT][QT];
#pragma unroll
    for (int i = 0; i < RT; i++)
#pragma unroll
        for (int j = 0; j < QT; j++) acc[i][j] = 0.0;
    int stage = 0;
    for (int it = 0; it < iters; it++) {
        const uint32_t xs = sbase + stage * STAGE_BYTES, qs = xs + X_BYTES;
#pragma unroll
        for (int c = 0; c < KC / 2; c++) {           // 16-byte chunk = two coordinates
            double2 a[RT], b[QT];
#pragma unroll
            for (int i = 0; i < RT; i++) {
                const int r = rg * RT + i;
                a[i] = lds128(xs + r * (KC * 8) + ((c ^ ((r ^ (r >> 3)) & 7)) << 4));
            }
#pragma unroll
            for (int j = 0; j < QT; j++) {
                const int q = qg + j * QG;
                b[j] = lds128(qs + q * (KC * 8) + ((c ^ (q & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < RT; i++)
#pragma unroll
                for (int j = 0; j < QT; j++) {
                    acc[i][j] = fma(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fma(a[i].y, b[j].y, acc[i][j]);
                }
        }
        if (++stage == STAGES) stage = 0;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < RT; i++)
#pragma unroll
        for (int j = 0; j < QT; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int RT, int QT>
void run(const char *name, double *out, int sms) {
    const int iters = 3000;
    const size_t smem = (size_t)STAGES * STAGE_BYTES;
    cudaFuncSetAttribute(dfma_kernel<RT, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dfma_kernel<RT, QT><<<sms, 256, smem>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    dfma_kernel<RT, QT><<<sms, 256, smem>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * ROWS * NQ * KC * (double)iters * sms;
    printf("%-32s %8.3f ms %7.2f TFLOP/s  (%s)\n", name, ms, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 256);
    run<8, 4>("DFMA 8 rows x 4 queries / thread", out, sms);
    run<4, 8>("DFMA 4 rows x 8 queries / thread", out, sms);
    run<8, 4>("DFMA 8 x 4 again", out, sms);
    return 0;
}
