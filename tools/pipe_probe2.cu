// pipe_probe2.cu -- how many issue cycles does an FP64 instruction cost depending on where its
// operands come from (registers without reuse / uniform registers), with and without FP32 fillers?
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: a[k] = fma(a[k], b, c)        b, c kernel params (uniform registers / reuse)
// MODE 1: a[k] = fma(a[k], x[k], y[k])  all operands in distinct registers
// MODE 2: a[k] = a[k] + x[k]            DADD two register operands
// MODE 3: a[k] = a[k] + b               DADD one uniform operand
template <int MODE, int NF>
__global__ void __launch_bounds__(256) probe(double *sink, int iters, double b, double c, float fb) {
    double a[8], x[8], y[8];
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = 1.0 + 0.001 * k + 1e-6 * threadIdx.x;
        x[k] = 1.0 + 1e-9 * (k + threadIdx.x);
        y[k] = 1e-9 * (k + 1 + threadIdx.x);
        f[k] = 1.0f + 0.01f * k + threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) a[k] = fma(a[k], b, c);
                if (MODE == 1) a[k] = fma(a[k], x[k], y[k]);
                if (MODE == 2) a[k] = __dadd_rn(a[k], x[k]);
                if (MODE == 3) a[k] = __dadd_rn(a[k], b);
                if (k < NF) f[k] = fmaf(f[k], fb, 0.5f);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + f[k] + x[k] + y[k];
    if (s == 123.456) sink[0] = s;
}

template <int MODE, int NF>
void run(const char *name) {
    double *sink;
    cudaMalloc(&sink, 8);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE, NF><<<blocks, 256>>>(sink, 100, 1.0000000001, 1e-9, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE, NF><<<blocks, 256>>>(sink, iters, 1.0000000001, 1e-9, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // cycles per FP64 instruction per SMSP at 1.965 GHz: 64 warps/SM -> 16 warps per SMSP
    const double fp64_per_smsp = 16.0 * iters * 32.0;
    printf("%-44s %8.2f ms  %.2f cycles per FP64 warp-instruction\n", name, ms, ms * 1e-3 * 1.965e9 / fp64_per_smsp);
    cudaFree(sink);
}

int main() {
    run<0, 0>("DFMA r,U,U");
    run<0, 8>("DFMA r,U,U + 1 FFMA each");
    run<1, 0>("DFMA r,r,r");
    run<1, 4>("DFMA r,r,r + 0.5 FFMA each");
    run<1, 8>("DFMA r,r,r + 1 FFMA each");
    run<2, 0>("DADD r,r");
    run<2, 8>("DADD r,r + 1 FFMA each");
    run<3, 0>("DADD r,U");
    run<3, 8>("DADD r,U + 1 FFMA each");
    return 0;
}
