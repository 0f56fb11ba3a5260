// pipe_probe3.cu -- issue cost of FP64 instructions whose register operand repeats in consecutive instructions
// (the operand reuse cache, SASS ".reuse") and of the instruction kinds of the pair kernel's chain.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: a[k] = a[k] + x            one register operand shared by all (reuse candidates)
// MODE 1: a[k] = a[k] + x[k & 1]     shared by every second instruction
// MODE 2: a[k] = a[k] * a[k]         DMUL, one register pair
// MODE 3: p = |a[k]| > U ; a[k] = p ? a[k] : -a[k]  (DSETP r,U + cheap consumer)
// MODE 4: a[k] = a[k] + x[k]         distinct (reference: 2 issue cycles)
template <int MODE, int NF>
__global__ void __launch_bounds__(256) probe(double *sink, int iters, double b, float fb) {
    double a[8], x[8];
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = 1.0 + 0.001 * k + 1e-6 * threadIdx.x;
        x[k] = 1e-9 * (k + 1 + threadIdx.x);
        f[k] = 1.0f + 0.01f * k + threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) a[k] = __dadd_rn(a[k], x[0]);
                if (MODE == 1) a[k] = __dadd_rn(a[k], x[k & 1]);
                if (MODE == 2) a[k] = __dmul_rn(a[k], a[k]);
                if (MODE == 3) {
                    if (fabs(a[k]) > b) f[k] = -f[k];
                }
                if (MODE == 4) a[k] = __dadd_rn(a[k], x[k]);
                if (k < NF) f[k] = fmaf(f[k], fb, 0.5f);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + f[k] + x[k];
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
    probe<MODE, NF><<<blocks, 256>>>(sink, 100, 1.0000000001, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE, NF><<<blocks, 256>>>(sink, iters, 1.0000000001, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fp64_per_smsp = 16.0 * iters * 32.0;
    printf("%-44s %8.2f ms  %.2f cycles per FP64 warp-instruction\n", name, ms, ms * 1e-3 * 1.965e9 / fp64_per_smsp);
    cudaFree(sink);
}

int main() {
    run<4, 0>("DADD r,r distinct");
    run<4, 8>("DADD r,r distinct + 1 FFMA each");
    run<0, 0>("DADD r,x (x shared by all)");
    run<0, 8>("DADD r,x (x shared by all) + 1 FFMA each");
    run<1, 0>("DADD r,x[k&1]");
    run<1, 8>("DADD r,x[k&1] + 1 FFMA each");
    run<2, 0>("DMUL r,r (same register)");
    run<2, 8>("DMUL r,r (same register) + 1 FFMA each");
    run<3, 0>("DSETP |r|,U + predicated FADD");
    run<3, 8>("DSETP |r|,U + predicated FADD + 1 FFMA each");
    return 0;
}
