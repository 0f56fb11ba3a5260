// pipe_probe.cu -- does an FP64 warp instruction block the issue port for its 2 pipe cycles on B200?
// Each thread runs 8 independent DFMA chains; per 8 DFMAs it also issues N independent FFMA (fma
// pipe), LOP3 (alu pipe) or a mix.  If co-issue works, time stays flat until N ~ 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI>
__global__ void __launch_bounds__(256) probe(double *sink, int iters, float fb, unsigned ib) {
    double a[8];
    float f[8];
    unsigned u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = 1.0 + 0.001 * k + 1e-6 * threadIdx.x;
        f[k] = 1.0f + 0.01f * k + threadIdx.x;
        u[k] = threadIdx.x * 2654435761u + k;
    }
    const double b = 1.0000000001, c = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] = fma(a[k], b, c);
                if (k < NF) f[k] = fmaf(f[k], fb, 0.5f);
                if (k < NI) u[k] = (u[k] ^ ib) + (u[k] >> 3);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + f[k] + u[k];
    if (s == 123.456) sink[0] = s;
}

template <int NF, int NI>
void run(const char *name) {
    double *sink;
    cudaMalloc(&sink, 8);
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = dev_sms * 8, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<NF, NI><<<blocks, 256>>>(sink, 100, 1.0001f, 0x5bd1e995u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<NF, NI><<<blocks, 256>>>(sink, iters, 1.0001f, 0x5bd1e995u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)blocks * 256 * iters * 32.0;
    printf("%-28s %8.2f ms  %.3e DFMA/s\n", name, ms, dfma / (ms * 1e-3));
    cudaFree(sink);
}

int main() {
    run<0, 0>("8 DFMA");
    run<2, 0>("8 DFMA + 2 FFMA");
    run<4, 0>("8 DFMA + 4 FFMA");
    run<8, 0>("8 DFMA + 8 FFMA");
    run<0, 2>("8 DFMA + 2x2 ALU");
    run<0, 4>("8 DFMA + 4x2 ALU");
    run<0, 8>("8 DFMA + 8x2 ALU");
    run<4, 4>("8 DFMA + 4 FFMA + 4x2 ALU");
    run<8, 8>("8 DFMA + 8 FFMA + 8x2 ALU");
    return 0;
}
