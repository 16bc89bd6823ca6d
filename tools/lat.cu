// dependent-chain latency / throughput microbenchmarks for FP64 on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, int iters, double a) {
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x = 1.0 / x + a;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_exp(double* out, long long* cyc, int iters, double a) {
    double x = 0.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) x = exp(x * a) * 0.25;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 8);
    const int iters = 1000;
    // single warp: pure latency
    k_chain<1><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3); cudaDeviceSynchronize();
    printf("DFMA dependent latency (1 warp, ILP1): %.2f cycles\n", (double)*cyc / (iters * 16));
    k_chain<2><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3); cudaDeviceSynchronize();
    printf("DFMA ILP2 per-instruction: %.2f cycles\n", (double)*cyc / (iters * 32));
    k_chain<4><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3); cudaDeviceSynchronize();
    printf("DFMA ILP4 per-instruction: %.2f cycles\n", (double)*cyc / (iters * 64));
    k_chain<8><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3); cudaDeviceSynchronize();
    printf("DFMA ILP8 per-instruction: %.2f cycles\n", (double)*cyc / (iters * 128));
    for (int w = 1; w <= 16; w *= 2) {  // warps per SM (one block on one SM), ILP1: how many warps saturate the pipe
        k_chain<1><<<1, 32 * w * 4>>>(out, cyc, iters, 0.999, 1e-3); cudaDeviceSynchronize();
        printf("%2d warps/scheduler ILP1: %.2f cycles per DFMA per warp -> %.2f DFMA/cycle/scheduler\n", w, (double)*cyc / (iters * 16), w * (iters * 16.0) / *cyc);
    }
    k_rcp<<<1, 32>>>(out, cyc, iters, 0.5); cudaDeviceSynchronize();
    printf("1.0/x + a dependent latency: %.2f cycles\n", (double)*cyc / (iters * 8));
    k_exp<<<1, 32>>>(out, cyc, iters, 0.7); cudaDeviceSynchronize();
    printf("exp(x*a)*c dependent latency: %.2f cycles\n", (double)*cyc / (iters * 4));
    return 0;
}
