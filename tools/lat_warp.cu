// dependent-chain latencies of the warp-level exchange primitives on sm_100a
// (SHFL, REDUX, VOTE, LDS broadcast, STS->LDS round trip): design input for kernel_rows.cuh
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_shfl(double* out, long long* cyc, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __shfl_sync(0xffffffffu, x, (k * 7 + 3) & 31) + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl32(double* out, long long* cyc, int iters) {
    int x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __shfl_sync(0xffffffffu, x, (k * 7 + 3) & 31) + 1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_redux(double* out, long long* cyc, int iters) {
    unsigned x = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __reduce_max_sync(0xffffffffu, x ^ (threadIdx.x + k)) + 1u;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_redux_half(double* out, long long* cyc, int iters) {  // two 16-lane groups, different masks
    unsigned x = threadIdx.x * 2654435761u;
    const unsigned m = threadIdx.x < 16 ? 0x0000ffffu : 0xffff0000u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __reduce_max_sync(m, x ^ (threadIdx.x + k)) + 1u;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_vote(double* out, long long* cyc, int iters) {
    unsigned x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __ballot_sync(0xffffffffu, (x >> (threadIdx.x & 7)) & 1) + threadIdx.x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int iters) {  // pointer chase through shared memory
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) nxt[i] = (i * 17 + 5) & 1023;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) p = nxt[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc, int iters) {  // lane 5 writes, syncwarp, all read (broadcast)
    __shared__ double buf[64];
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (threadIdx.x == ((k * 7 + 3) & 31)) buf[k] = x;
            __syncwarp();
            x = buf[k] + 1.0;
            __syncwarp();
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmul_dadd(double* out, long long* cyc, int iters, double a) {  // non-fused pair chain
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x = __dsub_rn(x, __dmul_rn(x, a));
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 16); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    k_shfl<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("shfl.idx f64 (2xSHFL) + DADD chain: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_shfl32<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("shfl.idx b32 + IADD chain: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_redux<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("redux.max.u32 (+xor,add) chain: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_redux_half<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("redux.max.u32, two 16-lane masks per warp: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_vote<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("vote.ballot (+shift,and,add) chain: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_lds<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("LDS pointer chase: %.2f cycles per load\n", (double)*cyc / (iters * 16));
    k_sts_lds<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("STS(one lane) -> syncwarp -> LDS broadcast (+DADD) round trip: %.2f cycles per step\n", (double)*cyc / (iters * 16));
    k_dmul_dadd<<<1, 32>>>(out, cyc, iters, 1e-3); cudaDeviceSynchronize();
    printf("DMUL -> DADD dependent pair: %.2f cycles per pair\n", (double)*cyc / (iters * 16));
    return 0;
}
