// second round: lane-varying shuffles, argmax variants, partial-mask votes
#include <cstdio>
#include <cuda_runtime.h>
#define TIMED(...) long long t0 = clock64(); for (int it = 0; it < iters; it++) { _Pragma("unroll") for (int k = 0; k < 8; k++) { __VA_ARGS__ } } long long t1 = clock64(); if (threadIdx.x == 0) *cyc = t1 - t0;
__global__ void k_shfl(double* out, long long* cyc, int iters) {
    double x = threadIdx.x; const int lane = threadIdx.x;
    TIMED(x = __shfl_sync(0xffffffffu, x, (lane + k + 1) & 31) + lane;)
    out[threadIdx.x] = x;
}
__global__ void k_shfl_half(double* out, long long* cyc, int iters) {
    double x = threadIdx.x; const int lane = threadIdx.x & 15; const unsigned m = threadIdx.x < 16 ? 0xffffu : 0xffff0000u;
    TIMED(x = __shfl_sync(m, x, (lane + k + 1) & 15, 16) + lane;)
    out[threadIdx.x] = x;
}
template <int L> __global__ void k_argmax_bfly(double* out, long long* cyc, int iters) {
    double x = threadIdx.x * 0.37; const int lane = threadIdx.x & (L - 1);
    const unsigned m = L == 32 ? 0xffffffffu : (threadIdx.x < 16 ? 0xffffu : 0xffff0000u);
    TIMED(
        double best = fabs(x - 7.0 - k); int bi = lane;
        _Pragma("unroll") for (int o = L / 2; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(m, best, o, L); const int oi = __shfl_xor_sync(m, bi, o, L);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        x = __shfl_sync(m, x, bi, L) + lane;
    )
    out[threadIdx.x] = x;
}
__global__ void k_argmax_redux(double* out, long long* cyc, int iters) {
    double x = threadIdx.x * 0.37; const int lane = threadIdx.x;
    TIMED(
        const double a = fabs(x - 7.0 - k);
        const unsigned hi = __double2hiint(a), lo = __double2loint(a);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const bool c1 = hi == mh;
        const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
        const bool c2 = c1 && lo == ml;
        const unsigned src = __ffs(__ballot_sync(0xffffffffu, c2)) - 1;
        x = __shfl_sync(0xffffffffu, x, src) + lane;
    )
    out[threadIdx.x] = x;
}
__global__ void k_vote_half(double* out, long long* cyc, int iters) {
    unsigned x = threadIdx.x; const unsigned m = threadIdx.x < 16 ? 0xffffu : 0xffff0000u;
    TIMED(x = __ballot_sync(m, (x >> (threadIdx.x & 7)) & 1) + threadIdx.x;)
    out[threadIdx.x] = x;
}
__global__ void k_syncwarp_half(double* out, long long* cyc, int iters) {
    __shared__ double buf[64];
    double x = threadIdx.x; const unsigned m = threadIdx.x < 16 ? 0xffffu : 0xffff0000u; const int g = threadIdx.x >> 4;
    TIMED(
        if ((threadIdx.x & 15) == ((k * 7 + 3) & 15)) buf[g * 32 + k] = x;
        __syncwarp(m);
        x = buf[g * 32 + k] + 1.0;
        __syncwarp(m);
    )
    out[threadIdx.x] = x;
}
__global__ void k_rcp(double* out, long long* cyc, int iters) {
    double x = 1.0 + threadIdx.x * 1e-3;
    TIMED(x = 1.0 / x + 0.5;)
    out[threadIdx.x] = x;
}
__global__ void k_rcp_fast(double* out, long long* cyc, int iters) {  // MUFU.RCP64H seed + 2 Newton steps, no special cases
    double x = 1.0 + threadIdx.x * 1e-3;
    TIMED(
        double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); y = fma(y, e, y);
        x = y + 0.5;
    )
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 16); cudaMallocManaged(&cyc, 8);
    const int iters = 2000; const double n = iters * 8.0;
    k_shfl<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("shfl f64 lane-varying + DADD: %.2f\n", *cyc / n);
    k_shfl_half<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("shfl f64 width16, two masks + DADD: %.2f\n", *cyc / n);
    k_argmax_bfly<32><<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("argmax butterfly L=32 + broadcast: %.2f\n", *cyc / n);
    k_argmax_bfly<16><<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("argmax butterfly L=16 (two groups) + broadcast: %.2f\n", *cyc / n);
    k_argmax_redux<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("argmax 2xredux+ballot L=32 + broadcast: %.2f\n", *cyc / n);
    k_vote_half<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("ballot with two 16-lane masks: %.2f\n", *cyc / n);
    k_syncwarp_half<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("STS->syncwarp(half)->LDS->syncwarp(half): %.2f\n", *cyc / n);
    k_rcp<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("1.0/x + a: %.2f\n", *cyc / n);
    k_rcp_fast<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); printf("rcp.approx + 2 Newton + a: %.2f\n", *cyc / n);
    return 0;
}
