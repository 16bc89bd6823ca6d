// Host emulation of the slice of CUDA the device library uses -- TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py compiles acme.jl_b200/csrc/{acmeb200,tpi,rows}.cu with g++ against THIS header
// (found as <cuda_runtime.h>) into tests/emu/_build/libacmeb200_emu.so, so that the kernels' logic --
// the very source the GPU runs -- can be executed and checked against the oracle on a machine without
// a GPU (tests/test_emu.py).  The product never loads it: acme.jl_b200/_lib.py loads
// libacmeb200.so and fails loudly without a CUDA device.
//
// Model: one CTA at a time; every CUDA thread is a fiber (ucontext) in one OS thread, switched only at
// collectives, so warp-synchronous code sees exactly the lock-step semantics it relies on.  Warp collectives
// synchronise the lanes of their mask (full warps, or the cooperative kernel's sub-warp groups); tensor-map tiles,
// bulk copies and mbarrier phases are emulated synchronously (csrc/tma.cuh, cuda.h here).
#pragma once
#define ACME_HOST_EMU 1
#define __CUDACC__ 1
#define __CUDA_ARCH__ 1000

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ /* system headers spell the attribute __attribute__((__noinline__)): leave the inner name empty */
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__
#define __shared__
#define __align__(n)

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct emu_uint3 { unsigned x, y, z; };
extern emu_uint3 threadIdx, blockIdx, blockDim, gridDim;

// ------------------------------------------------------------------ fibers
namespace acme_emu {
constexpr int MAX_THREADS = 1024;
struct Cta {
    int nthreads = 0, live = 0;
    unsigned warp_gen[MAX_THREADS / 32], warp_cnt[MAX_THREADS / 32], warp_live[MAX_THREADS / 32];
    // sub-warp groups (the cooperative kernel's lanes-per-instance masks): one barrier per (warp, mask)
    struct GroupBar { unsigned mask, gen, cnt; } group[MAX_THREADS / 32][4];
    unsigned live_mask[MAX_THREADS / 32];
    unsigned cta_gen = 0, cta_cnt = 0;
    uint64_t slot[MAX_THREADS];
    unsigned long progress = 0;
};
extern Cta g_cta;
void yield();
void warp_barrier(unsigned mask = 0xffffffffu);
void cta_barrier();
void run_cta(const std::function<void()>& body, unsigned block, unsigned bx, unsigned grid);
extern unsigned char* g_smem;  // dynamic shared memory of the running CTA

template <class K, class... A>
inline void launch(K kernel, unsigned grid, unsigned block, size_t /*smem*/, A... args) {
    for (unsigned b = 0; b < grid; b++) run_cta([&]() { kernel(args...); }, block, b, grid);
}
}  // namespace acme_emu

// ------------------------------------------------------------------ warp / CTA collectives (full mask)
inline void __syncthreads() { acme_emu::cta_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { acme_emu::warp_barrier(mask); }
template <class T>
inline T emu_exchange(T v, int src_lane_in_warp, unsigned mask = 0xffffffffu) {
    uint64_t bits = 0;
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    memcpy(&bits, &v, sizeof(T));
    const unsigned tid = threadIdx.x;
    acme_emu::g_cta.slot[tid] = bits;
    acme_emu::warp_barrier(mask);
    const uint64_t r = acme_emu::g_cta.slot[(tid & ~31u) | (unsigned)(src_lane_in_warp & 31)];
    acme_emu::warp_barrier(mask);
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = threadIdx.x & 31;
    return emu_exchange(v, (lane & ~(width - 1)) | (src & (width - 1)), mask);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int xr, int width = 32) {
    const int lane = threadIdx.x & 31;
    return emu_exchange(v, (lane & ~(width - 1)) | ((lane ^ xr) & (width - 1)), mask);
}
template <class F>
inline unsigned emu_reduce(unsigned mask, unsigned v, F f) {  // over the live lanes of `mask`
    const unsigned tid = threadIdx.x, base = tid & ~31u;
    acme_emu::g_cta.slot[tid] = v;
    acme_emu::warp_barrier(mask);
    const unsigned m = mask & acme_emu::g_cta.live_mask[tid >> 5];
    unsigned r = 0;
    bool first = true;
    for (unsigned l = 0; l < 32; l++)
        if (m >> l & 1u) {
            const unsigned x = (unsigned)acme_emu::g_cta.slot[base + l];
            r = first ? x : f(r, x);
            first = false;
        }
    acme_emu::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a + b; }); }
inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned tid = threadIdx.x, base = tid & ~31u;
    acme_emu::g_cta.slot[tid] = pred ? 1u : 0u;
    acme_emu::warp_barrier(mask);
    const unsigned m = mask & acme_emu::g_cta.live_mask[tid >> 5];
    unsigned r = 0;
    for (unsigned l = 0; l < 32; l++)
        if (m >> l & 1u) r |= (unsigned)acme_emu::g_cta.slot[base + l] << l;
    acme_emu::warp_barrier(mask);
    return r;
}
inline int __all_sync(unsigned m, int pred) { return emu_reduce(m, pred ? 1u : 0u, [](unsigned a, unsigned b) { return a & b; }) != 0; }
inline int __any_sync(unsigned m, int pred) { return emu_reduce(m, pred ? 1u : 0u, [](unsigned a, unsigned b) { return a | b; }) != 0; }

// ------------------------------------------------------------------ device intrinsics
inline int __double2hiint(double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b >> 32); }
inline int __double2loint(double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b & 0xffffffff); }
inline unsigned __float_as_uint(float v) { unsigned b; memcpy(&b, &v, 4); return b; }
inline float __uint_as_float(unsigned b) { float v; memcpy(&v, &b, 4); return v; }
inline double __hiloint2double(int hi, int lo) { const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; memcpy(&v, &b, 8); return v; }
inline double __longlong_as_double(long long b) { double v; memcpy(&v, &b, 8); return v; }
inline long long __double_as_longlong(double v) { long long b; memcpy(&b, &v, 8); return b; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }   // the volatile keeps the product
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }   // from being contracted into an FMA
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline void __threadfence_block() {}
inline void __threadfence() {}
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const unsigned char*)p - acme_emu::g_smem); }
using std::isfinite;

// ------------------------------------------------------------------ runtime API (host memory stands in for device memory)
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
       cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaDevAttrMultiProcessorCount = 16,
       cudaDevAttrMaxSharedMemoryPerMultiprocessor = 81 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; r++) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0 };
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* q);  // emu_runtime.cpp
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) { *v = attr == cudaDevAttrMultiProcessorCount ? 148 : 233472; return cudaSuccess; }
