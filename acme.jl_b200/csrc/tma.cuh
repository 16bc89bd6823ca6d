// Small shared device utilities: compile-time loops and the TMA (bulk async copy) + mbarrier
// primitives (PTX ISA 8.x) used by all kernels.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudart)
#include <cstdint>

namespace acme {

template <int V> struct IC { static constexpr int value = V; };

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(IC<I>{});
        static_for<I + 1, N>(f);
    }
}


#ifdef ACME_HOST_EMU
// ---- host emulation (tests/emu): a "shared address" is the offset into the emulated shared memory; an
// mbarrier is its count of completed phases (every use here is one arrival + the bytes of one copy);
// bulk and tensor copies happen synchronously, so the bulk async-groups have nothing to wait for.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t) { *reinterpret_cast<uint64_t*>(acme_emu::g_smem + bar) = 0; }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t, uint32_t) {}
__device__ __forceinline__ void mbar_arrive(uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // phase with this parity completed?
    while ((*reinterpret_cast<volatile uint64_t*>(acme_emu::g_smem + bar) & 1u) == (uint64_t)parity) acme_emu::yield();
}
__device__ __forceinline__ void emu_complete_phase(uint32_t bar) {
    *reinterpret_cast<uint64_t*>(acme_emu::g_smem + bar) += 1;
    acme_emu::g_cta.progress++;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    memcpy(acme_emu::g_smem + dst, src, bytes);
    emu_complete_phase(bar);
}
// tile <-> shared memory with the descriptor's swizzle (16-byte chunks XOR-permuted by address bits 7..9),
// zero fill on loads and clipping on stores outside the tensor -- what the TMA unit does
__device__ __forceinline__ uint32_t emu_swizzle(const CUtensorMap* tm, uint32_t off) { return off ^ (((off >> 7) & tm->swizzle_mask) << 4); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    for (unsigned r = 0; r < tm->box1; r++)
        for (unsigned c = 0; c < tm->box0; c++) {
            const long long i0 = (long long)c0 + c, i1 = (long long)c1 + r;
            double v = 0.0;
            if (i0 >= 0 && i1 >= 0 && (unsigned long long)i0 < tm->dim0 && (unsigned long long)i1 < tm->dim1)
                memcpy(&v, tm->base + (unsigned long long)i1 * tm->stride1_bytes + (unsigned long long)i0 * 8, 8);
            memcpy(acme_emu::g_smem + dst + emu_swizzle(tm, (r * tm->box0 + c) * 8), &v, 8);
        }
    emu_complete_phase(bar);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
    for (unsigned r = 0; r < tm->box1; r++)
        for (unsigned c = 0; c < tm->box0; c++) {
            const long long i0 = (long long)c0 + c, i1 = (long long)c1 + r;
            if (i0 >= 0 && i1 >= 0 && (unsigned long long)i0 < tm->dim0 && (unsigned long long)i1 < tm->dim1)
                memcpy(tm->base + (unsigned long long)i1 * tm->stride1_bytes + (unsigned long long)i0 * 8,
                       acme_emu::g_smem + src + emu_swizzle(tm, (r * tm->box0 + c) * 8), 8);
        }
}
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read1() {}
__device__ __forceinline__ void bulk_wait_read0() {}
__device__ __forceinline__ void bulk_wait_all() {}
__device__ __forceinline__ void fence_async_smem() {}
#else
// ---- TMA (bulk async copy) + mbarrier primitives, PTX ISA 8.x --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the mbarrier initialisation visible to the async proxy / the cluster
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy of a contiguous range (TMA, UBLKCP in SASS), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// One 2D tile of a (instances x samples) stream, global -> shared, through a tensor map
// (TMA, UTMALDG in SASS); completion is counted on an mbarrier.  Out-of-range rows/columns of the
// box are zero-filled, so partial tiles and partially filled warps need no special case.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// shared -> global tile store (UTMASTG), tracked by the issuing thread's bulk async-group;
// the part of the box outside the tensor is not written
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0),
                 "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


#endif  // ACME_HOST_EMU

}  // namespace acme
