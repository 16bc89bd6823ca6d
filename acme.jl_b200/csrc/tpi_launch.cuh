// Launch side of the thread-per-instance kernels (kernel_tpi.cuh): tensor maps of a launch's streams, the launcher
// and the registry entry of one compile-time shape.  A header, because it has two users: tpi.cu (the shapes built into
// the library) and the one-shape plugin libraries that acme_jl_b200.specialise generates for a model whose shape the
// library does not carry (csrc/shape_plugin.cu.in; registered at run time through acmeb200_register_tpi).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hostmodel.h"
#include "elements.cuh"
#include "kernel_tpi.cuh"

namespace acme {

template <class C>
static void fill_tpi_mats(const DevModel& dm, const double* blob, TpiMats<C>& M) {
    memset(&M, 0, sizeof M);
    auto cp = [&](double* dst, int off, int n) { for (int i = 0; i < n; i++) dst[i] = blob[off + i]; };
    cp(M.a, dm.o_a, C::NX * C::NX); cp(M.b, dm.o_b, C::NX * C::NU); cp(M.c, dm.o_c, C::NX * C::NN);
    cp(M.x0, dm.o_x0, C::NX);
    cp(M.dy, dm.o_dy, C::NY * C::NX); cp(M.ey, dm.o_ey, C::NY * C::NU); cp(M.fy, dm.o_fy, C::NY * C::NN);
    cp(M.y0, dm.o_y0, C::NY);
    if (C::NN > 0) {
        const DevSub& s = dm.subs[0];
        cp(M.dq, s.o_dq, C::NP * C::NX); cp(M.eq, s.o_eq, C::NP * C::NU);
        cp(M.pexp, s.o_pexp, C::NQ * C::NP); cp(M.q0, s.o_q0, C::NQ); cp(M.fq, s.o_fq, C::NQ * C::NN);
    }
}

// cuTensorMapEncodeTiled, fetched through the runtime (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Tensor map of one f64 stream of a launch: dim0 = the `inner` contiguous values of an instance,
// dim1 = `ninst` instances `stride` values apart, box {box_inner, box_outer = 32} = one warp tile.
// Sample-major streams swap the roles: dim0 = the channels of all instances at one sample, dim1 = samples
// `stride` values apart, box {32*channels, TPI_T}.
// False when the stream does not meet TMA's 16-byte alignment rules (the kernel then uses its
// synchronous path).
static bool make_tile_map(CUtensorMap* tm, const double* base, int64_t stride, int64_t ninst, int64_t inner,
                          int box_inner, int box_outer = 32) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !base || inner <= 0 || ninst <= 0) return false;
    if (ninst == 1 && stride < inner) stride = (inner + 1) & ~int64_t(1);  // a single row: the pitch is unused
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride & 1) || stride < inner || stride >= (int64_t(1) << 36)) return false;
    if ((box_inner & 1) || box_inner > 256 || inner >= (int64_t(1) << 32) || ninst >= (int64_t(1) << 32)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)ninst};
    const cuuint64_t strides[1] = {(cuuint64_t)stride * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    // the swizzle the kernel's tile accessors assume (tpi_swizzle_mask): by the row length of the box
    const int mask = tpi_swizzle_mask(box_inner * 8);
    const CUtensorMapSwizzle sw = mask == 7 ? CU_TENSOR_MAP_SWIZZLE_128B : mask == 3 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : mask == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// attributes, carve-out and the launch of one register-budget variant (WIDE: see kernel_tpi.cuh)
template <class C, bool WIDE>
cudaError_t launch_tpi_k(const acmeb200_model* m, const RunArgs& a, const TpiMats<C>& M, const SolverCfg& sc, const DevSub& cache,
                                const TpiMaps& maps, int64_t blocks, int sms, cudaStream_t stream) {
    constexpr int MINB = WIDE ? ACME_TPI_WIDE_MINB : ACME_TPI_MINB;
    const size_t smem = tpi_smem_bytes<C>();
    const int dev = m->device & (ACME_MAX_DEVICES - 1);  // function attributes are per device: one process may drive several
    if (smem > 48 * 1024) {  // long tiles: opt in to the large dynamic shared memory carve-out
        static bool attr_set_dev[ACME_MAX_DEVICES] = {};
        bool& attr_set = attr_set_dev[dev];
        if (!attr_set) {
            for (auto* k : {k_tpi<C, true, false, WIDE>, k_tpi<C, false, false, WIDE>, k_tpi<C, true, true, WIDE>, k_tpi<C, false, true, WIDE>}) {
                const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
            }
            attr_set = true;
        }
    }
    {
        // Shared-memory carve-out: only what the CTAs resident for THIS batch need, the rest of the
        // unified array stays L1.  The learning cache's stored points are scanned from global memory
        // every sample (config 5: 124 KB per SM); with the default maximum carve-out (200 KB) only 39 % of
        // those loads hit L1 and the kernel waits on L2 latency (profiles/k_tpi_r1.md).
        static int last_pct_dev[ACME_MAX_DEVICES], max_smem_dev[ACME_MAX_DEVICES] = {};
        static bool seen_dev[ACME_MAX_DEVICES] = {};
        int &last_pct = last_pct_dev[dev], &max_smem = max_smem_dev[dev];
        if (!seen_dev[dev]) {
            cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, m->device);
            last_pct = -1;
            seen_dev[dev] = true;
        }
        const int64_t resident = std::min<int64_t>(MINB, (blocks + sms - 1) / std::max(sms, 1));
        const int64_t need = resident * (int64_t)(smem + 1024);
        const int pct = (int)std::min<int64_t>(100, (need * 100 + max_smem - 1) / std::max(max_smem, 1));
        if (pct != last_pct) {
            for (auto* k : {k_tpi<C, true, false, WIDE>, k_tpi<C, false, false, WIDE>, k_tpi<C, true, true, WIDE>, k_tpi<C, false, true, WIDE>})
                cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            last_pct = pct;
        }
    }
    if (a.smaj) {  // sample-major streams: transposed tiles (kernel_tpi.cuh)
        if (m->blob_stride)
            ACME_LAUNCH((k_tpi<C, true, true, WIDE>), (unsigned)blocks, TPI_TPB, smem, stream, M, a, sc, cache, maps);
        else
            ACME_LAUNCH((k_tpi<C, false, true, WIDE>), (unsigned)blocks, TPI_TPB, smem, stream, M, a, sc, cache, maps);
    } else if (m->blob_stride)
        ACME_LAUNCH((k_tpi<C, true, false, WIDE>), (unsigned)blocks, TPI_TPB, smem, stream, M, a, sc, cache, maps);
    else
        ACME_LAUNCH((k_tpi<C, false, false, WIDE>), (unsigned)blocks, TPI_TPB, smem, stream, M, a, sc, cache, maps);
    return cudaGetLastError();
}

template <class C>
static cudaError_t launch_tpi(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    TpiMats<C> M;
    fill_tpi_mats<C>(m->dm, m->h_blob.data(), M);
    SolverCfg sc{m->dm.tol, m->dm.maxiter, m->dm.solver, make_exp_table()};
    DevSub cache;
    memset(&cache, 0, sizeof cache);
    if (m->dm.nsub > 0) cache = m->dm.subs[0];
    TpiMaps maps;
    memset(&maps, 0, sizeof maps);
    if (!a.init && a.smaj) {
        maps.in_ok = C::NU > 0 && a.u_stride != 0 &&
                     make_tile_map(&maps.u, a.U, a.u_stride, a.N, a.ninst * C::NU, 32 * C::NU, TPI_T);
        maps.out_ok = C::NY > 0 && make_tile_map(&maps.y, a.Y, a.y_stride, a.N, a.ninst * C::NY, 32 * C::NY, TPI_T);
    } else if (!a.init) {
        maps.in_ok = C::NU > 0 && a.u_stride != 0 &&
                     make_tile_map(&maps.u, a.U, a.u_stride, a.ninst, a.N * C::NU, TPI_T * C::NU);
        maps.out_ok = C::NY > 0 && make_tile_map(&maps.y, a.Y, a.y_stride, a.ninst, a.N * C::NY, TPI_T * C::NY);
    }
    const int64_t blocks = (a.ninst + TPI_TPB - 1) / TPI_TPB;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
#ifndef ACME_HOST_EMU
    if constexpr (C::NN > 0) {  // small batches of non-linear shapes: the 255-register build (kernel_tpi.cuh, ACME_TPI_WIDE_MINB)
        // (both stream layouts take the same build: they are bit-identical to each other, the two builds only to rounding)
        bool wide = blocks <= (int64_t)ACME_TPI_WIDE_MINB * std::max(sms, 1);
        if (const char* e = getenv("ACMEB200_TPI_WIDE")) wide = atoi(e) != 0;  // tests, tuning
        if (wide) return launch_tpi_k<C, true>(m, a, M, sc, cache, maps, blocks, sms, stream);
    }
#endif
    return launch_tpi_k<C, false>(m, a, M, sc, cache, maps, blocks, sms, stream);
}

template <class C>
static TpiEntry make_entry(const char* name) {
    return TpiEntry{name, C::NX, C::NU, C::NY, C::NP, C::NE, C::kinds, C::S_ROWS, &launch_tpi<C>};
}


}  // namespace acme
