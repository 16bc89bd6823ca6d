// Warp-per-instance kernel with the LU rows in registers (kernel_rows.cuh): launchers.
#include <cuda_runtime.h>

#include <string>

#include "hostmodel.h"
#include "kernel_rows.cuh"
#include "kernel_rowsg.cuh"

#include <cstdlib>

using namespace acme;

// BASELINE config 4: examples/superover.jl with the three potentiometers as inputs
using RowsSuperover = CoopStatic<11, 4, 1, 13, 29, 11, 8, 23>;

bool rows_matches(const DevModel& dm) { return RowsSuperover::matches(dm); }

template <class S, int WARPS>
static cudaError_t launch_rows(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    const size_t smem = rows_smem_bytes<S>(WARPS, m->dm.nconst);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_rows<S, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k_rows<S, WARPS><<<(unsigned)((a.ninst + WARPS - 1) / WARPS), WARPS * 32, smem, stream>>>(m->dm, a);
    return cudaGetLastError();
}

template <class S, int WARPS, int G, int MINB>
static cudaError_t launch_rowsg(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    const size_t smem = rowsg_smem_bytes<S, G>(WARPS, m->dm.nconst);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_rowsg<S, WARPS, G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int64_t per_cta = (int64_t)WARPS * G;
    k_rowsg<S, WARPS, G, MINB><<<(unsigned)((a.ninst + per_cta - 1) / per_cta), WARPS * 32, smem, stream>>>(m->dm, a);
    return cudaGetLastError();
}

cudaError_t launch_rows_kernel(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    // tuning knob (experiments): ACMEB200_ROWS_VARIANT = g<G>w<WARPS>, e.g. g2w1, g2w4, g1w1
    if (const char* v = getenv("ACMEB200_ROWS_VARIANT")) {
        const std::string s(v);
        if (s == "g1w1") return launch_rowsg<RowsSuperover, 1, 1, 8>(m, a, stream);
        if (s == "g2w1") return launch_rowsg<RowsSuperover, 1, 2, 8>(m, a, stream);
        if (s == "g2w4") return launch_rowsg<RowsSuperover, 4, 2, 4>(m, a, stream);
        if (s == "g2w4r") return launch_rowsg<RowsSuperover, 4, 2, 3>(m, a, stream);
        if (s == "g2w2") return launch_rowsg<RowsSuperover, 2, 2, 4>(m, a, stream);
    }
    // small batches: one warp per CTA spreads the instances evenly over the SMs
    if (a.ninst <= 148 * 16) return launch_rows<RowsSuperover, 1>(m, a, stream);
    return launch_rows<RowsSuperover, 4>(m, a, stream);
}
