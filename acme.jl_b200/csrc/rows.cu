// Warp-per-instance kernel with the LU rows in registers (kernel_rows.cuh): launchers.
#include <cuda_runtime.h>

#include "hostmodel.h"
#include "kernel_rows.cuh"

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

cudaError_t launch_rows_kernel(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    // small batches: one warp per CTA spreads the instances evenly over the SMs and may use 255
    // registers (8 resident warps per SM); larger batches need the 16 warps per SM of the 128-register build
    if (a.ninst <= 148 * 8) return launch_rows<RowsSuperover, 1>(m, a, stream);
    return launch_rows<RowsSuperover, 4>(m, a, stream);
}
