// Cooperative kernel (kernel_coop.cuh): lane-count choice and launchers.
#include <cuda_runtime.h>

#include <algorithm>

#include "hostmodel.h"
#include "kernel_coop.cuh"

using namespace acme;

// compile-time shapes of the cooperative kernel (BASELINE config 4: examples/superover.jl with the
// three potentiometers as inputs: nx 11, nu 4, ny 1, nn 13, nq 29, np 11, 8 elements, 23 jv entries)
using CoopSuperover = CoopStatic<11, 4, 1, 13, 29, 11, 8, 23>;

bool coop_static_matches(const DevModel& dm) { return CoopSuperover::matches(dm); }

int coop_lanes_for(const acmeb200_model* m) {
    if (m->blob_stride != 0 || m->dm.nsub == 0 || m->max_nn > MAX_ROWS || !m->rows_ok) return 0;
    const int need = std::max(m->max_nn, m->max_nelem);
    int lanes = need <= 8 ? 8 : (need <= 16 ? 16 : 32);
    // small batches are latency-bound: one instance per warp avoids the two groups of a warp
    // serialising when their Newton iteration counts differ, and doubles the warps in flight
    if (m->B * lanes < (int64_t)148 * 32 * 24) lanes = std::min(32, lanes * 2);
    size_t smem = lanes == 8 ? coop_smem_bytes<8>(m->dm) : lanes == 16 ? coop_smem_bytes<16>(m->dm) : coop_smem_bytes<32>(m->dm);
    while (smem > 200 * 1024 && lanes < 32) {  // fewer groups per CTA
        lanes *= 2;
        smem = lanes == 16 ? coop_smem_bytes<16>(m->dm) : coop_smem_bytes<32>(m->dm);
    }
    return smem <= 200 * 1024 ? lanes : 0;
}

template <int L, class P>
static cudaError_t launch_coop(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    const size_t smem = coop_smem_bytes<L>(m->dm);
    static bool attr_set_dev[ACME_MAX_DEVICES] = {};  // function attributes are per device: one process may drive several
    bool& attr_set = attr_set_dev[m->device & (ACME_MAX_DEVICES - 1)];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_coop<L, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    constexpr int GPC = COOP_TPB / L;
    ACME_LAUNCH((k_coop<L, P>), (unsigned)((a.ninst + GPC - 1) / GPC), COOP_TPB, smem, stream, m->dm, a);
    return cudaGetLastError();
}

cudaError_t launch_coop_kernel(const acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    if (m->coop_static == 1) {
        if (m->coop_lanes == 16) return launch_coop<16, CoopSuperover>(m, a, stream);
        return launch_coop<32, CoopSuperover>(m, a, stream);
    }
    if (m->coop_lanes == 8) return launch_coop<8, CoopDyn>(m, a, stream);
    if (m->coop_lanes == 16) return launch_coop<16, CoopDyn>(m, a, stream);
    return launch_coop<32, CoopDyn>(m, a, stream);
}
