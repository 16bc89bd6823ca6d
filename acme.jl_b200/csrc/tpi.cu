// Thread-per-instance kernels (kernel_tpi.cuh): registry of the compile-time shapes, tensor
// maps of the launch's streams, launchers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <deque>
#include <vector>

#include "tpi_shapes.h"

using namespace acme;

#ifndef ACME_HOST_EMU
// the 255-register build of the non-linear shapes lives in tpi_wide.cu
#define X(CFG) extern template cudaError_t acme::launch_tpi_k<CFG, true>(const acmeb200_model*, const RunArgs&, const TpiMats<CFG>&, const SolverCfg&, const DevSub&, const TpiMaps&, int64_t, int, cudaStream_t);
ACME_TPI_WIDE_SHAPES(X)
#undef X
#endif

static std::deque<TpiEntry>& tpi_registry() {  // a deque: models keep pointers to entries, registration must not move them
    static std::deque<TpiEntry> reg = {
        make_entry<CfgDiodeClipper>("tpi<diodeclipper nx1 nu1 ny1 np1 [diode,diode]>"),
        make_entry<CfgSallenKey>("tpi<linear nx2 nu1 ny1>"),
        make_entry<CfgBirdieFixed>("tpi<birdie nx3 nu1 ny1 np2 [bjt]>"),
        make_entry<CfgBirdieVol>("tpi<birdie nx3 nu2 ny1 np3 [bjt,pot]>"),
    };
    return reg;
}

// A shape compiled outside the library (acme_jl_b200.specialise: one translation unit with one TpiCfg, built by nvcc
// into its own shared object) joins the registry; later models of that shape run on its kernel.  The entry must stay
// alive (it lives in the plugin library, which is never unloaded).
extern "C" int acmeb200_register_tpi(const void* entry_) {
    const TpiEntry* e = static_cast<const TpiEntry*>(entry_);
    if (!e || !e->launch || !e->name) return ACMEB200_EINVAL;
    for (const TpiEntry& r : tpi_registry())
        if (r.launch == e->launch) return ACMEB200_OK;  // already there
    tpi_registry().push_back(*e);
    return ACMEB200_OK;
}

const TpiEntry* find_tpi(const DevModel& dm) {
    if (dm.nsub > 1) return nullptr;
    for (const TpiEntry& e : tpi_registry()) {
        if (e.nx != dm.nx || e.nu != dm.nu || e.ny != dm.ny) continue;
        if (dm.nsub == 0) {
            if (e.ne == 0) return &e;
            continue;
        }
        const DevSub& s = dm.subs[0];
        if (e.ne != s.nelem || e.np != s.np) continue;
        bool same = true;
        for (int k = 0; k < e.ne; k++) same = same && dm.elems[s.elem0 + k].kind == e.kinds[k];
        if (same) return &e;
    }
    return nullptr;
}


#ifdef ACME_TPI_PROF
// tuning builds only: the per-warp records of the last nonlinear thread-per-instance launch (kernel_tpi.cuh)
extern "C" int acmeb200_diag_tpi_prof(unsigned long long* out, int n_warps) {
    if (n_warps > TPI_PROF_WARPS) n_warps = TPI_PROF_WARPS;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_tpi_prof, sizeof(unsigned long long) * TPI_PROF_REC * n_warps);
}
#endif
