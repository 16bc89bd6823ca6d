// Host-side model object and the launch entry points of the per-kernel translation units
// (tpi.cu, coop.cu, rows.cu are compiled separately so that the library builds in parallel).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/acmeb200.h"
#include "devmodel.h"

struct acmeb200_model;

constexpr int ACME_MAX_DEVICES = 64;  // per-device launch state (function attributes) is kept in tables of this size

// kernel launch: CUDA's <<<>>> in the product; the host emulation of tests/emu runs the CTAs as fibers
#ifdef ACME_HOST_EMU
#define ACME_LAUNCH(kernel, grid, block, smem, stream, ...) acme_emu::launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__)
#else
#define ACME_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

struct TpiEntry {
    const char* name;
    int nx, nu, ny, np, ne;
    const int* kinds;
    int state_rows;
    cudaError_t (*launch)(const acmeb200_model*, const acme::RunArgs&, cudaStream_t);
};


struct acmeb200_model {
    acme::DevModel dm;
    int64_t B = 0;
    int device = 0;
    // device arrays
    double* d_blob = nullptr;
    int64_t blob_stride = 0;
    double* d_consts = nullptr;
    double* d_initz = nullptr;
    double* d_ws = nullptr;      // state/workspace of the selected kernel
    int64_t ws_rows = 0;
    uint32_t* d_status = nullptr;
    long long* d_first_fail = nullptr;
    acme::DevStats* d_stats = nullptr;
    std::vector<void*> d_cache;  // frozen, host-built solution stores (one image per sub-problem)
    std::vector<void*> d_dyn;    // learning per-instance solution stores and their scratch (kdcache.cuh)
    std::vector<size_t> dyn_bytes;  // bytes to clear at reset (0: scratch)
    int cache_capacity = 0;      // stored solutions per instance asked for by the descriptor (0: automatic)
    // host copies needed to (re)build kernel parameters
    std::vector<double> h_blob;  // blob of instance 0 (or the shared blob)
    const TpiEntry* tpi = nullptr;
    int coop_lanes = 0;  // 0: not the cooperative kernel
    int coop_static = 0; // 1: CoopSuperover compile-time shape
    int rows = 0;        // > 0: warp-per-instance kernel with LU rows in registers, shape index (rows.cu)
    bool rows_ok = true;
    bool has_cache = false;
    int max_nn = 0, max_nelem = 0;
    int kernel_mode = 0;
    std::string kernel_name;
    int64_t launches = 0;
    int64_t n_done = 0;
    // pinned staging for host-pointer runs
    double* h_stage[2] = {nullptr, nullptr};
    double* d_stage_u[2] = {nullptr, nullptr};
    double* d_stage_y[2] = {nullptr, nullptr};
    size_t stage_u_bytes = 0, stage_y_bytes = 0, hstage_bytes = 0;
    cudaStream_t copy_streams[2] = {nullptr, nullptr};
    cudaStream_t compute_stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};


// tpi.cu: thread-per-instance kernels with compile-time shapes
const TpiEntry* find_tpi(const acme::DevModel& dm);
// coop.cu: lanes-per-instance kernel, state in shared memory
int coop_lanes_for(const acmeb200_model* m);
bool coop_static_matches(const acme::DevModel& dm);
cudaError_t launch_coop_kernel(const acmeb200_model* m, const acme::RunArgs& a, cudaStream_t stream);
// rows.cu: warp-per-instance kernel, LU rows in registers
int rows_shape(const acme::DevModel& dm);  // 0: no instantiation, else the shape index stored in acmeb200_model::rows
cudaError_t launch_rows_kernel(const acmeb200_model* m, const acme::RunArgs& a, cudaStream_t stream);
int64_t rows_mirror_doubles(int shape, int cap);  // per-instance size of the kernel's leaf mirror (kernel_rows.cuh)
