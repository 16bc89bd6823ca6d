// The host emulation builds the generic, the thread-per-instance and the rows-in-registers kernels (see
// cuda_runtime.h here); the cooperative kernel (sub-warp masks) reports "not available".
#include "cuda_runtime.h"
#include "hostmodel.h"

int coop_lanes_for(const acmeb200_model*) { return 0; }
bool coop_static_matches(const acme::DevModel&) { return false; }
cudaError_t launch_coop_kernel(const acmeb200_model*, const acme::RunArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
