// The host emulation builds the generic and the rows-in-registers kernels only (see cuda_runtime.h here):
// the other kernel families (TMA tensor maps, sub-warp masks) report "not available".
#include "cuda_runtime.h"
#include "hostmodel.h"

const TpiEntry* find_tpi(const acme::DevModel&) { return nullptr; }
int coop_lanes_for(const acmeb200_model*) { return 0; }
bool coop_static_matches(const acme::DevModel&) { return false; }
cudaError_t launch_coop_kernel(const acmeb200_model*, const acme::RunArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
