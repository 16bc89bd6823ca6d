// The 255-register ("wide") build of the non-linear thread-per-instance kernels for batches that need at most
// ACME_TPI_WIDE_MINB CTAs per SM (kernel_tpi.cuh, tpi_launch.cuh): a translation unit of its own so that it compiles in
// parallel with tpi.cu.
#include <cuda_runtime.h>

#include "tpi_shapes.h"

using namespace acme;

#define X(CFG) template cudaError_t acme::launch_tpi_k<CFG, true>(const acmeb200_model*, const RunArgs&, const TpiMats<CFG>&, const SolverCfg&, const DevSub&, const TpiMaps&, int64_t, int, cudaStream_t);
ACME_TPI_WIDE_SHAPES(X)
#undef X
