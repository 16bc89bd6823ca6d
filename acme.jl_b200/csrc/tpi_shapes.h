// The compile-time shapes the library carries for the thread-per-instance kernels (kernel_tpi.cuh): the BASELINE circuits
// (SURVEY.md section 8 size table).  Two translation units build them: tpi.cu (registry, linear shape, the 128-register
// build of the non-linear shapes) and tpi_wide.cu (their 255-register build for small batches), compiled in parallel.
#pragma once
#include "tpi_launch.cuh"

namespace acme {
using CfgDiodeClipper = TpiCfg<1, 1, 1, 1, Diode, Diode>;  // examples/diodeclipper.jl
using CfgSallenKey = TpiCfg<2, 1, 1, 0>;                   // examples/sallenkey.jl (linear)
using CfgBirdieFixed = TpiCfg<3, 1, 1, 2, Bjt>;            // examples/birdie.jl, vol baked in
using CfgBirdieVol = TpiCfg<3, 2, 1, 3, Bjt, Pot>;         // examples/birdie.jl, vol as input
}  // namespace acme

#define ACME_TPI_WIDE_SHAPES(X) X(acme::CfgDiodeClipper) X(acme::CfgBirdieFixed) X(acme::CfgBirdieVol)
