#!/bin/bash
T=r2r
mkdir -p gpurun_out
KB_STIFF=1 KB_MODEL=clipper KB_N=2205 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tpi -s 3 -c 1 -o gpurun_out/prof_stiff_$T python tools/kbench_one.py > gpurun_out/ncu_stiff_$T.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_stiff_$T.log
KB_STIFF=1 KB_SOLVER="HomotopySolver{SimpleSolver}" KB_MODEL=clipper KB_N=2205 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tpi -s 3 -c 1 -o gpurun_out/prof_stiffH_$T python tools/kbench_one.py > gpurun_out/ncu_stiffH_$T.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_stiffH_$T.log
