#!/bin/bash
T=r2m
mkdir -p gpurun_out
KB_B=1024 KB_N=44100 KB_PROF_N=600 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -f -o gpurun_out/prof_rows_$T python tools/rows_prof2.py > gpurun_out/ncu_rows_$T.log 2>&1
tail -n 4 gpurun_out/ncu_rows_$T.log
KB_WARM=44100 KB_N=2205 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tpi -s 2 -c 1 -f -o gpurun_out/prof_birdie_$T python tools/birdie_prof.py > gpurun_out/ncu_birdie_$T.log 2>&1
tail -n 3 gpurun_out/ncu_birdie_$T.log
ls -la gpurun_out/*.ncu-rep
