#!/bin/bash
T=r2g
mkdir -p gpurun_out
KB_B=1024 KB_N=44100 KB_PROF_N=600 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -f -o gpurun_out/prof_rows_$T python tools/rows_prof2.py > gpurun_out/ncu_rows_$T.log 2>&1
KB_MODEL=clipper KB_N=4410 timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --clock-control none --import-source on -k regex:k_tpi -s 4 -c 1 -f -o gpurun_out/prof_clipper_$T python tools/kbench_one.py > gpurun_out/ncu_clipper_$T.log 2>&1
ls -la gpurun_out/
