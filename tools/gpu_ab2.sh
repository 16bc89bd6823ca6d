#!/bin/bash
T=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for L in "$@"; do
  ACMEB200_LIB=tools/libs/lib_$L.so KB_MODEL=birdie KB_N=4410 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
  ACMEB200_LIB=tools/libs/lib_$L.so KB_WARM=44100 KB_N=4410 timeout 300 python tools/birdie_prof.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
  [ $rep = 1 ] && ACMEB200_LIB=tools/libs/lib_$L.so KB_MODEL=clipper KB_N=8820 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
done
