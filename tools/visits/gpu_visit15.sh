#!/bin/bash
T=r2p
mkdir -p gpurun_out
ACMEB200_LIB=tools/libs/lib_prof.so timeout 600 python tools/tail_prof.py > gpurun_out/tail_$T.jsonl 2> gpurun_out/tail_$T.err; echo "tail exit $?"
tail -n 3 gpurun_out/tail_$T.err
for L in tools/libs/lib_base.so tools/libs/lib_m7.so tools/libs/lib_base.so tools/libs/lib_m7.so; do
  ACMEB200_LIB=$L KB_MODEL=clipper KB_N=8820 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
cut -c1-600 gpurun_out/tail_$T.jsonl
