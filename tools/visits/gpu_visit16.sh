#!/bin/bash
T=r2q
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or G1 or learning_cache or solver_state or state_persists or frozen_cache_lookup or K3" 2>&1 | tail -n 6
ACMEB200_LIB=tools/libs/lib_prof.so timeout 600 python tools/tail_prof.py > gpurun_out/tail_$T.jsonl 2> gpurun_out/tail_$T.err; echo "tail exit $?"
tail -n 3 gpurun_out/tail_$T.err
for L in tools/libs/lib_base.so tools/libs/lib_sm1.so tools/libs/lib_base.so tools/libs/lib_sm1.so; do
  ACMEB200_LIB=$L KB_MODEL=clipper KB_N=8820 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
  ACMEB200_LIB=$L KB_MODEL=birdie KB_N=4410 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
cut -c1-900 gpurun_out/tail_$T.jsonl
