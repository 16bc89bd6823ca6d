#!/bin/bash
T=r2w
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_z_sample_major.py -m gpu -x -q -k "config2 or config5 or G1 or learning_cache or solver_state or state_persists or frozen_cache_lookup or K3 or K7 or odd_lengths or per_instance or sample" 2>&1 | tail -n 4
for L in fin lpw; do
  ACMEB200_LIB=tools/libs/lib_$L.so KB_MODEL=clipper KB_N=8820 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
  ACMEB200_LIB=tools/libs/lib_$L.so KB_MODEL=birdie KB_N=4410 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
for W in 8 4; do
  ACMEB200_TPI_LPW=$W ACMEB200_LIB=tools/libs/lib_lpw.so KB_MODEL=birdie KB_N=4410 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
ACMEB200_LIB=tools/libs/lib_fin.so KB_MODEL=birdie KB_N=4410 KB_B=4096 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
ACMEB200_LIB=tools/libs/lib_lpw.so KB_MODEL=birdie KB_N=4410 KB_B=4096 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
ACMEB200_TPI_LPW=16 ACMEB200_LIB=tools/libs/lib_lpw.so KB_MODEL=clipper KB_N=8820 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
