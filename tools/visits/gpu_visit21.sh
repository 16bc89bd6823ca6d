#!/bin/bash
T=r3b
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $? after $SECONDS s" >> gpurun_out/tests_$T.log
tail -n 4 gpurun_out/tests_$T.log
SECONDS=0
ACMEB200_TPI_WIDE=0 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_z_sample_major.py tests/test_specialise.py -m gpu -x -q -k "not rows and not superover and not full_size" > gpurun_out/tests_narrow_$T.log 2>&1; echo "narrow-build tests exit $? after $SECONDS s" >> gpurun_out/tests_narrow_$T.log
tail -n 3 gpurun_out/tests_narrow_$T.log
for W in 1 0; do
  ACMEB200_TPI_WIDE=$W KB_MODEL=birdie KB_N=4410 KB_B=32768 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
  ACMEB200_TPI_WIDE=$W KB_MODEL=clipper KB_N=8820 timeout 300 python tools/kbench_one.py 2>&1 | tail -1 | tee -a gpurun_out/kb_$T.jsonl
done
