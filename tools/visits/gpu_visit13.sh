#!/bin/bash
T=r2n
mkdir -p gpurun_out
timeout 300 python tools/kbench_cfg3.py > gpurun_out/kb_$T.jsonl 2> gpurun_out/err_$T.log
KB_MODEL=birdie KB_B=32768 KB_N=22050 timeout 300 python tools/kbench_one.py >> gpurun_out/kb_$T.jsonl 2>> gpurun_out/err_$T.log
cat gpurun_out/kb_$T.jsonl
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $? after $SECONDS s" >> gpurun_out/tests_$T.log
tail -n 6 gpurun_out/tests_$T.log
