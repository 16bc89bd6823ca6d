#!/bin/bash
T=r2h
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_$T.log
tail -n 15 gpurun_out/tests_$T.log
timeout 300 python tools/kbench_cfg3.py > gpurun_out/kb_$T.jsonl 2> gpurun_out/err_$T.log
KB_MODEL=birdie KB_B=32768 KB_N=22050 timeout 300 python tools/kbench_one.py >> gpurun_out/kb_$T.jsonl 2>> gpurun_out/err_$T.log
cat gpurun_out/kb_$T.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
