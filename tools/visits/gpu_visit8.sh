#!/bin/bash
T=r2i
mkdir -p gpurun_out
CONFIGS=4,5 NO_CPU=1 timeout 900 python tests/tools/bench_configs.py > gpurun_out/configs_$T.md 2> gpurun_out/err_$T.log
cat gpurun_out/configs_$T.md
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_$T.log
tail -n 12 gpurun_out/tests_$T.log
