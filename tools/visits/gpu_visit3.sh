#!/bin/bash
# the reference's k-d tree cache on the device: first timing
T=r2c
mkdir -p gpurun_out
timeout 300 python tools/kbench_cfg3.py > gpurun_out/kb_$T.jsonl 2> gpurun_out/err_$T.log
KB_MODEL=birdie KB_B=32768 KB_N=22050 timeout 300 python tools/kbench_one.py >> gpurun_out/kb_$T.jsonl 2>> gpurun_out/err_$T.log
CONFIGS=4,5 NO_CPU=1 timeout 900 python tests/tools/bench_configs.py > gpurun_out/configs_$T.md 2>> gpurun_out/err_$T.log
cat gpurun_out/kb_$T.jsonl gpurun_out/configs_$T.md; tail -5 gpurun_out/err_$T.log
