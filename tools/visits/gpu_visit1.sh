#!/bin/bash
# round 2, first GPU visit: the -m gpu suite (new HC sweep tests), stream layouts, per-config baselines
T=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$T.txt 2>&1
nproc > gpurun_out/nproc_$T.txt; numactl -H >> gpurun_out/nproc_$T.txt 2>&1; lscpu | head -30 >> gpurun_out/nproc_$T.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_$T.log
for L in instance sample; do KB_LAYOUT=$L timeout 300 python tools/kbench_cfg3.py >> gpurun_out/layouts_$T.jsonl 2>> gpurun_out/err_$T.log; done
CONFIGS=4,5 NO_CPU=1 timeout 600 python tests/tools/bench_configs.py > gpurun_out/configs_$T.md 2>> gpurun_out/err_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
tail -3 gpurun_out/tests_$T.log; cat gpurun_out/layouts_$T.jsonl; cat gpurun_out/configs_$T.md; cat gpurun_out/smoke_$T.log | tail -2
