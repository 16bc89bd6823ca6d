#!/bin/bash
# k-d tree cache, second timing + ncu captures of the birdie (thread-per-instance) and superover (warp-per-instance) kernels
T=r2d
mkdir -p gpurun_out
timeout 300 python tools/kbench_cfg3.py > gpurun_out/kb_$T.jsonl 2> gpurun_out/err_$T.log
KB_MODEL=birdie KB_B=32768 KB_N=22050 timeout 300 python tools/kbench_one.py >> gpurun_out/kb_$T.jsonl 2>> gpurun_out/err_$T.log
CONFIGS=4 NO_CPU=1 timeout 900 python tests/tools/bench_configs.py > gpurun_out/configs_$T.md 2>> gpurun_out/err_$T.log
cat gpurun_out/kb_$T.jsonl gpurun_out/configs_$T.md
KB_WARM=44100 KB_N=2205 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tpi -s 2 -c 1 -f -o gpurun_out/prof_birdie_$T python tools/birdie_prof.py > gpurun_out/ncu_birdie_$T.log 2>&1
KB_B=1024 KB_N=44100 KB_PROF_N=600 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows -s 3 -c 1 -f -o gpurun_out/prof_rows_$T python tools/rows_prof2.py > gpurun_out/ncu_rows_$T.log 2>&1
tail -3 gpurun_out/ncu_birdie_$T.log gpurun_out/ncu_rows_$T.log; ls -la gpurun_out/*.ncu-rep; tail -5 gpurun_out/err_$T.log
