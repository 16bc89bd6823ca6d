#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
T=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$T.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $?" >> gpurun_out/tests_$T.log
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --samples 4416 > gpurun_out/bench_under_ncu_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tpi -s 2 -c 1 -f -o gpurun_out/prof_tpi_$T \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --samples 2208 > gpurun_out/ncu_full_$T.log 2>&1
# sample-major streams against the default layout: short kernel-only runs of configs 3 and 2, then the full bench line
for L in instance sample; do KB_LAYOUT=$L timeout 300 python tools/kbench_cfg3.py >> gpurun_out/layouts_$T.jsonl 2>> gpurun_out/bench_$T.err; done
timeout 600 python bench.py --layout sample --no-cpu > gpurun_out/bench_smaj_$T.json 2>> gpurun_out/bench_$T.err
tail -3 gpurun_out/tests_$T.log; cat gpurun_out/bench_$T.json; cat gpurun_out/layouts_$T.jsonl
