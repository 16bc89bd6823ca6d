#!/bin/bash
T=r2z
mkdir -p gpurun_out
KB_MODEL=clipper KB_N=2208 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tpi -s 3 -c 1 -o gpurun_out/prof_clipper_$T python tools/kbench_one.py > gpurun_out/ncu_clipper_$T.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_clipper_$T.log
KB_WARM=44100 KB_N=1102 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tpi -s 2 -c 1 -o gpurun_out/prof_birdie_$T python tools/birdie_prof.py > gpurun_out/ncu_birdie_$T.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/ncu_birdie_$T.log
SECONDS=0; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_$T.log 2>&1; echo "launch list exit $? after $SECONDS s"
tail -n 3 gpurun_out/launches_$T.csv | cut -c1-300
