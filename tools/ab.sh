#!/bin/bash
# A/B kernel timing of two library builds, interleaved: tools/ab.sh libA libB [reps]
A=$1; B=$2; R=${3:-2}
for i in $(seq $R); do
  for L in $A $B; do
    for M in clipper birdie sallenkey; do
      N=8820; [ $M = birdie ] && N=4410
      ACMEB200_LIB=$L KB_MODEL=$M KB_N=$N python tools/kbench_one.py 2>&1 | tail -1
    done
  done
done
