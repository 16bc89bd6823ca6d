#!/bin/bash
T=r3e
mkdir -p gpurun_out
for SM in default 100000; do
  if [ $SM = default ]; then unset ACMEB200_ROWS_SMALL_MAX; else export ACMEB200_ROWS_SMALL_MAX=$SM; fi
  SECONDS=0
  timeout 900 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e --sub none > gpurun_out/c4_$SM.json 2> gpurun_out/c4_$SM.err; echo "small_max=$SM exit $? after $SECONDS s"
  python - $SM <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/c4_{sys.argv[1]}.json'))
print(sys.argv[1], d.get('value'), d.get('ms_per_step'), (d.get('config') or {}).get('kernel'), (d.get('newton') or {}).get('mean_iters'), (d.get('parity') or {}).get('max_rel_err'))
PY
done
