#!/bin/bash
T=r2j
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $? after $SECONDS s" >> gpurun_out/tests_$T.log
tail -n 12 gpurun_out/tests_$T.log
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $? after $SECONDS s"
tail -n 3 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2j.json'))
print('top', d['value'], d['parity']['max_rel_err'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
for k in ('config1','config3','config4','config5'):
    v=d.get(k) or {}
    print(k, v.get('error') or v.get('value'), (v.get('parity') or {}).get('max_rel_err'), (v.get('roofline') or {}).get('frac'))
PY
timeout 900 python tools/ncu_refresh.py 2>&1 | tail -n 5
