#!/bin/bash
T=r2z
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$T.log 2>&1; echo "tests exit $? after $SECONDS s" >> gpurun_out/tests_$T.log
tail -n 4 gpurun_out/tests_$T.log
timeout 900 python tools/ncu_refresh.py 2>&1 | tail -n 4
SECONDS=0
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $? after $SECONDS s"
tail -n 3 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2z.json'))
print('top', d['value'], d['parity']['max_rel_err'], 'e2e', d['e2e']['value'], 'traffic', d['roofline']['traffic'], d.get('roofline_fp64',{}).get('frac'))
for k in ('config1','config3','config4','config5'):
    v=d.get(k) or {}
    print(k, v.get('error') or v.get('value'), (v.get('parity') or {}).get('max_rel_err'), (v.get('roofline') or {}).get('frac'), (v.get('newton') or {}).get('mean_iters'), (v.get('newton') or {}).get('stored_solutions_rank0'))
PY
