#!/bin/bash
# bench.py with sub-records: a shrunk smoke run first, then the real default line
T=r2b
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --samples 4410 --batch 8192 --c3-batch 8192 --sub-steps 1 > gpurun_out/bench_small_$T.json 2> gpurun_out/bench_small_$T.err; echo "small exit $?"
tail -5 gpurun_out/bench_small_$T.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_small_r2b.json'))
    for k in ('config1','config3','config4','config5'):
        v=d.get(k); print(k, (v.get('error') or v.get('value')) if v else None, v.get('parity',{}).get('max_rel_err') if v else None)
    print('top', d['value'], d['parity'])
except Exception as e: print('parse failed', e)
PY
SECONDS=0; timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $?"
echo "bench wall seconds: $SECONDS"
cat gpurun_out/bench_$T.json
