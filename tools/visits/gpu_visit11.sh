#!/bin/bash
# 8 GPUs: config 2 (kernel only) + the north-star config 4 (8192 superover instances sharded, NCCL gather)
T=r2l
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --sub 4,5 > gpurun_out/bench_8gpu_$T.json 2> gpurun_out/bench_8gpu_$T.err; echo "bench exit $? after $SECONDS s"
tail -n 2 gpurun_out/bench_8gpu_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_8gpu_r2l.json'))
print('top', d['value'], d['n_gpus'], 'e2e', (d.get('e2e') or {}).get('value'))
for k in ('config4','config5'):
    v=d.get(k) or {}
    print(k, v.get('error') or v.get('value'), (v.get('parity') or {}).get('max_rel_err'), (v.get('gather') or {}), (v.get('newton') or {}).get('mean_iters'), (v.get('e2e') or {}).get('value'))
PY
