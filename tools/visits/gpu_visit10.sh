#!/bin/bash
# 2 GPUs: the plain-C multi-GPU host, then the complete bench line under torchrun
T=r2k
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_c_abi.py -m gpu -x -q 2>&1 | tail -n 3
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_$T.json 2> gpurun_out/bench_2gpu_$T.err; echo "bench exit $? after $SECONDS s"
tail -n 3 gpurun_out/bench_2gpu_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu_r2k.json'))
print('top', d['value'], d['n_gpus'], d['parity']['max_rel_err'], 'e2e', d['e2e']['value'])
for k in ('config1','config3','config4','config5'):
    v=d.get(k) or {}
    print(k, v.get('error') or v.get('value'), (v.get('parity') or {}).get('max_rel_err'), (v.get('gather') or {}), (v.get('newton') or {}).get('mean_iters'))
PY
