#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 16 --warmup 8 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
tail -3 gpurun_out/r2q_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','e2e','single_batch')}); print(d['cfg5'])
PY
