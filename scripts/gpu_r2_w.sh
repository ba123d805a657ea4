#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 16 --warmup 5 > gpurun_out/r2w_bench_n4.json 2> gpurun_out/r2w_bench_n4.err
tail -3 gpurun_out/r2w_bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','e2e','single_batch')}); print(d['cfg5']['ms_total'], d['cfg5']['comm_ms_per_sweep'])
PY
