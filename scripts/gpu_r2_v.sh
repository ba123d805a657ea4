#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 1 --frames 32 --inflight 2 --sdf-all-faces 1 --cpu-seconds 1 > gpurun_out/r2v_bench_allfaces.json 2> gpurun_out/r2v_bench_allfaces.err; tail -2 gpurun_out/r2v_bench_allfaces.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench_allfaces.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','single_batch','rounds_per_step')}); print(d['config'].get('sdf_semantics')); print(d['roofline']['kernel_time_share_of_step'])
PY
