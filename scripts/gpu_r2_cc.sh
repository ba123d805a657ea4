#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/round_times.py 256 1 2>&1 | grep "^B="
timeout 400 python bench.py --steps 8 --warmup 4 --frames 64 --views 4 --sdf 0 --cpu-seconds 1 > gpurun_out/r2cc_cfg3.json 2> gpurun_out/r2cc_cfg3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2cc_cfg3.json')); print('cfg3', round(d['value']), d['single_batch']['value'], d['single_batch']['ms_per_step'])
PY
bash scripts/gpu_r2_final.sh
