#!/bin/bash
mkdir -p gpurun_out
for p in 1 2 3; do
  MVS_SDF_PASSES=$p timeout 300 python bench.py --steps 3 --warmup 2 --cpu-seconds 1 > gpurun_out/k_p$p.txt 2>/dev/null
  python - <<PY
import json
for line in open('gpurun_out/k_p$p.txt'):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('passes $p', round(d['ms_per_step'],2), d['rounds_per_step'], r['kernel_avg_us_instrumented_step'])
PY
done
