#!/bin/bash
mkdir -p gpurun_out
MVS_TRACE_HOST=1 timeout 900 python scripts/inflight_probe.py 8 2 > gpurun_out/r2o_probe.log 2>&1
grep -v Warn gpurun_out/r2o_probe.log | grep "^[1248] " ; grep "host loop" gpurun_out/r2o_probe.log | awk 'NR%6==0' | head -12
timeout 900 python bench.py --steps 8 --warmup 4 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -3 gpurun_out/r2o_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e','single_batch','rounds_per_step')})
print(d['roofline']['avg_launch_us'], d['roofline']['avg_active_frames_per_launch'], d['roofline']['frac'])
PY
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
