#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/round_times.py 256 64 1 2>&1 | grep "^B="
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_tests.log 2>&1; tail -3 gpurun_out/r2z_tests.log
timeout 600 python bench.py --steps 16 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -2 gpurun_out/r2z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','single_batch')})
print(d['roofline']['kernel_avg_us_instrumented_step'])
PY
