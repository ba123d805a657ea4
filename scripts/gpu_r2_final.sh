#!/bin/bash
# end-of-round validation: what the driver runs (GPU suite, smoke, bench both arms), plus the N3 timing
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; tail -4 gpurun_out/r2f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 300 python scripts/n3_time.py 256 32 2>&1 | grep sdf_all_faces
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e','single_batch','rounds_per_step','clocks')})
print(d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['avg_active_frames_per_launch'], d['roofline']['frac'], d['roofline']['traffic'])
PY
