#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vposer.py tests/test_gpu_seq_demo.py -m gpu -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 8 --warmup 4 --frames 256 --views 4 --vposer 1 > gpurun_out/r2bb_vposer_256x4.json 2> gpurun_out/r2bb_vposer.err; tail -2 gpurun_out/r2bb_vposer.err
timeout 400 python bench.py --steps 8 --warmup 4 --frames 1 --views 4 --vposer 1 > gpurun_out/r2bb_cfg1_vposer.json 2>> gpurun_out/r2bb_vposer.err
python - <<'PY'
import json
for n in ("vposer_256x4","cfg1_vposer"):
    d=json.load(open('gpurun_out/r2bb_%s.json'%n)); print(n, round(d['value']), round(d['ms_per_step'],2), d['single_batch']['value'], d['single_batch']['ms_per_step'])
PY
