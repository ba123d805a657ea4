#!/bin/bash
# the other configs through bench.py (8 lanes), and the reference arm on the 64 x 4 config
mkdir -p gpurun_out
timeout 400 python bench.py --steps 16 --warmup 8 --frames 1 --views 8 --cpu-seconds 1 > gpurun_out/r2r_cfg2.json 2> gpurun_out/r2r_cfg2.err
timeout 400 python bench.py --steps 16 --warmup 8 --frames 64 --views 4 --sdf 0 --cpu-seconds 1 > gpurun_out/r2r_cfg3.json 2> gpurun_out/r2r_cfg3.err
timeout 400 python bench.py --steps 16 --warmup 8 --frames 1 --views 4 --vposer 1 > gpurun_out/r2r_cfg1_vposer.json 2> gpurun_out/r2r_cfg1_vposer.err
timeout 400 python bench.py --steps 16 --warmup 8 --frames 256 --views 4 --vposer 1 > gpurun_out/r2r_vposer_256x4.json 2> gpurun_out/r2r_vposer_256x4.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --frames 64 --views 4 --sdf 0 --ref-seconds 70 > gpurun_out/r2r_ref_cfg3.json 2> gpurun_out/r2r_ref_cfg3.err
python - <<'PY'
import json
for n in ("cfg2","cfg3","cfg1_vposer","vposer_256x4","ref_cfg3"):
    try:
        d=json.loads(open('gpurun_out/r2r_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['value']), round(d['ms_per_step'],2), d.get('single_batch',{}).get('value'), d.get('single_batch',{}).get('ms_per_step'), d.get('e2e',{}).get('value'))
    except Exception as e:
        print(n, "failed", e)
PY
tail -2 gpurun_out/r2r_*.err | tail -20
