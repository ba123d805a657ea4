#!/bin/bash
# round-2 evidence: bench with 8 lanes, lanes probe to 16, full ncu captures of the dense-round kernels (full batch, round ~10)
# and of the frame-resident kernel, the launch list of one bench step.  Everything under gpurun_out/ (summaries -> profiles/).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 16 --warmup 8 --inflight 8 > gpurun_out/r2p_bench8.json 2> gpurun_out/r2p_bench8.err; tail -2 gpurun_out/r2p_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e','single_batch','rounds_per_step')})
PY
timeout 600 python scripts/inflight_probe.py 16 2 2>&1 | grep "^1*[12468] " | tail -3
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
for k in sdf_fused_kernel frame_step_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -o gpurun_out/r2p_$k python scripts/prof_closure.py lbfgs > gpurun_out/r2p_ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lbfgs_resident_kernel -c 1 -o gpurun_out/r2p_lbfgs_resident_kernel python scripts/prof_closure.py resident > gpurun_out/r2p_ncu_res.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2p_launches_bench.csv python bench.py --steps 1 --warmup 1 --inflight 1 --cpu-seconds 1 > gpurun_out/r2p_ncu_bench.log 2>&1
ls -la gpurun_out | grep " r2p_"
