#!/bin/bash
# end-of-round evidence: full ncu captures of the dense-round kernels (full batch, round ~10) and of the frame-resident
# kernel, the launch list of one bench step, all under gpurun_out/ (summaries are written into profiles/ afterwards)
mkdir -p gpurun_out
for k in sdf_fused_kernel frame_step_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -o gpurun_out/p_$k python scripts/prof_closure.py lbfgs > gpurun_out/p_ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lbfgs_resident_kernel -c 1 -o gpurun_out/p_lbfgs_resident_kernel python scripts/prof_closure.py resident > gpurun_out/p_ncu_res.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/p_launches_bench.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/p_ncu_bench.log 2>&1
ls -la gpurun_out | grep " p_"
