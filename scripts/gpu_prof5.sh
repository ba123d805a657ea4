#!/bin/bash
# full ncu captures of the dense-round kernels at full occupancy (round ~10) and at the tail, + launch list
mkdir -p gpurun_out
for k in sdf_fused_kernel frame_step_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -o gpurun_out/j_$k python scripts/prof_closure.py lbfgs > gpurun_out/j_ncu_$k.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/j_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/j_ncu0.log 2>&1
ls -la gpurun_out | grep " j_"
