#!/bin/bash
mkdir -p gpurun_out
for k in frame_step_kernel vertex_bwd_kernel sdf_part_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -o gpurun_out/g_$k python scripts/prof_closure.py lbfgs > gpurun_out/g_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/g_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/g_ncu3.log 2>&1
ls -la gpurun_out | grep g_
