#!/bin/bash
# launch list of a whole SDF stage (head and tail) + full captures of the round kernels at full and low occupancy
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/h_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/h_ncu0.log 2>&1
for k in frame_step_kernel vertex_bwd_kernel sdf_part_kernel skin_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -o gpurun_out/h_$k python scripts/prof_closure.py lbfgs > gpurun_out/h_ncu_$k.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 400 -c 1 -o gpurun_out/h_low_$k python scripts/prof_closure.py lbfgs > gpurun_out/h_ncu_low_$k.log 2>&1
done
ls -la gpurun_out | grep h_
