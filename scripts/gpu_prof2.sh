#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_frame_kernel -s 40 -c 1 -o gpurun_out/d_sdf_frame python scripts/prof_closure.py lbfgs > gpurun_out/d_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_step_kernel -s 40 -c 1 -o gpurun_out/d_frame_step python scripts/prof_closure.py lbfgs > gpurun_out/d_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/d_ncu3.log 2>&1
ls -la gpurun_out | grep d_
