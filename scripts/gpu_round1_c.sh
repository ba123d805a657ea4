#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_closure.csv python scripts/prof_closure.py closure > gpurun_out/c_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vertex_bwd_kernel -s 4 -c 1 -o gpurun_out/c_vertex_bwd python scripts/prof_closure.py closure > gpurun_out/c_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vertex_fwd_kernel -s 4 -c 1 -o gpurun_out/c_vertex_fwd python scripts/prof_closure.py closure > gpurun_out/c_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame_bwd_kernel -s 4 -c 1 -o gpurun_out/c_frame_bwd python scripts/prof_closure.py closure > gpurun_out/c_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/c_ncu5.log 2>&1
ls -la gpurun_out | tail -12
