#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sdf_bins_kernel -s 1 -c 1 -o gpurun_out/r2x_sdf_bins_kernel python scripts/prof_n3.py > gpurun_out/r2x_ncu_bins.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sdf_fused_kernel -s 1 -c 1 -o gpurun_out/r2x_sdf_fused_all python scripts/prof_n3.py > gpurun_out/r2x_ncu_fused.log 2>&1
ls -la gpurun_out | grep r2x_
