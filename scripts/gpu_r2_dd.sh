#!/bin/bash
# re-capture of the dense-round kernels on the final code (TMA-store epilogue, no dense v_posed write)
mkdir -p gpurun_out
for k in sdf_fused_kernel frame_step_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -o gpurun_out/r2dd_$k python scripts/prof_closure.py lbfgs > gpurun_out/r2dd_ncu_$k.log 2>&1
done
ls -la gpurun_out | grep "r2dd_.*rep"
