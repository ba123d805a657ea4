#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lbfgs_resident_kernel -c 1 -o gpurun_out/p_lbfgs_resident_kernel python scripts/prof_closure.py resident > gpurun_out/p_ncu_res.log 2>&1
ls -la gpurun_out | grep p_lbfgs
