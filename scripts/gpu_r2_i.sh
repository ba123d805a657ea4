#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py -m gpu -q 2>&1 | grep -v Warning | tail -60 > gpurun_out/r2i_resident.log; grep -n "^E " gpurun_out/r2i_resident.log | head -8
timeout 600 python scripts/e2e_sdf_diag.py > gpurun_out/r2i_e2e_diag.log 2>&1; tail -3 gpurun_out/r2i_e2e_diag.log
