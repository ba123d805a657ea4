#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_sdf_refpin.py -m gpu -q -s > gpurun_out/r2aa_tests.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed\|all-faces closure" gpurun_out/r2aa_tests.log | cut -c1-300 | head -20
timeout 300 python scripts/n3_time.py 256 8 2>&1 | grep sdf_all_faces
