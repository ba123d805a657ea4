#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_resident.py::test_dense_regime_lbfgs_matches_batched_lbfgs -m gpu -q -s > gpurun_out/r2n_tests.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r2n_tests.log | cut -c1-600 | head -20
