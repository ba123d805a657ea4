#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/inflight_probe.py 4 3 2>&1 | grep -v Warn | tail -8
timeout 1200 python -m pytest tests/test_gpu_fit_e2e.py::test_four_stage_fit_with_sdf_matches_live_reference_run tests/test_gpu_resident.py::test_dense_regime_lbfgs_matches_batched_lbfgs tests/test_gpu_seq_demo.py tests/test_gpu_vposer.py::test_lbfgs_step_in_latent_space tests/test_gpu_zz_init.py -m gpu -q -s > gpurun_out/r2m_tests.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed\|demo fit" gpurun_out/r2m_tests.log | cut -c1-300 | head -40
