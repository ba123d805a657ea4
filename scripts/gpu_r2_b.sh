#!/bin/bash
# round 2, second GPU call: SDF pin diagnostics, reference arms (host cores 32 / 64 workers, torch-CUDA), end-to-end fit tests
mkdir -p gpurun_out
timeout 600 python scripts/sdf_pin_diag.py > gpurun_out/r2b_sdf_pin_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_sdf_refpin.py tests/test_gpu_fit_e2e.py -m gpu -q -s 2>&1 | grep -v Warning | tail -120 > gpurun_out/r2b_tests.log
nproc > gpurun_out/r2b_host.txt; lscpu | head -20 >> gpurun_out/r2b_host.txt
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 --ref-seconds 60 --ref-workers 32 > gpurun_out/r2b_ref_cpu32.json 2> gpurun_out/r2b_ref_cpu32.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 --ref-seconds 60 --ref-workers 64 > gpurun_out/r2b_ref_cpu64.json 2> gpurun_out/r2b_ref_cpu64.err
timeout 400 python bench.py --impl reference --ref-device cuda --steps 3 --warmup 1 --ref-seconds 60 > gpurun_out/r2b_ref_cuda.json 2> gpurun_out/r2b_ref_cuda.err
timeout 400 python bench.py --impl reference --ref-device cuda --sdf 0 --steps 3 --warmup 1 --ref-seconds 45 > gpurun_out/r2b_ref_cuda_nosdf.json 2> gpurun_out/r2b_ref_cuda_nosdf.err
tail -30 gpurun_out/r2b_tests.log; head -c 400 gpurun_out/r2b_ref_cpu64.json; head -c 400 gpurun_out/r2b_ref_cuda.json
