#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/round_times.py 256 64 5 1 2>&1 | tail -4
timeout 600 python scripts/sdf_pin_diag.py > gpurun_out/r2h_sdf_pin_diag.log 2>&1; tail -2 gpurun_out/r2h_sdf_pin_diag.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
head -c 300 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
