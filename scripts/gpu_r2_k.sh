#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -25 > gpurun_out/r2k_tests.log; tail -12 gpurun_out/r2k_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 600 gpurun_out/r2k_bench.json
