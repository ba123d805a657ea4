#!/bin/bash
# other BASELINE.json configs as bench lines (parity-test cases; kept under profiles/ for the BASELINE.md table)
mkdir -p gpurun_out
timeout 300 python bench.py --frames 64 --views 4 --sdf 0 --steps 5 --warmup 3 > gpurun_out/c_cfg3.txt 2>/dev/null; tail -c 300 gpurun_out/c_cfg3.txt
timeout 300 python bench.py --frames 1 --views 8 --sdf 1 --steps 5 --warmup 3 > gpurun_out/c_cfg2.txt 2>/dev/null; tail -c 300 gpurun_out/c_cfg2.txt
timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/c_cfg4.txt 2>/dev/null; tail -c 300 gpurun_out/c_cfg4.txt
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -2
