#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_resident.py -x -q 2>&1 | tail -30 > gpurun_out/f_tc.txt; cat gpurun_out/f_tc.txt | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/t_pytest.txt; cat gpurun_out/t_pytest.txt | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/f_bench.txt 2> gpurun_out/f_bench.err; tail -c 1500 gpurun_out/f_bench.txt; tail -3 gpurun_out/f_bench.err
