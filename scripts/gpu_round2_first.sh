#!/bin/bash
# first GPU call of the next round: everything that was written after round 1's GPU minutes ran out, then the usual
# test + bench pass.  ~6 GPU-minutes.
mkdir -p gpurun_out
timeout 120 python scripts/init_time.py 256 8 > gpurun_out/g_init.txt 2>&1; tail -4 gpurun_out/g_init.txt
timeout 900 python -m pytest tests -q -m gpu -rxX > gpurun_out/g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.txt; tail -8 gpurun_out/g_pytest.txt
timeout 600 python bench.py > gpurun_out/g_bench.txt 2> gpurun_out/g_bench.err; tail -c 400 gpurun_out/g_bench.txt; tail -2 gpurun_out/g_bench.err
