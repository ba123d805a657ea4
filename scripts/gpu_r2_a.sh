#!/bin/bash
# round 2, first GPU call: whole GPU suite (incl. the new reference-kernel pin), bench baseline with the active-count trace
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sdf_refpin.py 2>&1 | tail -30 > gpurun_out/r2a_tests.log
timeout 900 python -m pytest tests/test_gpu_sdf_refpin.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2a_refpin.log
MVS_TRACE_NA=1 timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.log; tail -40 gpurun_out/r2a_refpin.log; head -c 600 gpurun_out/r2a_bench.json
